"""GPU: parity at the sizes BASELINE.json quotes (the fixtures and the ragged-batch tests use 40-row tables and B <= 1000).

* configs[1]: NASRec-Small supernet TRAINING STEP through the C++ executor, B = 512, 0.5 M-row tables, Zipf ids (hot rows
  with hundreds of duplicates in a batch) -- logits / loss / clip norm / updated weights and the touched embedding rows
  against the oracle's step on the same weights and batch;
* configs[0]: Criteo NASRec-Full best model (fixed, LN off), B = 256, capped tables -- forward + backward;
* configs[2]: NASRec-Full supernet scoring at the evaluation batch B = 8192 -- logits only (the oracle's forward takes a few
  seconds on the host cores).
Tolerances as everywhere: logits and log-loss 1e-5 relative, bit-exact row sets, gradient norms 5e-4 / 5e-2 near a kink."""
import numpy as np
import pytest
import torch

from nasrec_b200 import SuperNet, ops_config_lib
from nasrec_b200.native import NativeNet, NativeTrainer
from nasrec_b200.utils.train_utils import init_weights
from oracle import nasrec_oracle as orc
from tests.helpers import flipped_relu, load_golden, rel_err, relu_kink_candidates, relu_kink_margin

pytestmark = pytest.mark.gpu

CRITEO = [1461, 584, 10131227, 2202609, 306, 25, 12518, 634, 4, 93146, 5684, 8351593, 3195, 28, 14993, 5461307, 11,
          5653, 2174, 5, 7046548, 19, 16, 286182, 106, 142573]
CAP = 500000


def _cpu_state(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def _step_mismatch(tr, m, logits, loss, ref, rl, rloss, rnorm, new, sd, sets):
    """(None, 0) when the executor's step equals the oracle's at the strict tolerances, else (first miss, worst excess =
    largest ratio of a difference to its tolerance over everything checked)."""
    first, worst = None, 0.0

    def check(name, diff, tol):
        nonlocal first, worst
        if diff >= tol:
            first = first or name
            worst = max(worst, diff / tol)

    check("logits", rel_err(logits.cpu().numpy(), rl.numpy()), 1e-5)
    check("loss", abs(float(loss.item()) - rloss), 1e-5 * max(1.0, abs(rloss)))
    check("clip norm", abs(float(tr.last_total_norm.item()) - rnorm), 5e-4 * max(1.0, rnorm))
    for f in (0, 2, 9, 18, 25):
        k = "_embedding.%d.weight" % f
        moved = np.nonzero(np.abs((new[k] - sd[k]).numpy()).sum(1))[0]
        if not set(moved.tolist()) <= set(sets[f].tolist()):
            return k + " moved rows outside the batch", float("inf")
        check(k, float((new[k] - ref.params[k].detach()).abs().max()), 1e-5)
    for k in ("_final.weight", "_blocks.5._nodes.2._linear.weight", "_blocks.3.project_emb_dim.weight"):
        if k in new:
            check(k, float((new[k] - ref.params[k].detach()).abs().max()), 1e-5)
    return first, worst


class _flips:
    """Several flipped_relu contexts at once."""

    def __init__(self, flips):
        self.flips = flips

    def __enter__(self):
        import torch as _t
        self.orig = _t.relu
        count = [0]
        by_call = {}
        for call, idx, val in self.flips:
            by_call.setdefault(call, []).append((idx, val))

        def relu(x):
            c = count[0]
            count[0] += 1
            if c in by_call:
                off = _t.zeros_like(x).flatten()
                for idx, val in by_call[c]:
                    off[idx] = -2.0 * val
                x = x + off.view_as(x)
            return self.orig(x)

        _t.relu = relu
        return self

    def __exit__(self, *exc):
        import torch as _t
        _t.relu = self.orig
        return False


def test_small_supernet_training_step_B512_capped_tables():
    """Two steps at the headline size against the oracle, STRICT (logits / loss 1e-5, norm 5e-4, weights 1e-5 absolute).
    Conditioning: a step evaluates ~10^7 ReLU inputs and the closest ones to zero are ~1e-8 .. 5e-7 away (measured: -2.0e-8
    in step 0 of this very case, six within 5e-7 in step 1).  Two correct fp32 implementations can put such a unit on
    different sides; that switches one unit's gradient for one sample and, through first-step Adagrad (lr / eps = 12),
    moves that sample's rows by up to ~1e-4 (profiles/r02_notes.md: changing only the split-K order of ONE forward GEMM
    does it, and the oracle with that one unit flipped reproduces the CUDA weights to 2.5e-7).  The test therefore allows
    exactly this and nothing else: when the strict comparison fails, near-zero ReLU inputs of the oracle (|x| < 6e-6 of the
    call's median, at most six candidates) are pushed across zero greedily -- the flip that brings the oracle closest is
    kept, at most three in total -- and the strict comparison must hold against that oracle run, which then also is the
    state the next step starts from.  If the candidates tried do not account for the difference, the bounded-outlier
    check at the end of the loop applies."""
    import copy
    ne = [min(x, CAP) for x in CRITEO]
    torch.manual_seed(3)
    np.random.seed(3)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["autoctr"], use_layernorm=True, num_embeddings=ne,
                 path_sampling_strategy="default", anypath_choice="binomial-0.5", supernet_training_steps=0).to("cuda")
    m.materialize(13)
    m.apply(init_weights)
    cfg = dict(ops="autoctr", use_layernorm=True, fixed=False, num_blocks=7)
    tr = NativeTrainer(m, lr=0.12)
    ref = orc.OracleTrainer(_cpu_state(m), cfg, lr=0.12)          # one oracle run: its Adagrad accumulators carry over too
    for step in range(2):
        sd = _cpu_state(m)
        int_x, cat_x, y = orc.synth_batch(512, 13, ne, seed=1200 + step, zipf=True)
        logits, loss = tr.step(int_x.cuda(), cat_x.cuda(), y.cuda())
        assert tr.net is not None, tr.fallback_reason
        new = _cpu_state(m)
        sets = orc.embedding_row_sets(cat_x.numpy())
        before = copy.deepcopy(ref)
        rl, rloss, rnorm = ref.step(m.choice, int_x, cat_x, y)
        miss, score = _step_mismatch(tr, m, logits, loss, ref, rl, rloss, rnorm, new, sd, sets)
        if miss is not None:
            osd = {k: v.detach() for k, v in before.params.items()}
            cands = [(call, idx, val) for _, call, idx, val in relu_kink_candidates(osd, cfg, m.choice, int_x, cat_x)]
            kept, log = [], [("none", miss, score)]
            while miss is not None and len(kept) < 3 and cands:
                best = None
                for c in cands:
                    t = copy.deepcopy(before)
                    with _flips(kept + [c]):
                        out = t.step(m.choice, int_x, cat_x, y)
                    mi, sc = _step_mismatch(tr, m, logits, loss, t, *out, new, sd, sets)
                    if best is None or sc < best[0]:
                        best = (sc, mi, c, t)          # keep one candidate state at a time (each is ~1 GB of host memory)
                    del t
                    if mi is None:
                        break
                sc, mi, c, t = best
                log.append((c[:2], mi, sc))
                if sc >= score:
                    break                                  # no flip helps: not a kink effect
                kept.append(c)
                cands.remove(c)
                miss, score, ref = mi, sc, t
            if miss is not None:
                # More units sit within the GPU-vs-CPU rounding distance (~1e-6) of zero than can be tried one oracle step
                # at a time (about fifty per step at this size).  Last resort, still a hard, STRUCTURAL bound: logits, loss
                # and clip norm strict as above; a flipped unit changes the gradient of ONE sample, so every table row
                # that is off must be a row of one of at most four samples (greedy cover over the checked tables); at most
                # 2 % of a table's touched rows off, none by more than 5e-3; dense weights: < 5 % of the elements off (a
                # rank-1 outlier per flip), none by more than 5e-3.
                assert not miss.startswith(("logits", "loss", "clip")) and "outside" not in miss, (step, log)
                cat = cat_x.numpy()
                flagged = set()
                for f in (0, 2, 9, 18, 25):
                    k = "_embedding.%d.weight" % f
                    row_err = (new[k] - ref.params[k].detach()).abs().max(1).values
                    off_rows = torch.nonzero(row_err >= 1e-5).flatten().tolist()
                    assert len(off_rows) <= max(2, len(sets[f]) // 50), (step, k, log)
                    assert float(row_err.max()) < 5e-3, (step, k, log)
                    flagged |= {(f, r) for r in off_rows}
                cover = []
                while flagged and len(cover) < 5:
                    gain = [sum((f, int(cat[bi, f])) in flagged for f in (0, 2, 9, 18, 25)) for bi in range(cat.shape[0])]
                    bi = int(np.argmax(gain))
                    if gain[bi] == 0:
                        break
                    cover.append(bi)
                    flagged -= {(f, int(cat[bi, f])) for f in (0, 2, 9, 18, 25)}
                assert not flagged and len(cover) <= 4, (step, "off rows are not the rows of a few samples", cover, sorted(flagged)[:8], log)
                for k in ("_final.weight", "_blocks.5._nodes.2._linear.weight", "_blocks.3.project_emb_dim.weight"):
                    if k in new:
                        d = (new[k] - ref.params[k].detach()).abs()
                        assert float(d.max()) < 5e-3 and float((d >= 1e-5).double().mean()) < 0.05, (step, k, log)


def test_criteo_full_best_B256_capped_tables():
    meta, _ = load_golden("fixed_best")
    mm = meta["models"]["criteo_xlarge"]
    ne = [min(x, CAP) for x in CRITEO]
    torch.manual_seed(4)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=False, num_embeddings=ne,
                 path_sampling_strategy="fixed-path", fixed=True, fixed_choice=mm["choice"]).to("cuda")
    m.materialize(13)
    m.apply(init_weights)
    sd = _cpu_state(m)
    cfg = dict(ops="xlarge", use_layernorm=False, fixed=True, num_blocks=7)
    int_x, cat_x, y = orc.synth_batch(256, 13, ne, seed=1300, zipf=True)
    m.zero_grad(set_to_none=True)
    logits = m(int_x.cuda(), cat_x.cuda())
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.cuda())
    loss.backward()
    lr, lo, gr = orc.loss_and_grads(sd, cfg, mm["choice"], int_x, cat_x, y)
    assert rel_err(logits.detach().cpu().numpy(), lr.numpy()) < 1e-5
    assert abs(float(loss) - float(lo)) < 1e-5 * max(1.0, abs(float(lo)))
    gtol = 5e-4 if relu_kink_margin(sd, cfg, mm["choice"], int_x, cat_x) >= 6e-6 else 5e-2
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    for n, g in gr.items():
        gn = float(g.double().norm())
        if gn == 0.0 or n.startswith("_embedding"):
            continue
        assert abs(float(grads[n].double().norm()) - gn) <= gtol * max(gn, 1e-3) + 1e-7, n
    sets = orc.embedding_row_sets(cat_x.numpy())
    for f in (0, 2, 11, 25):
        got = np.nonzero(np.abs(grads["_embedding.%d.weight" % f].numpy()).sum(1))[0]
        want = np.nonzero(np.abs(gr["_embedding.%d.weight" % f].numpy()).sum(1))[0]
        assert got.tolist() == want.tolist()
        assert set(got.tolist()) <= set(sets[f].tolist())


def test_xlarge_scoring_B8192_logits():
    from nasrec_b200.search import generate_random_choice
    ne = [min(x, CAP) for x in CRITEO]
    torch.manual_seed(5)
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib["xlarge"], use_layernorm=True, num_embeddings=ne,
                 path_sampling_strategy="full-path").to("cuda")
    m.materialize(13)
    m.apply(init_weights)
    m.requires_grad_(False)
    sd = _cpu_state(m)
    cfg = dict(ops="xlarge", use_layernorm=True, fixed=False, num_blocks=7)
    np.random.seed(77)
    cands = [generate_random_choice(7, ops_config_lib["xlarge"]) for _ in range(2)]
    int_x, cat_x, _ = orc.synth_batch(8192, 13, ne, seed=1400, zipf=True)
    net = NativeNet(m, state_of=None, pgrad_bytes=1 << 20)
    z = net.forward_multi([NativeNet.encode_choice(c["macro"], c["micro"]) for c in cands], int_x.cuda().contiguous(),
                          cat_x.cuda().contiguous())
    for k, c in enumerate(cands):
        with torch.no_grad():
            ref = orc.supernet_forward(sd, cfg, c, int_x, cat_x).reshape(-1)
        assert rel_err(z[k].cpu().numpy(), ref.numpy()) < 1e-5, k


@pytest.mark.parametrize("B,zipf", [(4096, True), (4096, False), (16384, True), (40000, True), (1500, True)])
def test_big_sorted_row_reduction_matches_one_cta_kernel_and_oracle(B, zipf):
    """csrc/emb_big.cu (multi-CTA radix-sort reduction, any B) against the one-CTA-per-table kernel (B <= 16384) and the
    oracle's row sets: same unique rows in the same order, same counts; summed gradients equal bit for bit for rows with at
    most 256 duplicates (same ascending-sample order) and to fp32 rounding beyond; an out-of-range id is dropped and flagged."""
    from nasrec_b200 import _lib
    ne = [min(x, CAP) for x in CRITEO]
    F = len(ne)
    _, cat_x, _ = orc.synth_batch(B, 13, ne, seed=1500, zipf=zipf)
    cat = cat_x.cuda()
    cat[7, 3] = ne[3] + 5                      # out of range: must be dropped by both kernels
    g = torch.Generator().manual_seed(1)
    gout = torch.randn(B, F, 16, generator=g).cuda()
    rows = torch.tensor(ne, dtype=torch.int64, device="cuda")

    def run(big):
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        uniq = torch.full((F, B), -1, dtype=torch.int64, device="cuda")
        nuniq = torch.zeros(F, dtype=torch.int32, device="cuda")
        rg = torch.zeros(F, B, 16, device="cuda")
        sumsq = torch.zeros(F, device="cuda")
        if big:
            nb = _lib.query("nasrec_emb_grad_sort_reduce_big_ws_bytes", B, F)
            ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
            _lib.call("nasrec_emb_grad_sort_reduce_big", cat.data_ptr(), rows.data_ptr(), err.data_ptr(), gout.data_ptr(), B, F,
                      uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(), ws.data_ptr(), nb)
        else:
            scratch = torch.empty(F, B + 1, dtype=torch.int32, device="cuda")
            _lib.call("nasrec_emb_grad_sort_reduce_checked", cat.data_ptr(), rows.data_ptr(), err.data_ptr(), gout.data_ptr(), B, F,
                      uniq.data_ptr(), nuniq.data_ptr(), rg.data_ptr(), sumsq.data_ptr(), scratch.data_ptr())
        torch.cuda.synchronize()
        return uniq.cpu(), nuniq.cpu(), rg.cpu(), sumsq.cpu(), int(err.item())

    u1, n1, r1, s1, e1 = run(True)
    assert e1 == 1
    catn = cat_x.numpy().copy()
    sets = orc.embedding_row_sets(catn)
    for f in range(F):
        want = sets[f] if f != 3 else np.unique(np.delete(catn[:, 3], 7))
        assert u1[f, :n1[f]].tolist() == want.tolist(), f
    # reference sums in float64
    f = 2
    ids = catn[:, f]
    ref = np.zeros((int(n1[f]), 16))
    pos = {int(r): i for i, r in enumerate(u1[f, :n1[f]].tolist())}
    gh = gout[:, f].double().cpu().numpy()
    for b in range(B):
        ref[pos[int(ids[b])]] += gh[b]
    assert np.abs(r1[f, :n1[f]].numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
    if B <= 16384:
        u0, n0, r0, s0, e0 = run(False)
        assert e0 == 1 and torch.equal(n0, n1) and torch.equal(u0, u1)
        counts = np.bincount(ids, minlength=ne[f])
        short = [i for i, r in enumerate(u1[f, :n1[f]].tolist()) if counts[r] <= 256]
        assert torch.equal(r0[f, short], r1[f, short])
        # hot rows (thousands of duplicates): the two kernels add in different (both fixed) orders -> fp32 rounding only
        assert float((r0 - r1).abs().max()) <= 2e-6 * float(r1.abs().max()) * 16
        assert torch.allclose(s0, s1, rtol=1e-4)
