"""GPU: whole-network parity of the CUDA path (through SuperNet.forward -> C ABI)
against (1) the committed outputs of the unmodified reference (tests/golden) and
(2) the oracle run live on the same seeded inputs, incl. ragged / edge batches.

Bar (BASELINE.json north_star): logits and log-loss within 1e-5 relative (fp32);
bit-exact embedding row sets; gradients within 2e-4 relative per tensor."""
import numpy as np
import pytest
import torch

from nasrec_b200 import SuperNet, ops_config_lib
from nasrec_b200.utils.train_utils import FusedTrainer, reference_style_step
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden, rel_err, relu_kink_margin

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-5
GRAD_TOL = 5e-4          # per-tensor gradient norms, well-conditioned cases
KINK_GRAD_TOL = 5e-2     # cases where a ReLU pre-activation lies within fp32 rounding of zero
KINK_MARGIN = 6e-6       # ~ the fp32 rounding error of a pre-activation; see tests/helpers.relu_kink_margin


def _build(cfg, ne, nd, shapes, seed, choice=None):
    fixed = cfg["fixed"]
    m = SuperNet(num_blocks=cfg["num_blocks"], ops_config=ops_config_lib[cfg["ops"]],
                 use_layernorm=cfg["use_layernorm"], num_embeddings=ne, sparse_input_size=len(ne),
                 path_sampling_strategy="fixed-path" if fixed else "full-path", fixed=fixed,
                 fixed_choice=choice if fixed else None)
    m = m.to("cuda")
    m.materialize(nd)
    sd = orc.fill_state_dict(shapes, seed)
    m.load_state_dict(sd, strict=True)
    return m, sd


def _run_case(m, cfg, choice, int_x, cat_x, y):
    if not cfg["fixed"]:
        m.configure_choice(choice)
        m.configure_path_sampling_strategy("fixed-path")
    m.zero_grad(set_to_none=True)
    logits = m(int_x.cuda(), cat_x.cuda())
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    return logits.detach().cpu(), float(loss), grads


def _compare(logits, loss, grads, logits_ref, loss_ref, gn_ref, small, emb_rows, margin=1.0):
    assert rel_err(logits.numpy(), logits_ref) < LOGIT_TOL
    assert abs(loss - loss_ref) < LOGIT_TOL * max(1.0, abs(loss_ref))
    gtol = GRAD_TOL if margin >= KINK_MARGIN else KINK_GRAD_TOL
    for n, g in gn_ref.items():
        assert n in grads, "missing grad for " + n
        got = float(grads[n].double().norm())
        assert abs(got - g) <= gtol * max(g, 1e-3) + 1e-7, (n, got, g, margin)
    for n, g in grads.items():
        if n not in gn_ref:
            assert float(g.abs().max()) == 0.0, "unexpected grad for " + n
    for n, g in small.items():     # full small tensors: relative L2 distance
        d = float(np.linalg.norm(grads[n].numpy().astype(np.float64) - g)) / max(float(np.linalg.norm(g)), 1e-12)
        assert d < 2 * gtol, (n, d, margin)
    for f, rows in emb_rows.items():
        got = np.nonzero(np.abs(grads["_embedding.%s.weight" % f].numpy()).sum(1))[0].tolist()
        assert got == rows, "embedding row set of table %s" % f


@pytest.mark.parametrize("name", ["supernet_autoctr_criteo", "supernet_xlarge_criteo", "supernet_xlarge_kdd",
                                  "supernet_xlarge_avazu", "supernet_zeros_criteo"])
def test_supernet_matches_reference_golden(name):
    meta, arr = load_golden(name)
    m, sd = _build(meta["cfg"], meta["num_embeddings"], meta["nd"], meta["shapes"], meta["state_seed"])
    strict = 0
    for ci, case in enumerate(meta["cases"]):
        int_x, cat_x, y = orc.synth_batch(meta["batch"], meta["nd"], meta["num_embeddings"], seed=case["batch_seed"],
                                          all_zero_dense=(meta["dataset"] == "avazu"))
        logits, loss, grads = _run_case(m, meta["cfg"], case["choice"], int_x, cat_x, y)
        small = {k.split("/", 1)[1]: v for k, v in arr.items() if k.startswith("grad_%d/" % ci)}
        margin = relu_kink_margin(sd, meta["cfg"], case["choice"], int_x, cat_x)
        strict += margin >= KINK_MARGIN
        _compare(logits, loss, grads, arr["logits_%d" % ci], case["loss"], case["grad_norms"], small,
                 case["emb_rows"], margin)
    assert strict >= len(meta["cases"]) // 3, "too few well-conditioned cases for a meaningful gradient check"


def test_fixed_best_models_match_reference_golden():
    meta, arr = load_golden("fixed_best")
    for tag, mm in meta["models"].items():
        m, sd = _build(mm["cfg"], mm["num_embeddings"], mm["nd"], mm["shapes"], mm["state_seed"], mm["choice"])
        int_x, cat_x, y = orc.synth_batch(mm["batch"], mm["nd"], mm["num_embeddings"], seed=mm["batch_seed"],
                                          all_zero_dense=(mm["dataset"] == "avazu"))
        logits, loss, grads = _run_case(m, mm["cfg"], mm["choice"], int_x, cat_x, y)
        small = {k.split("/", 2)[2]: v for k, v in arr.items() if k.startswith("grad/%s/" % tag)}
        margin = relu_kink_margin(sd, mm["cfg"], mm["choice"], int_x, cat_x)
        _compare(logits, loss, grads, arr["logits/" + tag], mm["loss"], mm["grad_norms"], small, mm["emb_rows"],
                 margin)


@pytest.mark.parametrize("B", [1, 3, 257, 1000])
def test_supernet_matches_oracle_live_ragged_batches(B):
    """Same seeded inputs through the oracle and the CUDA path at batch sizes the
    fixtures do not cover (1, odd, > one tile), EA candidate choices, Zipf ids."""
    meta, _ = load_golden("supernet_xlarge_criteo")
    smeta, _ = load_golden("samplers")
    cfg, ne, nd = meta["cfg"], meta["num_embeddings"], meta["nd"]
    m, sd = _build(cfg, ne, nd, meta["shapes"], 77)
    for ci, choice in enumerate(smeta["ea_candidates"]["xlarge"][:3]):
        int_x, cat_x, y = orc.synth_batch(B, nd, ne, seed=900 + ci, zipf=True)
        logits, loss, grads = _run_case(m, cfg, choice, int_x, cat_x, y)
        lr, lo, gr = orc.loss_and_grads(sd, cfg, choice, int_x, cat_x, y)
        assert rel_err(logits.numpy(), lr.numpy()) < LOGIT_TOL
        assert abs(loss - float(lo)) < LOGIT_TOL * max(1.0, abs(float(lo)))
        gtol = 1e-3 if relu_kink_margin(sd, cfg, choice, int_x, cat_x) >= KINK_MARGIN else KINK_GRAD_TOL
        for n, g in gr.items():
            gn = float(g.double().norm())
            if gn == 0.0:
                assert n not in grads or float(grads[n].abs().max()) == 0.0
                continue
            assert n in grads, n
            assert rel_err(grads[n].numpy(), g.numpy()) < gtol, n
        sets = orc.embedding_row_sets(cat_x.numpy())
        for f in range(len(ne)):
            got = np.nonzero(np.abs(grads["_embedding.%d.weight" % f].numpy()).sum(1))[0]
            assert set(got.tolist()) <= set(sets[f].tolist())


def test_ffma_mode_matches_reference_golden():
    """The fp32 CUDA-core GEMM mode (NASREC_GEMM_MODE=0) stays available and parity-green."""
    from nasrec_b200 import _lib
    prev = _lib.LIB.gemm_mode()
    _lib.LIB.set_gemm_mode(0)
    try:
        test_supernet_matches_reference_golden("supernet_autoctr_criteo")
        test_fixed_best_models_match_reference_golden()
    finally:
        _lib.LIB.set_gemm_mode(prev)


def test_default_gemm_mode_is_tensor_core():
    from nasrec_b200 import _lib
    import os
    if "NASREC_GEMM_MODE" not in os.environ:
        assert _lib.LIB.gemm_mode() == 3


def test_no_grad_and_frozen_modes():
    meta, arr = load_golden("supernet_autoctr_criteo")
    m, _ = _build(meta["cfg"], meta["num_embeddings"], meta["nd"], meta["shapes"], meta["state_seed"])
    case = meta["cases"][1]
    int_x, cat_x, y = orc.synth_batch(meta["batch"], meta["nd"], meta["num_embeddings"], seed=case["batch_seed"])
    m.configure_choice(case["choice"])
    m.configure_path_sampling_strategy("fixed-path")
    with torch.no_grad():
        out = m(int_x.cuda(), cat_x.cuda())
    assert not out.requires_grad
    assert rel_err(out.cpu().numpy(), arr["logits_1"]) < LOGIT_TOL
    m.set_mode_to_finelune_last_only()
    m.zero_grad(set_to_none=True)
    out = m(int_x.cuda(), cat_x.cuda())
    torch.nn.functional.binary_cross_entropy_with_logits(out, y.cuda()).backward()
    got = {n for n, p in m.named_parameters() if p.grad is not None}
    assert got == {"_final.weight", "_final.bias"}
    assert abs(float(m._final.weight.grad.double().norm()) - case["grad_norms"]["_final.weight"]) < 1e-4 * max(
        1.0, case["grad_norms"]["_final.weight"])


def test_training_steps_match_reference_both_optimizer_paths():
    """3 reference steps (Adagrad eps=1e-2, clip 5.0): stock torch optimizer on our model
    (drop-in path) and the fused sparse/dense optimizer (fast path) both reproduce the
    reference's logits, losses, clip norms and updated weights."""
    meta, arr = load_golden("train_steps")
    for tag, run in meta["runs"].items():
        for path in ("torch-optim", "fused"):
            m, _ = _build(run["cfg"], run["num_embeddings"], 13, run["shapes"], run["state_seed"])
            m.configure_path_sampling_strategy("fixed-path")
            if path == "torch-optim":
                opt = torch.optim.Adagrad(m.parameters(), lr=run["lr"], eps=1e-2)
            else:
                tr = FusedTrainer(m, lr=run["lr"], eps=1e-2, clip=5.0)
            for si, ch in enumerate(run["choices"]):
                int_x, cat_x, y = orc.synth_batch(8, 13, run["num_embeddings"], seed=300 + si)
                m.configure_choice(ch)
                if path == "torch-optim":
                    logits, loss = reference_style_step(m, opt, torch.nn.BCEWithLogitsLoss(), int_x.cuda(),
                                                        cat_x.cuda(), y.cuda(), 5.0)
                    loss = float(loss)
                else:
                    logits, loss = tr.step(int_x.cuda(), cat_x.cuda(), y.cuda())
                    loss = float(loss.item())
                    assert abs(float(tr.last_total_norm.item()) - run["total_norms"][si]) < 1e-3 * max(
                        1.0, run["total_norms"][si])
                # step 0 sees identical weights (1e-5 bar); later steps see weights after Adagrad updates,
                # whose g/(sqrt(g^2)+eps) form amplifies last-bit gradient differences of tiny gradients
                tol = LOGIT_TOL if si == 0 else 5e-4
                assert rel_err(logits.detach().cpu().numpy(), arr["%s/logits_%d" % (tag, si)]) < tol, (tag, path, si)
                assert abs(loss - run["losses"][si]) < 2e-4, (tag, path, si)
            sd = m.state_dict()
            assert rel_err(sd["_final.weight"].cpu().numpy(), arr[tag + "/final_weight"]) < 5e-4, (tag, path)
            assert rel_err(sd["_embedding.0.weight"].cpu().numpy(), arr[tag + "/emb0"]) < 5e-4, (tag, path)
            for k, (s1, s2) in run["checksums"].items():
                v = sd[k].double()
                assert abs(float(v.abs().sum()) - s2) <= 2e-4 * max(1.0, s2), (tag, path, k)


def test_block_standalone_api_matches_oracle():
    """SuperNetBlock.forward on the reference's zero-padded tensors."""
    from nasrec_b200.supernet.supernet import SuperNetBlock
    cfg = ops_config_lib["xlarge"]
    blk = SuperNetBlock(cfg, True, 1024, 64, 16, "relu", path_sampling_strategy="fixed-path").cuda()
    choice = {"active_nodes": [1, 4], "dense_in_dims": 256, "sparse_in_dims": 48, "dense_sparse_interact": 1,
              "deep_fm": 1}
    g = torch.Generator().manual_seed(3)
    dense, left, right = (torch.randn(9, 1037, generator=g) for _ in range(3))
    sparse = torch.randn(9, 98, 16, generator=g)
    blk.configure_choice(choice)
    do, so = blk((dense.cuda(), sparse.cuda(), left.cuda(), right.cuda()))
    shapes = {"_blocks.0." + k: list(v.shape) for k, v in blk.state_dict().items()}
    sd = orc.fill_state_dict(shapes, 5)
    blk.load_state_dict({k[len("_blocks.0."):]: v for k, v in sd.items()}, strict=True)
    do, so = blk((dense.cuda(), sparse.cuda(), left.cuda(), right.cuda()))
    rd, rs = orc.block_forward(sd, 0, orc.OPS_CONFIG["xlarge"], True, False, choice, dense, sparse, left, right)
    assert do.shape == rd.shape and so.shape == rs.shape
    assert rel_err(do.detach().cpu().numpy(), rd.numpy()) < 2e-5
    assert rel_err(so.detach().cpu().numpy(), rs.numpy()) < 2e-5


def test_cuda_graph_step_matches_eager_step():
    """Whole-step CUDA-graph replay (fixed best model) == the eager fused step, bit for bit."""
    from nasrec_b200.utils.graph import GraphedFusedTrainer
    meta, _ = load_golden("fixed_best")
    mm = meta["models"]["criteo_xlarge"]
    batches = [orc.synth_batch(64, mm["nd"], mm["num_embeddings"], seed=500 + i) for i in range(6)]
    finals = []
    for graphed in (False, True):
        m, _ = _build(mm["cfg"], mm["num_embeddings"], mm["nd"], mm["shapes"], mm["state_seed"], mm["choice"])
        tr = FusedTrainer(m, lr=0.16)
        if graphed:
            tr = GraphedFusedTrainer(tr)
        losses = []
        for b in batches:
            _, loss = tr.step(b[0].cuda(), b[1].cuda(), b[2].cuda())
            losses.append(float(loss.item()))
        finals.append((losses, {k: v.clone() for k, v in m.state_dict().items()}))
    assert finals[0][0] == finals[1][0]
    for k in finals[0][1]:
        assert torch.equal(finals[0][1][k], finals[1][1][k]), k


def test_subnet_evaluator_matches_oracle_and_graph_replay():
    """One-shot scoring: shared gather + per-candidate CUDA-graph replay == eager == oracle."""
    from nasrec_b200.search import SubnetEvaluator
    meta, _ = load_golden("supernet_xlarge_criteo")
    smeta, _ = load_golden("samplers")
    cfg, ne, nd = meta["cfg"], meta["num_embeddings"], meta["nd"]
    m, sd = _build(cfg, ne, nd, meta["shapes"], 31)
    m.requires_grad_(False)
    cands = smeta["ea_candidates"]["xlarge"][:3]
    host = [orc.synth_batch(96, nd, ne, seed=700 + i) for i in range(4)]
    batches = [tuple(t.cuda() for t in b) for b in host]
    ev = SubnetEvaluator(m)
    eager = ev.score(cands, batches, use_cuda_graph=False)
    graph = ev.score(cands, batches, use_cuda_graph=True)
    for ch, e, g in zip(cands, eager, graph):
        assert e == g
        logits = torch.cat([orc.supernet_forward(sd, cfg, ch, b[0], b[1]).reshape(-1) for b in host])
        ys = torch.cat([b[2].reshape(-1) for b in host])
        acc, auc, loss = orc.binary_metrics(logits.detach().numpy(), ys.numpy())
        assert abs(e["test_loss"] - loss) < 1e-5
        assert abs(e["test_auroc"] - auc) < 1e-4
        assert abs(e["test_acc"] - acc) < 1e-6
