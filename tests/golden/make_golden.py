"""
Generates tests/golden/*.json|*.npz by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on deterministic
weights and inputs.  The committed fixtures are what pins oracle/nasrec_oracle.py
and the CUDA path; nothing at test time reads /root/reference.

    python tests/golden/make_golden.py            # regenerate everything

Weights come from oracle.fill_state_dict(shapes, seed) (copied INTO the
reference model), inputs from oracle.synth_batch(seed) -- both regenerate
bit-identically on any box, so the fixtures only hold shapes, choices and the
reference's outputs (logits, loss, per-tensor grad norms, a few small grads).
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

# shims: fvcore is imported at module scope by train_utils
fv = types.ModuleType("fvcore"); fvnn = types.ModuleType("fvcore.nn"); fvnn.FlopCountAnalysis = object
fv.nn = fvnn; sys.modules["fvcore"] = fv; sys.modules["fvcore.nn"] = fvnn
np.int = int

from nasrec.supernet.supernet import SuperNet, ops_config_lib          # noqa: E402  (reference)
from nasrec.searcher.tokenizer import Tokenizer                          # noqa: E402  (reference)
from oracle import nasrec_oracle as orc                                   # noqa: E402

DATASETS = {
    "criteo": dict(nd=13, F=26, ne=[1461, 584, 10131227, 2202609, 306, 25, 12518, 634, 4, 93146, 5684, 8351593,
                                    3195, 28, 14993, 5461307, 11, 5653, 2174, 5, 7046548, 19, 16, 286182, 106, 142573]),
    "avazu": dict(nd=1, F=23, ne=[10000, 241, 8, 8, 4738, 7746, 27, 8553, 560, 37, 2686409, 6729487, 8252, 6, 5,
                                  2627, 9, 10, 436, 5, 69, 173, 61]),
    "kdd": dict(nd=3, F=10, ne=[26274, 641708, 14848, 22122011, 1188090, 3735797, 2934102, 20004011, 4, 8]),
}
ROW_CAP = 40   # tiny tables: the fixtures pin arithmetic, not capacity


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return [jsonable(v) for v in o.tolist()]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    return o


def build(ds, ops, ln, fixed=False, choice=None, strategy="full-path", anypath="uniform", steps=0):
    D = DATASETS[ds]
    ne = [min(n, ROW_CAP) for n in D["ne"]]
    m = SuperNet(num_blocks=7, ops_config=ops_config_lib[ops], use_layernorm=ln, num_embeddings=ne,
                 sparse_input_size=D["F"], path_sampling_strategy="fixed-path" if fixed else strategy,
                 fixed=fixed, fixed_choice=choice, anypath_choice=anypath, supernet_training_steps=steps)
    int_x, cat_x, _ = orc.synth_batch(2, D["nd"], ne, seed=7, all_zero_dense=(ds == "avazu"))
    with torch.no_grad():
        m(int_x, cat_x)                       # lazy warm-up (train_utils.py:413-433)
    return m, ne


def load_filled(m, seed):
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = orc.fill_state_dict(shapes, seed)
    m.load_state_dict(sd, strict=True)
    return shapes


def run_case(m, ne, ds, choice, B, seed, fixed):
    D = DATASETS[ds]
    int_x, cat_x, y = orc.synth_batch(B, D["nd"], ne, seed=seed, all_zero_dense=(ds == "avazu"))
    if not fixed:
        m.configure_choice(choice)
        m.configure_path_sampling_strategy("fixed-path")
    m.zero_grad()
    logits = m(int_x, cat_x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    gn = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
    small = {n: p.grad.numpy().copy() for n, p in m.named_parameters()
             if p.grad is not None and p.grad.numel() <= 4096 and not n.startswith("_embedding")}
    emb_rows = {str(f): np.nonzero(np.abs(m._embedding[f].weight.grad.numpy()).sum(1))[0].tolist()
                for f in range(D["F"])}
    return dict(logits=logits.detach().numpy().copy(), loss=float(loss), grad_norms=gn, small_grads=small,
                emb_rows=emb_rows)


def sample_choices(ds, ops, strategy, anypath, seed, n, exhausted=True, steps=15000):
    """Reference samplers (supernet.py:432-511, 1009-1061) -> recorded choices."""
    m, ne = build(ds, ops, True, strategy="full-path", anypath=anypath, steps=steps)
    m.configure_path_sampling_strategy(strategy)
    if exhausted:
        m._supernet_train_steps_counter = steps + 5
        for b in m._blocks:
            b._supernet_train_steps_counter = steps + 5
    D = DATASETS[ds]
    int_x, cat_x, _ = orc.synth_batch(2, D["nd"], ne, seed=3, all_zero_dense=(ds == "avazu"))
    np.random.seed(seed)
    out = []
    with torch.no_grad():
        for _ in range(n):
            m(int_x, cat_x)
            out.append(jsonable(m.choice))
    return out


def save(name, meta, arrays):
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, "arrays:", len(arrays))


def supernet_fixture(name, ds, ops, nchoices, B=5):
    m, ne = build(ds, ops, True)
    shapes = load_filled(m, seed=11)
    cfg = dict(ops=ops, use_layernorm=True, fixed=False, num_blocks=7)
    choices = [orc.full_path_choice(cfg)]
    choices += sample_choices(ds, ops, "default", "binomial-0.5", seed=0, n=nchoices)
    choices += sample_choices(ds, ops, "any-path", "uniform", seed=1, n=nchoices)
    meta = dict(dataset=ds, cfg=cfg, num_embeddings=ne, nd=DATASETS[ds]["nd"], state_seed=11, batch=B,
                shapes={k: list(v) for k, v in shapes.items()}, cases=[])
    arrays = {}
    for ci, ch in enumerate(choices):
        r = run_case(m, ne, ds, ch, B, seed=100 + ci, fixed=False)
        meta["cases"].append(dict(choice=ch, batch_seed=100 + ci, loss=r["loss"], grad_norms=r["grad_norms"],
                                  emb_rows=r["emb_rows"]))
        arrays["logits_%d" % ci] = r["logits"]
        for n_, g in r["small_grads"].items():
            if n_.startswith("_final") or "._nodes.4._mha" in n_ or "_ln" in n_ or "layernorm" in n_:
                arrays["grad_%d/%s" % (ci, n_)] = g
    save(name, meta, arrays)


def fixed_fixture():
    cfgdir = "/root/reference/nasrec/configs"
    files = {
        "criteo_xlarge": ("criteo", "criteo/ea_criteo_kaggle_xlarge_best_1shot.json"),
        "criteo_autoctr": ("criteo", "criteo/ea_criteo_kaggle_autoctr_best_1shot.json"),
        "avazu_xlarge": ("avazu", "avazu/ea_avazu_kaggle_xlarge_best_1shot.json"),
        "avazu_autoctr": ("avazu", "avazu/ea_avazu_kaggle_autoctr_best_1shot.json"),
        "kdd_xlarge": ("kdd", "kdd/ea_kdd_kaggle_xlarge_best_1shot.json"),
        "kdd_autoctr": ("kdd", "kdd/ea_kdd_kaggle_autoctr_best_1shot.json"),
    }
    meta = dict(models={})
    arrays = {}
    for key, (ds, rel) in files.items():
        choice = json.load(open(os.path.join(cfgdir, rel)))
        for ln in ([False, True] if key == "criteo_xlarge" else [False]):   # main_train.py:262 forces False
            tag = key + ("_ln" if ln else "")
            m, ne = build(ds, choice["config"], ln, fixed=True, choice=choice)
            shapes = load_filled(m, seed=5)
            r = run_case(m, ne, ds, choice, 6, seed=42, fixed=True)
            meta["models"][tag] = dict(dataset=ds, choice=jsonable(choice), num_embeddings=ne, nd=DATASETS[ds]["nd"],
                                       cfg=dict(ops=choice["config"], use_layernorm=ln, fixed=True, num_blocks=7),
                                       state_seed=5, batch=6, batch_seed=42, loss=r["loss"],
                                       shapes={k: list(v) for k, v in shapes.items()},
                                       grad_norms=r["grad_norms"], emb_rows=r["emb_rows"],
                                       dense_params=int(sum(p.numel() for n, p in m.named_parameters()
                                                            if not n.startswith("_embedding"))))
            arrays["logits/" + tag] = r["logits"]
            for n_, g in r["small_grads"].items():
                if n_.startswith("_final") or "_mha" in n_:
                    arrays["grad/%s/%s" % (tag, n_)] = g
    save("fixed_best", meta, arrays)


def step_fixture():
    """3 reference training steps (train_utils.py:262-286) on an autoctr and an
    xlarge supernet: Adagrad(lr, eps=1e-2) + clip 5.0, dense embedding grads."""
    meta = dict(runs={})
    arrays = {}
    for tag, ops, lr in (("autoctr", "autoctr", 0.12), ("xlarge", "xlarge", 0.12)):
        m, ne = build("criteo", ops, True)
        shapes = load_filled(m, seed=21)
        choices = sample_choices("criteo", ops, "default", "binomial-0.5", seed=4, n=3)
        opt = torch.optim.Adagrad(m.parameters(), lr=lr, eps=1e-2)
        m.configure_path_sampling_strategy("fixed-path")
        losses, norms = [], []
        for si, ch in enumerate(choices):
            int_x, cat_x, y = orc.synth_batch(8, 13, ne, seed=300 + si)
            m.configure_choice(ch)
            opt.zero_grad()
            logits = m(int_x, cat_x)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
            loss.backward()
            tn = torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
            opt.step()
            losses.append(float(loss)); norms.append(float(tn))
            arrays["%s/logits_%d" % (tag, si)] = logits.detach().numpy().copy()
        sd = m.state_dict()
        checks = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
        arrays[tag + "/final_weight"] = sd["_final.weight"].numpy().copy()
        arrays[tag + "/emb0"] = sd["_embedding.0.weight"].numpy().copy()
        meta["runs"][tag] = dict(cfg=dict(ops=ops, use_layernorm=True, fixed=False, num_blocks=7), lr=lr,
                                 num_embeddings=ne, state_seed=21, choices=choices, losses=losses,
                                 total_norms=norms, checksums=checks,
                                 shapes={k: list(v) for k, v in shapes.items()})
    save("train_steps", meta, arrays)


def bf16_fixture():
    """bf16 compute (BASELINE config #4; the reference's --use_amp hooks, train_utils.py:146,247-286): the reference under
    torch.autocast("cpu", dtype=bfloat16) -- linear layers in bf16 with fp32 accumulation and bf16 outputs, LayerNorm /
    softmax / loss in fp32, fp32 master weights -- forward + backward on the Avazu and Criteo NASRec-Full best models and
    on an autoctr supernet choice.  Pins the stated bf16 tolerance of GEMM mode 2."""
    cfgdir = "/root/reference/nasrec/configs"
    meta = dict(models={})
    arrays = {}
    cases = [("avazu_xlarge", "avazu", "avazu/ea_avazu_kaggle_xlarge_best_1shot.json", False),
             ("criteo_xlarge", "criteo", "criteo/ea_criteo_kaggle_xlarge_best_1shot.json", False)]
    for tag, ds, rel, ln in cases:
        choice = json.load(open(os.path.join(cfgdir, rel)))
        m, ne = build(ds, choice["config"], ln, fixed=True, choice=choice)
        shapes = load_filled(m, seed=5)
        D = DATASETS[ds]
        int_x, cat_x, y = orc.synth_batch(64, D["nd"], ne, seed=42, all_zero_dense=(ds == "avazu"))
        m.zero_grad()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            logits = m(int_x, cat_x)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.float(), y)
        loss.backward()
        gn = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        m.zero_grad()
        ref32 = m(int_x, cat_x)
        meta["models"][tag] = dict(dataset=ds, choice=jsonable(choice), num_embeddings=ne, nd=D["nd"],
                                   cfg=dict(ops=choice["config"], use_layernorm=ln, fixed=True, num_blocks=7), state_seed=5,
                                   batch=64, batch_seed=42, loss=float(loss), grad_norms=gn,
                                   shapes={k: list(v) for k, v in shapes.items()},
                                   fp32_vs_bf16_logit_rms=float((ref32.detach() - logits.detach().float()).pow(2).mean().sqrt()),
                                   logit_rms=float(ref32.detach().pow(2).mean().sqrt()))
        arrays["logits/" + tag] = logits.detach().float().numpy().copy()
        arrays["logits_fp32/" + tag] = ref32.detach().numpy().copy()
    save("bf16_autocast", meta, arrays)


def lr_fixture():
    """lr sequences of the reference schedulers (utils/lr_schedule.py) as the EA fine-tune
    and main_train.py drive them: construct, step(epoch=-1), then step() per batch."""
    from nasrec.utils.lr_schedule import CosineAnnealingWarmupRestarts, ConstantWithWarmup
    out = {}

    def opt(lr):
        return torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)

    for tag, kw, n in (("cosine_20_w2", dict(first_cycle_steps=20, warmup_steps=2, max_lr=0.04, min_lr=1e-8), 45),
                       ("cosine_mult2_gamma", dict(first_cycle_steps=10, warmup_steps=3, max_lr=0.1, min_lr=0.001,
                                                   cycle_mult=2.0, gamma=0.5), 40)):
        o = opt(0.04)
        s = CosineAnnealingWarmupRestarts(o, **kw)
        seq = [o.param_groups[0]["lr"]]
        s.step(epoch=-1); seq.append(o.param_groups[0]["lr"])
        for _ in range(n):
            s.step(); seq.append(o.param_groups[0]["lr"])
        seeks = []
        for e in (0, 1, 5, 19, 20, 33, 71):
            s.step(epoch=e); seeks.append([e, o.param_groups[0]["lr"], s.cycle, s.step_in_cycle])
        out[tag] = dict(kwargs=kw, opt_lr=0.04, seq=seq, seeks=seeks)
    o = opt(0.12)
    s = ConstantWithWarmup(o, num_warmup_steps=5)
    seq = [o.param_groups[0]["lr"]]
    for _ in range(12):
        o.step(); s.step(); seq.append(o.param_groups[0]["lr"])
    out["constant_w5"] = dict(opt_lr=0.12, seq=seq)
    with open(os.path.join(HERE, "lr_schedules.json"), "w") as f:
        json.dump(out, f)
    print("wrote lr_schedules")


def finetune_fixture():
    """The EA's one-shot scoring recipe on the reference (eval_subnet_from_supernet.py:118-200,
    train_utils.py:262-300,129-180): last-layer-only fine-tune with Adagrad + cosine warm-up
    schedule + clip 5.0 on a pinned candidate, then log-loss on held-out batches."""
    from nasrec.utils.lr_schedule import CosineAnnealingWarmupRestarts
    meta = dict(cands=[])
    arrays = {}
    m, ne = build("criteo", "xlarge", True)
    shapes = load_filled(m, seed=33)
    base = {k: v.clone() for k, v in m.state_dict().items()}
    tok = Tokenizer(7, ops_config_lib["xlarge"])
    np.random.seed(77)
    cands = [jsonable(tok.generate_random_choice()) for _ in range(2)]
    steps, lr = 12, 0.04
    for ci, ch in enumerate(cands):
        m.load_state_dict(base, strict=True)
        m.configure_choice(ch)
        m.configure_path_sampling_strategy("fixed-path")
        m.set_mode_to_finelune_last_only()
        opt = torch.optim.Adagrad(m.parameters(), lr=lr, eps=1e-2)
        sch = CosineAnnealingWarmupRestarts(opt, first_cycle_steps=steps, warmup_steps=steps // 10, max_lr=lr,
                                            min_lr=1e-8)
        sch.step(epoch=-1)
        losses, lrs = [], []
        m.train()
        for b in range(steps):
            int_x, cat_x, y = orc.synth_batch(16, 13, ne, seed=500 + b)
            opt.zero_grad()
            loss = torch.nn.functional.binary_cross_entropy_with_logits(m(int_x, cat_x), y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            losses.append(float(loss))
            if b < steps - 1:
                sch.step()
        m.eval()
        outs, ys = [], []
        with torch.no_grad():
            for b in range(3):
                int_x, cat_x, y = orc.synth_batch(32, 13, ne, seed=900 + b)
                outs.append(m(int_x, cat_x)); ys.append(y)
        z = torch.cat(outs); yy = torch.cat(ys)
        test_loss = float(torch.nn.functional.binary_cross_entropy_with_logits(z, yy))
        arrays["cand%d/final_weight" % ci] = m._final.weight.detach().numpy().copy()
        arrays["cand%d/final_bias" % ci] = m._final.bias.detach().numpy().copy()
        arrays["cand%d/eval_logits" % ci] = z.numpy().copy()
        meta["cands"].append(dict(choice=ch, losses=losses, lrs=lrs, test_loss=test_loss))
    meta.update(cfg=dict(ops="xlarge", use_layernorm=True, fixed=False, num_blocks=7), num_embeddings=ne,
                state_seed=33, steps=steps, lr=lr, train_seeds=[500, 16], eval_seeds=[900, 32, 3],
                shapes={k: list(v) for k, v in shapes.items()})
    save("ea_finetune", meta, arrays)


def tokenizer_fixture():
    """Tokenizer goldens (searcher/tokenizer.py:158-265): token/hash of random candidates and a
    chain of mutate_spec draws from numpy's global RNG."""
    out = {}
    for ops in ("xlarge", "autoctr"):
        tok = Tokenizer(7, ops_config_lib[ops])
        np.random.seed(4321)
        cands = [tok.generate_random_choice() for _ in range(3)]
        rec = dict(cands=jsonable(cands), tokens=[tok.tokenize(c).tolist() for c in cands],
                   hashes=[tok.hash_token(tok.tokenize(c)) for c in cands])
        np.random.seed(99)
        chain, cur = [], cands[0]
        for _ in range(16):
            cur = tok.mutate_spec(cur)
            chain.append(dict(choice=jsonable(cur), hash=tok.hash_token(tok.tokenize(cur))))
        rec["mutations_seed99"] = chain
        out[ops] = rec
    with open(os.path.join(HERE, "tokenizer.json"), "w") as f:
        json.dump(out, f)
    print("wrote tokenizer")


def transform_fixture():
    """Raw batch -> (int_x, cat_x, y) through the reference's VanillaTransform* (data_pipes.py:135-252)
    on synthetic raw rows: negative / zero / large ints, hex ids of 1-8 digits, upper case, and
    missing fields (empty string -> row 0)."""
    from nasrec.utils import data_pipes as dp
    from nasrec.utils.config import NUM_EMBEDDINGS_CRITEO, NUM_EMBEDDINGS_AVAZU, NUM_EMBEDDINGS_KDD
    out = {}
    rng = np.random.RandomState(2024)
    for ds, fn, nd, F, ne in (("criteo", dp.VanillaTransformCriteo, 13, 26, NUM_EMBEDDINGS_CRITEO),
                              ("avazu", dp.VanillaTransformAvazu, 1, 23, NUM_EMBEDDINGS_AVAZU),
                              ("kdd", dp.VanillaTransformKDD, 3, 10, NUM_EMBEDDINGS_KDD)):
        B = 48
        ints = rng.randint(-3, 60, size=(B, nd)).astype(np.int64)
        ints[rng.rand(B, nd) < 0.1] = rng.randint(1 << 20, 1 << 23)
        hexs = []
        for f in range(F):
            col = []
            for b in range(B):
                u = rng.rand()
                if u < 0.12:
                    col.append("")
                else:
                    digits = int(rng.randint(1, 9))
                    v = "%x" % int(rng.randint(0, 16 ** digits, dtype=np.int64))
                    col.append(v.upper() if u > 0.9 else v)
            hexs.append(col)
        label = rng.randint(0, 2, size=B).astype(np.int64)
        batch = {"label": torch.tensor(label)}
        for c in range(nd):
            batch["int_%d" % c] = torch.tensor(ints[:, c])
        for f in range(F):
            batch["cat_%d" % f] = list(hexs[f])
        int_x, cat_x, y = fn(batch)
        out[ds] = dict(nd=nd, F=F, num_embeddings=list(ne), ints=ints.tolist(), hex=hexs, label=label.tolist(),
                       int_x=int_x.numpy().astype(np.float32).tolist(), cat_x=cat_x.numpy().tolist(),
                       y=y.numpy().tolist(), int_dtype=str(int_x.dtype), cat_dtype=str(cat_x.dtype))
    with open(os.path.join(HERE, "input_transform.json"), "w") as f:
        json.dump(out, f)
    print("wrote input_transform")


def sampler_fixture():
    """RNG-order goldens: what the reference draws from numpy's
    global legacy RNG, forward by forward."""
    meta = dict(streams=[])
    for ops in ("xlarge", "autoctr", "xlarge-zeros"):
        for strategy, anypath in (("default", "binomial-0.5"), ("default", "uniform"), ("single-path", "uniform"),
                                  ("any-path", "binomial-0.5"), ("any-path", "uniform")):
            for exhausted in (True, False):
                for seed in (0, 5):
                    ch = sample_choices("criteo", ops, strategy, anypath, seed, n=4, exhausted=exhausted, steps=6)
                    meta["streams"].append(dict(ops=ops, strategy=strategy, anypath=anypath, exhausted=exhausted,
                                                seed=seed, steps=6, choices=ch))
    # fixed-path sampled once, then frozen (supernet.py:476-491, 1035-1048)
    m, ne = build("criteo", "xlarge", True)
    m.configure_path_sampling_strategy("fixed-path")
    int_x, cat_x, _ = orc.synth_batch(2, 13, ne, seed=3)
    np.random.seed(9)
    with torch.no_grad():
        m(int_x, cat_x); c1 = jsonable(m.choice); m(int_x, cat_x); c2 = jsonable(m.choice)
    assert c1 == c2
    meta["fixed_path_seed9"] = c1
    # EA candidate generator (tokenizer.py:267-336) -- defines config #3's candidates
    meta["ea_candidates"] = {}
    for ops in ("xlarge", "autoctr"):
        tok = Tokenizer(7, ops_config_lib[ops])
        np.random.seed(1234)
        meta["ea_candidates"][ops] = [jsonable(tok.generate_random_choice()) for _ in range(6)]
    with open(os.path.join(HERE, "samplers.json"), "w") as f:
        json.dump(meta, f)
    print("wrote samplers")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["samplers", "fixed", "autoctr", "xlarge", "kdd", "steps", "lr", "finetune", "tokenizer", "transform", "avazu",
                             "zeros", "bf16"]
    if "samplers" in which:
        sampler_fixture()
    if "fixed" in which:
        fixed_fixture()
    if "autoctr" in which:
        supernet_fixture("supernet_autoctr_criteo", "criteo", "autoctr", nchoices=3)
    if "xlarge" in which:
        supernet_fixture("supernet_xlarge_criteo", "criteo", "xlarge", nchoices=3)
    if "kdd" in which:
        supernet_fixture("supernet_xlarge_kdd", "kdd", "xlarge", nchoices=2)
    if "avazu" in which:
        supernet_fixture("supernet_xlarge_avazu", "avazu", "xlarge", nchoices=1)
    if "bf16" in which:
        bf16_fixture()
    if "zeros" in which:        # the xlarge-zeros search space (supernet.py:151-168): Zeros2D / Zeros3D nodes
        supernet_fixture("supernet_zeros_criteo", "criteo", "xlarge-zeros", nchoices=4)
    if "steps" in which:
        step_fixture()
    if "lr" in which:
        lr_fixture()
    if "finetune" in which:
        finetune_fixture()
    if "tokenizer" in which:
        tokenizer_fixture()
    if "transform" in which:
        transform_fixture()
