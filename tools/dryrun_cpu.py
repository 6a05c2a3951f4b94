"""Developer aid (no GPU here): exercise all host-side control flow of the CUDA path
with the kernel launches stubbed out.  Arithmetic results are meaningless; this only
shakes out Python-level errors before spending GPU time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import _lib
import nasrec_b200.engine as eng
import nasrec_b200.supernet.modules as mods
import nasrec_b200.supernet.supernet as sn
import nasrec_b200.utils.train_utils as tu

calls = {}
def fake_call(name, *args):
    assert name in _lib._SIGS, name
    assert len(args) + 1 == len(_lib._SIGS[name][0]), (name, len(args) + 1, len(_lib._SIGS[name][0]))
    calls[name] = calls.get(name, 0) + 1
_lib.call = fake_call; eng.call = fake_call
_lib.LIB.workspace = torch.empty(1)
import contextlib
_lib.pin_stream = contextlib.nullcontext
# bypass CUDA checks
class _T(torch.Tensor): pass
import builtins
orig_is_cuda = torch.Tensor.is_cuda
torch.Tensor.is_cuda = property(lambda self: True)

from oracle import nasrec_oracle as orc
from tests.helpers import load_golden

def run_model(cfg, ne, nd, shapes, choice, B=5, fused=False):
    fixed = cfg["fixed"]
    m = sn.SuperNet(num_blocks=7, ops_config=sn.ops_config_lib[cfg["ops"]], use_layernorm=cfg["use_layernorm"],
                    num_embeddings=ne, sparse_input_size=len(ne), path_sampling_strategy="fixed-path" if fixed else "full-path",
                    fixed=fixed, fixed_choice=choice if fixed else None)
    m.materialize(nd)
    if not fixed:
        m.configure_choice(choice); m.configure_path_sampling_strategy("fixed-path")
    int_x, cat_x, y = orc.synth_batch(B, nd, ne, seed=1)
    if fused:
        tr = tu.FusedTrainer(m, lr=0.1)
        tr.step(int_x, cat_x, y)
    else:
        out = m(int_x, cat_x)
        out.sum().backward()
        n = sum(p.grad is not None for p in m.parameters())
        return n

for name in ["supernet_autoctr_criteo", "supernet_xlarge_criteo", "supernet_xlarge_kdd"]:
    meta, _ = load_golden(name)
    for case in meta["cases"]:
        n = run_model(meta["cfg"], meta["num_embeddings"], meta["nd"], meta["shapes"], case["choice"])
        assert n >= len(case["grad_norms"]), (name, n, len(case["grad_norms"]))
        run_model(meta["cfg"], meta["num_embeddings"], meta["nd"], meta["shapes"], case["choice"], fused=True)
    print(name, "ok")
meta, _ = load_golden("fixed_best")
for tag, mm in meta["models"].items():
    n = run_model(mm["cfg"], mm["num_embeddings"], mm["nd"], mm["shapes"], mm["choice"])
    assert n >= len(mm["grad_norms"]), (tag, n, len(mm["grad_norms"]))
    run_model(mm["cfg"], mm["num_embeddings"], mm["nd"], mm["shapes"], mm["choice"], fused=True)
    print(tag, "ok")
smeta, _ = load_golden("samplers")
meta, _ = load_golden("supernet_xlarge_criteo")
for ch in smeta["ea_candidates"]["xlarge"]:
    run_model(meta["cfg"], meta["num_embeddings"], 13, meta["shapes"], ch)
import nasrec_b200.search as srch
G, _ = load_golden("ea_finetune")
m = sn.SuperNet(num_blocks=7, ops_config=sn.ops_config_lib["xlarge"], use_layernorm=True,
                num_embeddings=G["num_embeddings"], sparse_input_size=26, path_sampling_strategy="full-path")
m.materialize(13)
tr = [orc.synth_batch(16, 13, G["num_embeddings"], seed=500 + b) for b in range(5)]
ev = [orc.synth_batch(32, 13, G["num_embeddings"], seed=900 + b) for b in range(2)]
srch.binary_metrics_device = lambda z, y: (0.5, 0.5, 0.7)
se = srch.SubnetEvaluator(m, use_native=False)
for ts in (8192, 40, 1):
    r = se.finetune_and_score(G["cands"][0]["choice"], tr, ev, lr=0.04, trunk_samples=ts)
S = srch.Searcher(se, srch.Tokenizer(7, sn.ops_config_lib["xlarge"]), tr, ev)
np.random.seed(3)
S.regularized_evolution_from_supernet(n_generations=2, n_childs=2, init_population=4, sample_size=2, top_k=1)
print("ea ok")
# module-level standalone
for ln, fixed in ((True, False), (False, True)):
    for cls, kw, ins, d in (
        (mods.ElasticLinear, dict(max_dims_or_dims=128, activation="relu"), [torch.randn(4, 50)], 32),
        (mods.ElasticLinear3D, dict(max_dims_or_dims=64, activation="relu", embedding_dim=16), [torch.randn(4, 30, 16)], 16),
        (mods.DotProduct, dict(max_dims_or_dims=128, embedding_dim=16), [torch.randn(4, 50), torch.randn(4, 30, 16)], 32),
        (mods.Sum, dict(max_dims_or_dims=128, activation="relu"), [torch.randn(4, 50), torch.randn(4, 20)], 32),
        (mods.SigmoidGating, dict(max_dims_or_dims=128, activation="relu"), [torch.randn(4, 50), torch.randn(4, 20)], 32),
        (mods.Transformer, dict(max_dims_or_dims=64, activation="relu", embedding_dim=16), [torch.randn(4, 30, 16)], 16),
        (mods.FactorizationMachine3D, dict(max_dims_or_dims=128), [torch.randn(4, 30, 16)], 32)):
        mod = cls(fixed=fixed, use_layernorm=ln, **kw)
        xs = [x.clone().requires_grad_(True) for x in ins]
        dd = kw["max_dims_or_dims"] if fixed else d
        out = mod(*xs, dd)
        out.sum().backward()
        assert all(x.grad is not None for x in xs), cls
        assert all(p.grad is not None for p in mod.parameters()), (cls, [n for n, p in mod.named_parameters() if p.grad is None])
print("modules ok")
print(sorted(calls.items()))
