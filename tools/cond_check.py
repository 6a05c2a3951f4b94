"""GPU diagnostic: is a gradient mismatch vs the fp32 reference a bug or conditioning?
Compares grads of GEMM modes 0 (FFMA) and 3 (3xTF32) and the fp32 oracle against an fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import SuperNet, ops_config_lib, _lib
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden
from tests.test_gpu_supernet import _build, _run_case

name = sys.argv[1] if len(sys.argv) > 1 else "supernet_autoctr_criteo"
meta, arr = load_golden(name)
cfg = meta["cfg"]
for ci, case in enumerate(meta["cases"]):
    int_x, cat_x, y = orc.synth_batch(meta["batch"], meta["nd"], meta["num_embeddings"], seed=case["batch_seed"])
    sd = orc.fill_state_dict(meta["shapes"], meta["state_seed"])
    sd64 = {k: v.double() for k, v in sd.items()}
    torch.set_default_dtype(torch.float64)
    l64, _, g64 = orc.loss_and_grads(sd64, cfg, case["choice"], int_x.double(), cat_x, y.double())
    torch.set_default_dtype(torch.float32)
    _, _, g32 = orc.loss_and_grads(sd, cfg, case["choice"], int_x, cat_x, y)
    res = {}
    for mode in (0, 3, 3):
        _lib.LIB.set_gemm_mode(mode)
        m, _ = _build(cfg, meta["num_embeddings"], meta["nd"], meta["shapes"], meta["state_seed"])
        logits, loss, grads = _run_case(m, cfg, case["choice"], int_x, cat_x, y)
        worst = ("", 0.0)
        for n, g in g64.items():
            gn = float(g.norm())
            if gn == 0 or n not in grads: continue
            e = float((grads[n].double() - g).norm() / gn)
            if e > worst[1]: worst = (n, e)
        lerr = float((logits.double() - l64).abs().max() / l64.abs().max())
        print("case %d mode %d: logits err vs fp64 %.2e  worst grad rel-L2 err vs fp64 %.2e (%s)" % (ci, mode, lerr, worst[1], worst[0]))
    worst = ("", 0.0)
    for n, g in g64.items():
        gn = float(g.norm())
        if gn == 0: continue
        e = float((g32[n].double() - g).norm() / gn)
        if e > worst[1]: worst = (n, e)
    print("case %d fp32 CPU oracle: worst grad rel-L2 err vs fp64 %.2e (%s)" % (ci, worst[1], worst[0]))
