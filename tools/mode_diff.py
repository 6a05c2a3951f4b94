"""GPU diagnostic: per-tensor gradient difference between GEMM mode 0 (FFMA) and mode 3 (3xTF32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nasrec_b200 import _lib
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden
from tests.test_gpu_supernet import _build, _run_case
name = sys.argv[1] if len(sys.argv) > 1 else "supernet_autoctr_criteo"
ci = int(sys.argv[2]) if len(sys.argv) > 2 else 0
meta, arr = load_golden(name)
cfg = meta["cfg"]; case = meta["cases"][ci]
int_x, cat_x, y = orc.synth_batch(meta["batch"], meta["nd"], meta["num_embeddings"], seed=case["batch_seed"])
out = {}
for mode in (0, 3):
    _lib.LIB.set_gemm_mode(mode)
    m, _ = _build(cfg, meta["num_embeddings"], meta["nd"], meta["shapes"], meta["state_seed"])
    out[mode] = _run_case(m, cfg, case["choice"], int_x, cat_x, y)
print("logits diff", float((out[0][0] - out[3][0]).abs().max()))
rows = []
for n, g in out[0][2].items():
    g3 = out[3][2][n]
    gn = float(g.double().norm())
    if gn == 0: continue
    rows.append((float((g.double() - g3.double()).norm() / gn), n, tuple(g.shape)))
for e, n, sh in sorted(rows, reverse=True)[:40]:
    print("%.2e  %-55s %s" % (e, n, sh))
print("---- smallest")
for e, n, sh in sorted(rows)[:25]:
    print("%.2e  %-55s %s" % (e, n, sh))
