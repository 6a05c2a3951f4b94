import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nasrec_b200 import _lib
from oracle import nasrec_oracle as orc
from tests.helpers import load_golden
from tests.test_gpu_supernet import _build, _run_case
meta, arr = load_golden("supernet_autoctr_criteo")
cfg = meta["cfg"]; case = meta["cases"][0]
int_x, cat_x, y = orc.synth_batch(meta["batch"], meta["nd"], meta["num_embeddings"], seed=case["batch_seed"])
_lib.LIB.set_gemm_mode(3)
m, _ = _build(cfg, meta["num_embeddings"], meta["nd"], meta["shapes"], meta["state_seed"])
out = _run_case(m, cfg, case["choice"], int_x, cat_x, y)
print("done", float(out[1]))
