#!/bin/bash
exec > gpurun_out/r2_probe6.log 2>&1
T="tests/test_gpu_baseline_sizes.py::test_small_supernet_training_step_B512_capped_tables"
run() { echo "== $*"; env "$@" python -m pytest $T -x -q -m gpu 2>&1 | grep -E "^E +assert [0-9]|passed|failed" | head -3; }
run NASREC_TILE_POLICY=1
run NASREC_TC_NS=1
run NASREC_TC_BN=32
run NASREC_TC_BN=64
run NASREC_TC_BN=64 NASREC_TC_NS=1
run NASREC_TC_BN=32 NASREC_TC_NS=1
run NASREC_GEMM_TMA=0
run NASREC_GEMM_MODE=4
