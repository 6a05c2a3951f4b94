#!/bin/bash
exec > gpurun_out/r2_probe13.log 2>&1
echo "== mode 0 (fp32 FFMA)"; WITH_ORACLE=1 NASREC_GEMM_MODE=0 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -6
echo "== mode 3 default plan"; WITH_ORACLE=1 python tools/step_dump.py /tmp/b.npz 1 | grep "vs oracle" | head -6
echo "== mode 4 default plan"; WITH_ORACLE=1 NASREC_GEMM_MODE=4 python tools/step_dump.py /tmp/b.npz 1 | grep "vs oracle" | head -6
echo "== mode 3 policy 1"; WITH_ORACLE=1 NASREC_TILE_POLICY=1 python tools/step_dump.py /tmp/c.npz 1 | grep "vs oracle" | head -4
python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -3
