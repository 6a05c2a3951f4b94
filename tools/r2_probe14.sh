#!/bin/bash
exec > gpurun_out/r2_probe14.log 2>&1
echo "== default plan, no side stream"; NO_OVERLAP=1 WITH_ORACLE=1 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -4
echo "== mode 0, no side stream"; NO_OVERLAP=1 WITH_ORACLE=1 NASREC_GEMM_MODE=0 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -4
echo "== default plan, with side stream, CUDA_LAUNCH_BLOCKING"; CUDA_LAUNCH_BLOCKING=1 WITH_ORACLE=1 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -4
