#!/bin/bash
exec > gpurun_out/r2_probe4.log 2>&1
python -m pytest tests/test_gpu_gemm_tma.py -x -q -m gpu 2>&1 | tail -5
for k in 32 256 1024 4096; do python tools/gemm_prof2.py 512 1024 $k 3; done
for d in 15 6 4 2 1; do NASREC_GEMM_DBG=$d python tools/gemm_prof2.py 512 1024 4096 3 | grep fwd; done
python tools/gemm_prof2.py 8192 1024 1024 3
NASREC_TC_BN=128 python tools/gemm_prof2.py 8192 1024 1024 3
