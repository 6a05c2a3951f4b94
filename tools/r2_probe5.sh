#!/bin/bash
exec > gpurun_out/r2_probe5.log 2>&1
python -m pytest tests/test_gpu_gemm_tma.py -x -q -m gpu 2>&1 | tail -15
for k in 256 1024 4096; do python tools/gemm_prof2.py 512 1024 $k 3; done
for n in 64 256; do python tools/gemm_prof2.py 512 $n 1024 3; done
python tools/gemm_prof2.py 8192 1024 1024 3
for bn in 32 64 128; do for ns in 1 2 4 8; do NASREC_TC_BN=$bn NASREC_TC_NS=$ns python tools/gemm_prof2.py 512 1024 1024 3 | sed "s/^/ns=$ns /"; done; done
