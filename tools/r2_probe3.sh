#!/bin/bash
exec > gpurun_out/r2_probe3.log 2>&1
for k in 32 256 1024 4096; do for d in 0 15 7 6 2 4 1; do NASREC_GEMM_DBG=$d python tools/gemm_prof2.py 512 1024 $k 3 | grep fwd; done; done
for n in 256 64; do for k in 256 1024; do for d in 0 15; do NASREC_GEMM_DBG=$d python tools/gemm_prof2.py 512 $n $k 3 ; done; done; done
for bn in 64 128; do NASREC_TC_BN=$bn python tools/gemm_prof2.py 512 1024 1024 3; done
