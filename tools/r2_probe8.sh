#!/bin/bash
exec > gpurun_out/r2_probe8.log 2>&1
python tools/gemm_sweep.py
for bn in 64 128; do for ns in 2 4 8; do NASREC_TC_BN=$bn NASREC_TC_NS=$ns python tools/gemm_sweep.py; done; done
NASREC_TC_BN=32 NASREC_TC_NS=8 python tools/gemm_sweep.py
