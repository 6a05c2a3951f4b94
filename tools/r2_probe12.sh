#!/bin/bash
exec > gpurun_out/r2_probe12.log 2>&1
NASREC_TC_BN=64 NASREC_TC_NS=2 python tools/gemm_sweep.py 2>&1 | head -4
NASREC_TC_BN=64 NASREC_TC_NS=1 python tools/gemm_sweep.py 2>&1 | head -2
