#!/bin/bash
exec > gpurun_out/r2_probe16.log 2>&1
export SWEEP_FIRST=1 NASREC_TC_BN=64 NASREC_TC_NS=2
timeout 600 compute-sanitizer --tool memcheck python tools/gemm_sweep.py 2>&1 | tail -15
timeout 600 compute-sanitizer --tool racecheck python tools/gemm_sweep.py 2>&1 | tail -25
