#!/bin/bash
# one GPU visit: full -m gpu suite, default bench (with the per-launch GEMM trace)
python -m pytest tests -x -q -m gpu > gpurun_out/r2e_gputests.log 2>&1; tail -3 gpurun_out/r2e_gputests.log
rm -f gpurun_out/r2e_trace.txt
NASREC_GEMM_TRACE=gpurun_out/r2e_trace.txt python bench.py --no-cpu --no-extras > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; cut -c1-400 gpurun_out/r2e_bench.json
