#!/bin/bash
# closing GPU run of round 2: the whole GPU suite on the final tree (with the four-threads-per-token attention kernels),
# the attention kernels old vs new, and the fixed-model bench line
timeout 140 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_gputests.log 2>&1; tail -2 gpurun_out/r2y_gputests.log; grep "^E  " gpurun_out/r2y_gputests.log | head -5 | cut -c1-300
NASREC_ATTN_OLD=1 timeout 25 python tools/attn_prof.py > gpurun_out/r2y_attn_old.log 2>&1; tail -7 gpurun_out/r2y_attn_old.log | cut -c1-200
timeout 25 python tools/attn_prof.py > gpurun_out/r2y_attn_new.log 2>&1; tail -7 gpurun_out/r2y_attn_new.log | cut -c1-200
timeout 40 python bench.py --config criteo_full_best --no-extras --no-cpu > gpurun_out/r2y_bench_cfb.json 2> gpurun_out/r2y_bench_cfb.err; cut -c1-300 gpurun_out/r2y_bench_cfb.json
NASREC_ATTN_FWD4_MAXB=100000 timeout 25 python tools/attn_prof.py > gpurun_out/r2y_attn_new_all.log 2>&1; tail -3 gpurun_out/r2y_attn_new_all.log | cut -c1-200
