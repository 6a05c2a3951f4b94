#!/bin/bash
python -m pytest tests -x -q -m gpu > gpurun_out/r2s_gputests.log 2>&1; tail -3 gpurun_out/r2s_gputests.log; grep "^E  " gpurun_out/r2s_gputests.log | head -4 | cut -c1-600
python bench.py --no-cpu > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; cut -c1-250 gpurun_out/r2s_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
for k,v in d["extra"]["hbm_kernels"].items():
    if isinstance(v,dict) and "sort_reduce" in k: print(k, round(v["us"],1))
PY
