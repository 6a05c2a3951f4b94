#!/bin/bash
python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gputests.log 2>&1; tail -3 gpurun_out/r2g_gputests.log
python bench.py --no-cpu --no-extras > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; cut -c1-330 gpurun_out/r2g_bench.json
for v in "NASREC_SPROJ_WGRAD_OLD=1" "NASREC_TILE_POLICY=1"; do echo; echo "== $v"; env $v python bench.py --no-cpu --no-extras 2>/dev/null | cut -c1-330; done
