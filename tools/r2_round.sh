#!/bin/bash
python -m pytest tests -x -q -m gpu > gpurun_out/r2u_gputests.log 2>&1; tail -3 gpurun_out/r2u_gputests.log; grep "^E  " gpurun_out/r2u_gputests.log | head -4 | cut -c1-600
python bench.py --no-cpu --no-extras > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; cut -c1-250 gpurun_out/r2u_bench.json
