#!/bin/bash
python -m pytest tests -x -q -m gpu > gpurun_out/r2v_gputests.log 2>&1; tail -3 gpurun_out/r2v_gputests.log; grep "^E  " gpurun_out/r2v_gputests.log | head -6 | cut -c1-500
python bench.py --config criteo_full_best --no-cpu --no-extras 2>/dev/null | cut -c1-260
