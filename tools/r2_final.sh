#!/bin/bash
# final single-GPU evidence of the round: tests, smoke, reference arm, default bench (all extras)
python -m pytest tests -x -q -m gpu > gpurun_out/r2z_gputests.log 2>&1; tail -2 gpurun_out/r2z_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; cut -c1-300 gpurun_out/r2z_bench.json
python bench.py --config criteo_full_best --no-extras > gpurun_out/r2z_bench_cfb.json 2> gpurun_out/r2z_bench_cfb.err; cut -c1-260 gpurun_out/r2z_bench_cfb.json
