#!/bin/bash
# final single-GPU evidence of the round: smoke, reference arm, default bench (all extras), ncu launch list
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
python bench.py --impl reference > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err; cut -c1-300 gpurun_out/r2z_bench_ref.json
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; cut -c1-300 gpurun_out/r2z_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 2800 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 6 --warmup 3 --no-extras --no-cpu > gpurun_out/r2z_b_ncu.log 2>&1
tail -c 200 gpurun_out/r2z_b_ncu.log
