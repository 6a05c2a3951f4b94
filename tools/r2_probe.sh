#!/bin/bash
# GPU probe: TMA delivery / MMA issue micro-benchmarks + GEMM shape timings (results -> gpurun_out/r2_probe.log)
exec > gpurun_out/r2_probe.log 2>&1
set -x
./build/tma_bench
./build/tma_bench s 4
./build/tma_bench s 2
./build/tma_bench m 1 t 1
./build/tma_bench m 1 t 1 s 2
./build/tma_bench m 1 t 1 s 1
./build/tma_bench m 37 t 4
./build/tma_bench m 4 t 37
./build/tma_bench t 8 b 128 n 2 s 4
./build/tma_bench t 8 b 128 n 2 s 4 m 16
./build/tma_bench t 8 b 128 n 1 s 4 m 16
./build/tma_bench a 128 b 32 n 2 w 16
./build/tma_bench m 1 t 1 a 128 b 128 n 2 s 4
for N in 32 64 128; do ./build/mma_bench N $N t 1; ./build/mma_bench N $N t 0; done
python tools/gemm_prof.py 512 3 1024
python tools/gemm_prof.py 8192 3 1024
python tools/gemm_prof.py 512 2 1024
