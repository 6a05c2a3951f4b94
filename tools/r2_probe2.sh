#!/bin/bash
exec > gpurun_out/r2_probe2.log 2>&1
for d in 0 1 2 4 8 6 7 14 15; do echo "== DBG $d"; NASREC_GEMM_DBG=$d python tools/gemm_prof.py 512 3 1024 | grep "tma=1"; done
for bn in 64 128; do echo "== BN $bn"; NASREC_TC_BN=$bn python tools/gemm_prof.py 512 3 1024 | grep "tma=1"; done
echo "== policy 0"; NASREC_TILE_POLICY=0 python tools/gemm_prof.py 512 3 1024 | grep "tma=1"
for d in 0 1 2 4 8 15; do echo "== M8192 DBG $d"; NASREC_GEMM_DBG=$d python tools/gemm_prof.py 8192 3 1024 | grep "tma=1"; done
