#!/bin/bash
python -m pytest tests/test_gpu_gemm_tma.py -x -q -m gpu 2>&1 | tail -3
for v in "NASREC_SMALL_K=0" "NASREC_SMALL_K=64"; do echo "== $v"; env $v python tools/gemm_prof2.py 512 1024 32 3; env $v python tools/gemm_prof2.py 512 16 32 3 | head -2; done
for v in "NASREC_SMALL_K=0" "NASREC_SMALL_K=64"; do echo; echo "== $v"; env $v python bench.py --no-cpu --no-extras 2>/dev/null | cut -c1-300; done
