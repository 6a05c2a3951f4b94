#!/bin/bash
# Builds libnasrec_b200.so (sm_100a) in-tree.  Used by __graft_entry__.build().
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
pids=()
for f in gemm ln emb emb_big interact optim attn fused_ops metrics net; do
  if [ ! -f "$HERE/obj/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/obj/$f.o" ] || [ "$HERE/common.cuh" -nt "$HERE/obj/$f.o" ] || [ "$HERE/gemm_tc.cuh" -nt "$HERE/obj/$f.o" ] || [ "$HERE/gemm_common.cuh" -nt "$HERE/obj/$f.o" ] || [ "$HERE/gemm_tma.cuh" -nt "$HERE/obj/$f.o" ] || [ "$HERE/../../include/nasrec_b200.h" -nt "$HERE/obj/$f.o" ]; then
    $NVCC $FLAGS -c "$HERE/$f.cu" -o "$HERE/obj/$f.o" > "$HERE/obj/$f.log" 2>&1 &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p || { cat "$HERE"/obj/*.log; exit 1; }; done
$NVCC -shared -o "$OUT/libnasrec_b200.so" "$HERE"/obj/{gemm,ln,emb,emb_big,interact,optim,attn,fused_ops,metrics,net}.o -lcudart
echo "built $OUT/libnasrec_b200.so"
