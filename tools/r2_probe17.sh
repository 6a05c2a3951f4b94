#!/bin/bash
exec > gpurun_out/r2_probe17.log 2>&1
export NASREC_SPLIT_KINDS=1 NASREC_SPLIT_LO=1 NASREC_SPLIT_HI=2 WITH_ORACLE=1
echo "== launch 1 split, LDG kernel + workspace"; NASREC_FORCE_NO_TMA=1 NASREC_TC_BN=64 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -3
echo "== launch 1 split, cluster bn=32 ns=2"; NASREC_TC_BN=32 NASREC_TC_NS=2 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -3
echo "== launch 1 split, cluster bn=64 ns=2"; NASREC_TC_BN=64 NASREC_TC_NS=2 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -3
echo "== nothing split but LO/HI set, bn=64"; NASREC_SPLIT_LO=100000 NASREC_SPLIT_HI=100001 NASREC_TC_BN=64 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -3
