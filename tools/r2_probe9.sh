#!/bin/bash
exec > gpurun_out/r2_probe9.log 2>&1
for k in 1 2 4; do echo "== kinds $k"; WITH_ORACLE=1 NASREC_SPLIT_KINDS=$k NASREC_TC_BN=64 python tools/step_dump.py /tmp/a.npz 1 | grep "vs oracle" | head -3; done
