#!/bin/bash
exec > gpurun_out/r2_probe11.log 2>&1
NASREC_SPLIT_VERBOSE=1 NASREC_SPLIT_KINDS=1 NASREC_SPLIT_LO=0 NASREC_SPLIT_HI=8 NASREC_TC_BN=64 python tools/step_dump.py /tmp/a.npz 1 2>&1 >/dev/null | grep -A1 "split launch"
