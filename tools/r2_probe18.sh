#!/bin/bash
exec > gpurun_out/r2_probe18.log 2>&1
ORACLE_FLIP=3 WITH_ORACLE=1 python tools/step_dump.py /tmp/a.npz 1 | grep -v "^loss"
