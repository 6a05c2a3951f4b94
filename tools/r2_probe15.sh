#!/bin/bash
exec > gpurun_out/r2_probe15.log 2>&1
ORACLE_SPLIT=1 WITH_ORACLE=1 python tools/step_dump.py /tmp/a.npz 1 | grep "oracle"
