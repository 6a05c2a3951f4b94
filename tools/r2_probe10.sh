#!/bin/bash
exec > gpurun_out/r2_probe10.log 2>&1
NASREC_TC_BN=64 NASREC_TC_NS=1 python tools/step_dump.py /tmp/ref.npz 1 > /dev/null
for lo in 0 6 12 18 24 30 36 42 48; do hi=$((lo+6)); echo "== range $lo $hi"; NASREC_SPLIT_KINDS=1 NASREC_SPLIT_LO=$lo NASREC_SPLIT_HI=$hi NASREC_TC_BN=64 python tools/step_dump.py /tmp/a.npz 1 > /dev/null; python tools/step_cmp.py /tmp/ref.npz /tmp/a.npz 2>/dev/null | head -2; done
