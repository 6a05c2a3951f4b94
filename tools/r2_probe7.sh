#!/bin/bash
exec > gpurun_out/r2_probe7.log 2>&1
echo "== BN64 NS1";  WITH_ORACLE=1 NASREC_TC_BN=64 NASREC_TC_NS=1 python tools/step_dump.py /tmp/a.npz 1 | grep -v choice
echo "== BN64 split"; WITH_ORACLE=1 NASREC_TC_BN=64 python tools/step_dump.py /tmp/b.npz 1 | grep -v choice
echo "== default plan"; WITH_ORACLE=1 python tools/step_dump.py /tmp/c.npz 1 | grep -v choice
echo "== policy 1"; WITH_ORACLE=1 NASREC_TILE_POLICY=1 python tools/step_dump.py /tmp/d.npz 1 | grep -v choice
