#!/bin/bash
for n in 1 2; do
NASREC_LN_DEFER=0 python tools/step_dump.py /tmp/a$n.npz $n > /dev/null 2>&1
NASREC_LN_DEFER=1 python tools/step_dump.py /tmp/b$n.npz $n > /dev/null 2>&1
echo "== after $n step(s): defer off vs on"; python tools/step_cmp.py /tmp/a$n.npz /tmp/b$n.npz 2>/dev/null | head -8
done
