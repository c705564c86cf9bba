#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call46.log
: > $L
timeout -k 10 900 python -m pytest tests -m gpu -q -x >> $L 2>&1; echo "rc=$?" >> $L
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
grep -E "passed|failed|FAILED|Error|smoke ok|^rc=" $L | tail -6
