#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
sed -n '/^if \[ "\$1" == "ncu" \]/,$p' tools/run_round.sh > /tmp/ncu_part.sh
bash /tmp/ncu_part.sh ncu 2>&1 | tail -20
ls -la gpurun_out/*.ncu-rep 2>/dev/null
