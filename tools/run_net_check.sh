#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/net_check.log
: > $LOG
for t in ${@:-infer train2 train5 time}; do
  timeout 300 python tools/gpu_net_check.py $t >> $LOG 2>&1
  echo "--- exit $? ($t)" >> $LOG
done
cat $LOG
