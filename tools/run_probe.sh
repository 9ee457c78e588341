#!/bin/bash
# Runs each hardware probe in its own process (a trapped kernel poisons the CUDA context) with a hard timeout.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for t in ${@:-tma umma_kmajor umma_mnmajor umma_unaligned conv head convt dgrad wgrad}; do
  timeout 180 python tools/gpu_probe.py $t >> $LOG 2>&1
  echo "--- exit $? ($t)" >> $LOG
done
cat $LOG
