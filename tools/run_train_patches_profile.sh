#!/bin/bash
# GPU box: time the training-sample kernels, then one ncu --set full capture of both (summarised on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/profile_train_patches.py 20 > gpurun_out/tp_time.log 2>&1 || { tail -5 gpurun_out/tp_time.log; exit 1; }
cat gpurun_out/tp_time.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "regex:train_" -s 8 -c 4 \
  -o gpurun_out/prof_tp python tools/profile_train_patches.py 3 > gpurun_out/ncu_tp.log 2>&1
echo "ncu exit $?"
python tools/ncu_summary.py report gpurun_out/prof_tp.ncu-rep gpurun_out/tp_ncu.csv
python tools/ncu_hot.py gpurun_out/prof_tp.ncu-rep 0 25 > gpurun_out/tp_hot_gather.txt 2>&1
rm -f gpurun_out/prof_tp.ncu-rep
cat gpurun_out/tp_ncu.csv
