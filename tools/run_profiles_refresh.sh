#!/bin/bash
# partial refresh: launch list + weight-gradient GEMMs + backward elementwise kernels (after the split-K / pool-map changes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_train.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:wgrad_halo_kernel|wgrad_gemm_kernel" -s 21 -c 21 -o gpurun_out/prof_wgrad_train $CMD > gpurun_out/ncu_wgrad_train.log 2>&1
echo "ncu wgrad exit $?"
python tools/ncu_summary.py report gpurun_out/prof_wgrad_train.ncu-rep gpurun_out/ncu_wgrad_train.csv
python tools/ncu_hot.py gpurun_out/prof_wgrad_train.ncu-rep 0 60 > gpurun_out/hot_wgrad_train_l0.txt 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:bn_bwd|pool_bwd|head_ce_fused|split_input|first_conv|view_colsum" -s 0 -c 14 -o gpurun_out/prof_elem2_train $CMD > gpurun_out/ncu_elem2_train.log 2>&1
echo "ncu elem2 exit $?"
python tools/ncu_summary.py report gpurun_out/prof_elem2_train.ncu-rep gpurun_out/ncu_elem2_train.csv
rm -f gpurun_out/*.ncu-rep
