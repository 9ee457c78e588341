#!/bin/bash
# ncu --set full of the HBM-bound backward kernels, the first-conv kernels and the fused head/CE kernel (one each level)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_train.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:bn_bwd|pool_bwd|head_ce_fused|split_input|first_conv|view_colsum|wgrad_unpack_all|sgd_kernel" -s 0 -c 24 -o gpurun_out/prof_elem2_train $CMD > gpurun_out/ncu_elem2_train.log 2>&1
echo "ncu elem2 exit $?"
python tools/ncu_summary.py report gpurun_out/prof_elem2_train.ncu-rep gpurun_out/ncu_elem2_train.csv
rm -f gpurun_out/*.ncu-rep
