#!/bin/bash
# Final profile set of a round: launch list + ncu --set full captures of one train step's tensor-core kernels and of
# the main HBM-bound kernels, plus one inference step.  Run under gpurun; summaries are made on the CPU box with
# tools/ncu_summary.py and committed under profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_train.log 2>&1 || { echo "plain train run failed"; tail -5 gpurun_out/plain_train.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 42 -c 42 -o gpurun_out/prof_conv_train $CMD > gpurun_out/ncu_conv_train.log 2>&1
echo "ncu conv train exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:wgrad_halo_kernel|wgrad_gemm_kernel" -s 21 -c 21 -o gpurun_out/prof_wgrad_train $CMD > gpurun_out/ncu_wgrad_train.log 2>&1
echo "ncu wgrad exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:first_conv|bn_bwd|bn_apply|pool_bwd|head_ce_fused|pack_conv|wgrad_unpack_all|sgd_kernel" -s 0 -c 16 -o gpurun_out/prof_elem_train $CMD > gpurun_out/ncu_elem_train.log 2>&1
echo "ncu elem exit $?"
CMDI="python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline"
$CMDI > gpurun_out/plain_infer.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:conv_igemm_kernel|first_conv" -s 22 -c 22 -o gpurun_out/prof_conv_infer $CMDI > gpurun_out/ncu_conv_infer.log 2>&1
echo "ncu conv infer exit $?"
# summaries are produced here (the raw reports are too big to travel back: gpurun_out/ is capped at 64 MiB)
for n in conv_train wgrad_train elem_train conv_infer; do
  python tools/ncu_summary.py report gpurun_out/prof_$n.ncu-rep gpurun_out/ncu_$n.csv
done
python tools/ncu_hot.py gpurun_out/prof_conv_train.ncu-rep 1 60 > gpurun_out/hot_conv_train_l0.txt 2>&1
python tools/ncu_hot.py gpurun_out/prof_wgrad_train.ncu-rep 0 60 > gpurun_out/hot_wgrad_train_l0.txt 2>&1
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/*.ncu-rep
