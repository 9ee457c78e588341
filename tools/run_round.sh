#!/bin/bash
# GPU-box round script: parity tests, headline bench (train) + inference bench, per-kernel breakdowns.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/breakdown_train.csv > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
echo "bench train exit $?"; cat gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
cat gpurun_out/breakdown_train.csv
timeout 300 python bench.py --mode infer --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/breakdown_infer.csv > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "bench infer exit $?"; cat gpurun_out/bench_infer.json; tail -5 gpurun_out/bench_infer.err
cat gpurun_out/breakdown_infer.csv
if [ "$1" == "ncu" ]; then
  # launch list of one short bench run (serialised, cold-cache: compare SHARES); then full captures of the top kernels
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 600 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  $CMD > gpurun_out/plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -c 11 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv exit $?"
  $CMD > gpurun_out/plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:wgrad_halo_kernel|wgrad_gemm_kernel" -c 6 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
  echo "ncu wgrad exit $?"
  $CMD > gpurun_out/plain4.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:first_conv|bn_bwd|head_bwd|bn_apply|pool_bwd" -c 12 -o gpurun_out/prof_elem $CMD > gpurun_out/ncu_elem.log 2>&1
  echo "ncu elem exit $?"
fi
if [ "$1" == "extra" ]; then
  timeout 600 python bench.py --mode survey --steps 10 --warmup 3 --batch 93 > gpurun_out/bench_survey.json 2> gpurun_out/bench_survey.err
  echo "bench survey exit $?"; cat gpurun_out/bench_survey.json; tail -5 gpurun_out/bench_survey.err
  timeout 600 python bench.py --mode train --in-ch 6 --size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress_train.json 2> gpurun_out/bench_stress_train.err
  echo "bench stress train exit $?"; cat gpurun_out/bench_stress_train.json; tail -5 gpurun_out/bench_stress_train.err
  timeout 600 python bench.py --mode infer --in-ch 6 --size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress_infer.json 2> gpurun_out/bench_stress_infer.err
  echo "bench stress infer exit $?"; cat gpurun_out/bench_stress_infer.json; tail -5 gpurun_out/bench_stress_infer.err
fi
