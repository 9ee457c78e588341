#!/bin/bash
# GPU-box script of round 2: parity tests, headline bench (+ breakdown), optional extras.  Usage: tools/run_r2.sh [tests] [bench] [infer] [ncu]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -q -m gpu -s -x > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -60 gpurun_out/pytest_gpu.log ;;
tests_all)
  timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  grep -v "^$" gpurun_out/pytest_gpu.log | tail -120 ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/breakdown_train.csv > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
  echo "bench train exit $?"; cat gpurun_out/bench_train.json; tail -5 gpurun_out/bench_train.err
  head -16 gpurun_out/breakdown_train.csv ;;
quick)
  timeout 600 python bench.py --steps 20 --warmup 3 --quick --no-cpu-baseline --profile-out gpurun_out/breakdown_train.csv > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
  echo "bench quick exit $?"; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
  head -16 gpurun_out/breakdown_train.csv ;;
infer)
  timeout 300 python bench.py --mode infer --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/breakdown_infer.csv > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
  echo "bench infer exit $?"; cat gpurun_out/bench_infer.json; tail -5 gpurun_out/bench_infer.err
  head -12 gpurun_out/breakdown_infer.csv ;;
ref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench ref exit $?"; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err ;;
ncu)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick"
  $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 600 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list exit $?"
  $CMD > gpurun_out/plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 42 -s 126 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv exit $?"
  python tools/ncu_summary.py report gpurun_out/prof_conv.ncu-rep gpurun_out/ncu_conv_train.csv; rm -f gpurun_out/prof_conv.ncu-rep
  $CMD > gpurun_out/plain3.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:wgrad_halo|wgrad_gemm" -c 21 -s 63 -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_wgrad.log 2>&1
  echo "ncu wgrad exit $?"
  python tools/ncu_summary.py report gpurun_out/prof_wgrad.ncu-rep gpurun_out/ncu_wgrad_train.csv; rm -f gpurun_out/prof_wgrad.ncu-rep ;;
ncu_elem)
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick"
  $CMD > gpurun_out/plain4.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:bn_bwd|bn_apply|bn_finalize|head_ce|first_conv|partial_sum|view_colsum|unpack|pack_|sgd" -c 60 -s 330 -o gpurun_out/prof_elem $CMD > gpurun_out/ncu_elem.log 2>&1
  echo "ncu elem exit $?"
  python tools/ncu_summary.py report gpurun_out/prof_elem.ncu-rep gpurun_out/ncu_elem_train.csv; rm -f gpurun_out/prof_elem.ncu-rep
  $CMD > gpurun_out/plain6.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pool_kernel|head_ce_fused|sgd_kernel|upsample|wgrad_unpack" -c 16 -s 32 -o gpurun_out/prof_elem2 $CMD > gpurun_out/ncu_elem2.log 2>&1
  echo "ncu elem2 exit $?"
  python tools/ncu_summary.py report gpurun_out/prof_elem2.ncu-rep gpurun_out/ncu_elem2_train.csv; rm -f gpurun_out/prof_elem2.ncu-rep ;;
ncu_infer)
  CMD="python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain5.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_igemm|first_conv" -c 22 -s 66 -o gpurun_out/prof_infer $CMD > gpurun_out/ncu_infer.log 2>&1
  echo "ncu infer exit $?"
  python tools/ncu_summary.py report gpurun_out/prof_infer.ncu-rep gpurun_out/ncu_conv_infer.csv; rm -f gpurun_out/prof_infer.ncu-rep ;;
stress)
  timeout 600 python bench.py --mode train --in-ch 6 --size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --quick > gpurun_out/bench_stress_train.json 2> gpurun_out/bench_stress_train.err
  echo "bench stress train exit $?"; cat gpurun_out/bench_stress_train.json; tail -3 gpurun_out/bench_stress_train.err
  timeout 600 python bench.py --mode infer --in-ch 6 --size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress_infer.json 2> gpurun_out/bench_stress_infer.err
  echo "bench stress infer exit $?"; cat gpurun_out/bench_stress_infer.json; tail -3 gpurun_out/bench_stress_infer.err ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit $?"; tail -5 gpurun_out/smoke.log ;;
*) echo "unknown step $what" ;;
esac
done
