#!/bin/bash
# N-GPU check of the data-parallel path: tools/run_dp.sh N [extra]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_dp.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_dp.txt 2>&1
if [ "$2" != "stress" ] && [ "$2" != "benchonly" ]; then
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check_$N.log 2>&1
echo "dp_check exit $?"; tail -25 gpurun_out/dp_check_$N.log
CRIMAC_AR_MULTICAST=0 DP_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/dp_check_${N}_p2p.log 2>&1
echo "dp_check (no multicast) exit $?"; tail -8 gpurun_out/dp_check_${N}_p2p.log
fi
if [ "$2" == "bench" ] || [ "$2" == "benchonly" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
  echo "bench dp$N exit $?"; cat gpurun_out/bench_dp$N.json; tail -5 gpurun_out/bench_dp$N.err
fi
if [ "$2" == "stress" ]; then
  # BASELINE configs[4]: 6 frequencies, 512x512 patches, batch 64 per GPU, train + infer
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --in-ch 6 --size 512 --batch 64 --steps 8 --warmup 3 --quick --no-cpu-baseline > gpurun_out/bench_stress_train_dp$N.json 2> gpurun_out/bench_stress_train_dp$N.err
  echo "stress train dp$N exit $?"; cat gpurun_out/bench_stress_train_dp$N.json; tail -3 gpurun_out/bench_stress_train_dp$N.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --mode infer --in-ch 6 --size 512 --batch 64 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stress_infer_dp$N.json 2> gpurun_out/bench_stress_infer_dp$N.err
  echo "stress infer dp$N exit $?"; cat gpurun_out/bench_stress_infer_dp$N.json; tail -3 gpurun_out/bench_stress_infer_dp$N.err
fi
