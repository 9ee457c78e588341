#!/bin/bash
# N-GPU check of the data-parallel path: tools/run_dp.sh N [extra]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_dp.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_dp.txt 2>&1
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/dp_check_$N.log 2>&1
echo "dp_check exit $?"; tail -25 gpurun_out/dp_check_$N.log
CRIMAC_AR_MULTICAST=0 DP_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/dp_check_${N}_p2p.log 2>&1
echo "dp_check (no multicast) exit $?"; tail -8 gpurun_out/dp_check_${N}_p2p.log
if [ "$2" == "bench" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
  echo "bench dp$N exit $?"; cat gpurun_out/bench_dp$N.json; tail -5 gpurun_out/bench_dp$N.err
fi
