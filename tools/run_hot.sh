#!/bin/bash
# warp-stall hot spots of one kernel of the inference bench: tools/run_hot.sh <kernel regex> <launch skip> <out name>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_hot.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -c 1 -s $2 -o gpurun_out/prof_hot $CMD > gpurun_out/ncu_hot.log 2>&1
echo "ncu exit $?"
python tools/ncu_hot.py gpurun_out/prof_hot.ncu-rep 0 70 > gpurun_out/$3 2>&1
python tools/ncu_summary.py report gpurun_out/prof_hot.ncu-rep gpurun_out/$3.csv
rm -f gpurun_out/prof_hot.ncu-rep
