#!/bin/bash
# usage: tools/run_ncu_one.sh <kernel regex> <out name> [count] [extra bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="$1"; O="$2"; C="${3:-4}"; shift 3
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $@"
$CMD > gpurun_out/plain_$O.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base ${NCU_NAME_BASE:-function} -k "regex:$K" -s 0 -c $C -o gpurun_out/prof_$O $CMD > gpurun_out/ncu_$O.log 2>&1
echo "ncu $O exit $?"
