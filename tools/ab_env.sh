#!/bin/bash
# same-box A/B of one environment switch: tools/ab_env.sh VAR VALUE_A VALUE_B   (value "-" = unset)
cd "$(dirname "$0")/.."
VAR=$1; A=$2; B=$3
for v in "$A" "$B" "$A" "$B"; do
  if [ "$v" == "-" ]; then unset $VAR; else export $VAR="$v"; fi
  timeout 300 python bench.py --steps 40 --warmup 5 --quick --no-cpu-baseline 2> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train $VAR=$v', round(d['ms_per_step'],4), round(d['value'],1), d['clocks']['sm_mhz'])"
  timeout 300 python bench.py --mode infer --steps 40 --warmup 5 --no-cpu-baseline 2>> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('infer $VAR=$v', round(d['ms_per_step'],4), round(d['value'],1))"
done
