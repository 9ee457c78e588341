#!/bin/bash
# same-box A/B of two builds of the library: tools/ab_lib.sh tools/ab_libs/lib_base.so tools/ab_libs/lib_new.so
cd "$(dirname "$0")/.."
DST=crimac-classifiers-unet_b200/libcrimac_b200.so
cp $DST /tmp/lib_keep.so
for v in "$1" "$2" "$1" "$2"; do
  cp "$v" $DST
  timeout 300 python bench.py --steps 40 --warmup 5 --quick --no-cpu-baseline --profile-out gpurun_out/ab_bd.csv 2> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train $v', round(d['ms_per_step'],4), round(d['value'],1), d['clocks']['sm_mhz'])"
  sed -n 3,6p gpurun_out/ab_bd.csv | cut -d, -f1-5 | cut -c1-60,100-
  grep "^\"bn_apply\|^\"bn_relu" gpurun_out/ab_bd.csv
  timeout 300 python bench.py --mode infer --steps 40 --warmup 5 --no-cpu-baseline 2>> gpurun_out/ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('infer $v', round(d['ms_per_step'],4), round(d['value'],1))"
done
cp /tmp/lib_keep.so $DST
