#!/bin/bash
# ncu --set full capture (with source) of the hot evaluation kernel at BASELINE config 2; summary + report -> gpurun_out/
TAG=${1:-r02_eval}
mkdir -p gpurun_out /tmp/prof
ncu --set full --clock-control none --import-source on -k regex:k_eval_staged -s 4 -c 1 -o /tmp/prof/$TAG -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/prof/$TAG.ncu-rep > gpurun_out/$TAG.txt 2>&1
cp /tmp/prof/$TAG.ncu-rep gpurun_out/ 2>/dev/null
tail -40 gpurun_out/$TAG.txt
