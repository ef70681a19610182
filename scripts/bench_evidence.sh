#!/bin/bash
# the driver's two bench commands on one GPU, outputs under gpurun_out/
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_1gpu.err
tail -c 600 gpurun_out/r02_bench_1gpu.json
