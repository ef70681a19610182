#!/bin/bash
# ncu captures of this round's kernels (run under gpurun, one GPU). Text summaries land in gpurun_out/ (the .ncu-rep
# files are summarised on the box with scripts/ncu_summary.py and only the hot kernel's report is kept: 64 MiB limit).
mkdir -p gpurun_out /tmp/prof
# (1) launch list of the default bench command (times are cold-cache and serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
# (2) full capture of the hot kernel at config 2 (B = 256)
ncu --set full --clock-control none --import-source on -k regex:k_eval_staged -s 4 -c 1 -o /tmp/prof/eval_staged -f \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_staged.log 2>&1
# (3) full capture of the chunked 4-warp variant (B = 64, T = 20000)
ncu --set full --clock-control none -k regex:k_eval_staged -s 2 -c 1 -o /tmp/prof/eval_staged_chunked_b64 -f \
  python bench.py --workload 2000x5x20000_b64 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_chunked.log 2>&1
# (4) the two phases of the target-sharded evaluation, two shard engines on this one GPU
ncu --set full --clock-control none -k regex:k_eval_tshard -s 2 -c 2 -o /tmp/prof/eval_tshard -f \
  python scripts/tshard_one_gpu.py > gpurun_out/ncu_tshard.log 2>&1
for R in eval_staged eval_staged_chunked_b64 eval_tshard; do
  python scripts/ncu_summary.py /tmp/prof/$R.ncu-rep > gpurun_out/$R.txt 2>&1
done
cp /tmp/prof/eval_staged.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/ /tmp/prof
