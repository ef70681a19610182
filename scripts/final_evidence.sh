#!/bin/bash
# Round-2 evidence on one GPU: the driver's bench command, its ncu launch list, the full capture of the hot kernel and of
# the kernels around it. Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out /tmp/prof
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split('(')[0]
    ns = float(r[-1].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/r02_launches.txt', 'w') as f:
    f.write('ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-70s launches %4d  total %10.1f us  share %5.1f %%\n' % (k[:70], v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
bash scripts/profile_hot.sh r02_eval_final > /dev/null 2>&1
cuobjdump -sass -fun '_ZN3pqa13k_eval_stagedILi5ELi2ELi8EEEvNS_12StagedParamsE' probqa_b200/lib/libPqaCore.so > /tmp/prof/hot.sass 2>/dev/null
python scripts/sass_loops.py /tmp/prof/hot.sass > gpurun_out/r02_hot_loops.txt 2>/dev/null
bash scripts/profile_others.sh > /dev/null 2>&1
bash scripts/profile_multi.sh "8 32" > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
bash scripts/client_loop.sh > /dev/null 2>&1
bash scripts/sweep_mid_batches.sh > gpurun_out/r02_mid_batches.txt 2>&1
tail -3 gpurun_out/r02_launches.txt; tail -12 gpurun_out/r02_eval_final.txt; tail -12 gpurun_out/r02_mid_batches.txt
