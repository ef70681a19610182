mkdir -p gpurun_out
OUT=gpurun_out/r02_variants3.log; : > $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-28s value %.4g  kernel_ms %.4f  e2e %.4g' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))
" "$1"; }
$B | summ base >> $OUT
PQA_B200_LIB=probqa_b200/lib/exp/nocheck/libPqaCore.so $B 2>&1 | summ nocheck >> $OUT
$B --workload 1000x5x1000_b128 | summ b128 >> $OUT
python bench.py --workload 10000x5x10000_b1024 --steps 3 --warmup 3 --no-cpu-baseline | summ config3 >> $OUT
cat $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_sharded.py tests/test_gpu_big_configs.py -x -q -m gpu 2>&1 > gpurun_out/r02_tests_v8.log; tail -3 gpurun_out/r02_tests_v8.log
