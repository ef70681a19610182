mkdir -p gpurun_out
OUT=gpurun_out/r02_variants.log; : > $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-28s value %.4g  kernel_ms %.4f  e2e %.4g' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))
" "$1"; }
$B | summ base >> $OUT
$B --lanes 1 | summ lanes1 >> $OUT
$B --lanes 4 | summ lanes4 >> $OUT
for v in nocheck tree nopass1 nopass2 nocheck_nopass1; do
  PQA_B200_LIB=probqa_b200/lib/exp/$v/libPqaCore.so $B 2>&1 | summ $v >> $OUT
done
$B --workload 1000x5x1000_b128 | summ b128 >> $OUT
$B --workload 1000x5x1000_b64 | summ b64 >> $OUT
$B --workload 1000x5x1000_b32 | summ b32 >> $OUT
$B --workload 1000x5x1000_b8 | summ b8 >> $OUT
python bench.py --workload 2000x5x20000_b64 --steps 5 --warmup 3 --no-cpu-baseline | summ 2000x5x20000_b64 >> $OUT
python bench.py --workload 10000x5x10000_b1024 --steps 3 --warmup 3 --no-cpu-baseline | summ config3 >> $OUT
cat $OUT
timeout 900 python -m pytest tests/test_gpu_big_configs.py tests/test_gpu_parity.py -x -q -m gpu -k "big or config or staged or benched" > gpurun_out/r02_tests_flush.log 2>&1; tail -5 gpurun_out/r02_tests_flush.log
python tests/tools/diag_big.py 2000 10000 2>&1 | tail -2
python tests/tools/diag_big.py 300 100000 2>&1 | tail -2
