mkdir -p gpurun_out
OUT=gpurun_out/r02_variants4.log; : > $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-28s value %.4g  kernel_ms %.4f  e2e %.4g' % (sys.argv[1], d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))
" "$1"; }
$B | summ base_a1d1 >> $OUT
for v in a0d1 a1d0 a0d0; do
  PQA_B200_LIB=probqa_b200/lib/exp/$v/libPqaCore.so $B 2>&1 | summ $v >> $OUT
done
cat $OUT
