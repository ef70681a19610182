mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 > gpurun_out/r02_tests_v9.log; tail -5 gpurun_out/r02_tests_v9.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_v9.log 2>&1; python -c "
import json
for l in open('gpurun_out/r02_bench_v9.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g kernel_ms %.4f' % (d['value'], d['roofline']['kernel_ms'])); print(d['single_quiz'])
"
