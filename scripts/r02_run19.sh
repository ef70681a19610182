for fm in 4 64; do for w in 1000x5x1000_b8 1000x5x1000_b16 1000x5x1000_b32 1000x5x1000_b64; do
PQA_B200_FEW_MAX=$fm python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('few_max=%s %-18s e2e ms/step %.4f  e2e q-evals/s %.4g   (resident: %.4f ms)' % (sys.argv[1], d['config']['workload'], d['e2e']['ms_per_step'], d['e2e']['value'], d['ms_per_step']))
" $fm; done; done
PQA_B200_FEW_MAX=64 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
