#!/bin/bash
# e2e time of one NextQuestion batch of 4 ... 64 quizzes at 1000x5x1000 for settings "MULTI_MIN:FEW_MAX:KL1_MAX"
# (medium-batch kernel from MULTI_MIN quizzes, fused launch up to FEW_MAX, four threads per quiz up to KL1_MAX)
for cfg in ${CFGS:-5:4:16 1000:16:16}; do
  IFS=: read mm fm k1 <<< "$cfg"
  echo "### PQA_B200_MULTI_MIN=$mm PQA_B200_FEW_MAX=$fm PQA_B200_MULTI_KL1_MAX=$k1"
  for b in ${BATCHES:-4 8 16 24 32 64}; do
    PQA_B200_MULTI_MIN=$mm PQA_B200_FEW_MAX=$fm PQA_B200_MULTI_KL1_MAX=$k1 python bench.py --workload 1000x5x1000_b$b --steps 100 --warmup 10 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  B=%-4d resident %.4f ms   e2e %.4f ms' % ($b, d['ms_per_step'], d['e2e']['ms_per_step']))"
  done
done
