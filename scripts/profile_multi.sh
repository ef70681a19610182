#!/bin/bash
# ncu --set full of k_eval_multi at 1000x5x1000 for the batch sizes in $1 (default "8 16 32"); summaries -> gpurun_out/
mkdir -p gpurun_out /tmp/prof
for b in ${1:-8 16 32}; do
cat > /tmp/multi.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth
Q, K, T, B = 1000, 5, 1000, $b
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=16, rng_seed=3)
eng.fill_binary_search_kb(3)
ids = eng.start_quiz_batch(B)
r = np.arange(B, dtype=np.uint64) * 7919
for step in range(4):
    ch = eng.next_question_batch(ids, r)
    eng.record_answer_batch(ids, [int(c) % K for c in ch])
PY
ncu --set full --import-source on --clock-control none -k regex:k_eval_multi -s 3 -c 1 -o /tmp/prof/multi_$b -f python /tmp/multi.py > gpurun_out/multi_ncu_$b.log 2>&1
python scripts/ncu_summary.py /tmp/prof/multi_$b.ncu-rep > gpurun_out/r02_eval_multi_b$b.txt 2>&1
cp /tmp/prof/multi_$b.ncu-rep gpurun_out/
done
tail -30 gpurun_out/r02_eval_multi_b*.txt
