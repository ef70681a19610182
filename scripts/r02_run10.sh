cat > /tmp/oneq.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth
Q,K,T=1000,5,1000
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K,Q,T,init_amount=0.1), emulated_workers=16, rng_seed=3)
eng.upload_kb(*synth.binary_search_kb(Q,K,T,0.1,3))
q = eng.start_quiz()
for _ in range(100): eng.next_question(q)
t0=time.perf_counter()
for _ in range(2000): eng.next_question(q)
print("%-14s us per call %.2f" % (sys.argv[1], (time.perf_counter()-t0)/2000*1e6))
PY
python /tmp/oneq.py base
for v in few_pf1b4 few_b4 few_b16 few_nop1 few_nop2 few_nosel few_nop12 few_nothing; do PQA_B200_LIB=probqa_b200/lib/exp/$v/libPqaCore.so python /tmp/oneq.py $v; done
timeout 600 python -m pytest tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -3
