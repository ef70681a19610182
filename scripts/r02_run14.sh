mkdir -p gpurun_out
python /tmp/oneq.py 2>/dev/null
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
print("one quiz: us per call %.2f" % ((time.perf_counter()-t0)/2000*1e6))
PY
python /tmp/oneq.py
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 > gpurun_out/r02_tests_v11.log; tail -3 gpurun_out/r02_tests_v11.log
