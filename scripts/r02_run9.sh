mkdir -p gpurun_out
cat > /tmp/oneq.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth
Q,K,T=1000,5,1000
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K,Q,T,init_amount=0.1), emulated_workers=16, rng_seed=3)
eng.upload_kb(*synth.binary_search_kb(Q,K,T,0.1,3))
q = eng.start_quiz()
for _ in range(50): eng.next_question(q)
t0=time.perf_counter()
for _ in range(500): eng.next_question(q)
print("us per call", (time.perf_counter()-t0)/500*1e6)
PY
python /tmp/oneq.py
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_eval_few -s 20 -c 5 --csv python /tmp/oneq.py 2>&1 | grep -i "k_eval_few" | awk -F, '{print $(NF)}' | tr '\n' ' '
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or selection or staged_kernel" 2>&1 | tail -2
