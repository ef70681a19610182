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
qs = eng.start_quiz_batch(4)
r = np.arange(4, dtype=np.uint64)
for n in (2, 4):
    for _ in range(50): eng.next_question_batch(qs[:n], r[:n])
    t0=time.perf_counter()
    for _ in range(1000): eng.next_question_batch(qs[:n], r[:n])
    print("  batch of %d: us per call %.2f" % (n, (time.perf_counter()-t0)/1000*1e6))
PY
python /tmp/oneq.py persist
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_eval_few -s 200 -c 2 --csv python scripts/oneq2.py 2>&1 | grep -i "k_eval_few" | awk -F'","' '{print $(NF-2), $(NF-1), $(NF)}'
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or selection or error" 2>&1 | tail -2
