mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 > gpurun_out/r02_tests_v13.log; tail -3 gpurun_out/r02_tests_v13.log
cat > /tmp/t1.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa
for (Q,K,T,B) in ((1000,5,1000,256),(2000,5,100000,64)):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K,Q,T,init_amount=0.1), emulated_workers=16, rng_seed=3, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    ids = eng.start_quiz_batch(B)
    qs = np.arange(B) % Q
    def t(f, n=20):
        f(); eng.synchronize(); t0=time.perf_counter()
        for _ in range(n): f()
        eng.synchronize(); return (time.perf_counter()-t0)/n*1e3
    def ra():
        eng.set_active_question_batch(ids, qs); eng.record_answer_batch(ids, qs % K)
    print("T=%d B=%d: set_active+record_answer %.3f ms; list_top_targets(10) %.3f ms; start_quiz_batch %.3f ms" % (T, B, t(ra), t(lambda: eng.list_top_targets_batch(ids, 10)), t(lambda: eng.release_quiz_batch(eng.start_quiz_batch(B)))))
PY
python /tmp/t1.py
