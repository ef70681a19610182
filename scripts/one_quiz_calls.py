"""Per-call wall time of the reference ABI's one-quiz calls (NextQuestion / RecordAnswer / ListTopTargets) at 1000x5x1000.
PQA_B200_LIB selects the library (A/B of builds)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from probqa_b200 import engine as pqa, synth

Q, K, T = 1000, 5, 1000
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=16, rng_seed=3)
eng.upload_kb(*synth.binary_search_kb(Q, K, T, 0.1, 3))
res = {"next_question": [], "record_answer": [], "list_top_targets": []}
for rep in range(3):
    tn = ta = tl = 0.0
    n = 0
    for quiz_no in range(60):
        q = eng.start_quiz()
        for step in range(20):
            t0 = time.perf_counter(); i = eng.next_question(q)
            t1 = time.perf_counter(); eng.record_answer(q, (i + step) % K)
            t2 = time.perf_counter(); eng.list_top_targets(q, 1)
            t3 = time.perf_counter()
            tn += t1 - t0; ta += t2 - t1; tl += t3 - t2; n += 1
        eng.release_quiz(q)
    res["next_question"].append(tn / n * 1e6); res["record_answer"].append(ta / n * 1e6); res["list_top_targets"].append(tl / n * 1e6)
print({k: [round(x, 2) for x in v] for k, v in res.items()})
