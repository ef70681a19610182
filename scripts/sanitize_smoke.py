"""Tiny run that touches every kernel (all evaluation variants, chunked and single-chunk), meant to be run under
compute-sanitizer --tool memcheck / racecheck / initcheck on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probqa_b200 import engine as pqa, synth

Q, K, T, W = 12, 5, 70, 3
kb = synth.gamma_kb(Q, K, T, 0.1)
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=9)
eng.upload_kb(*kb)
ids = eng.start_quiz_batch(40)
for step in range(2):
    qs = eng.next_question_batch(ids)
    eng.record_answer_batch(ids, (qs + step) % K)
ref = None
for which, chunk, lanes, n in ((1, 0, 0, 40), (2, 0, 0, 3), (2, 0, 1, 40), (2, 0, 2, 40), (2, 0, 4, 40), (2, 32, 1, 5), (2, 32, 2, 40), (2, 32, 4, 40)):
    eng.set_eval_kernel(which, chunk, 0, lanes)
    pri = eng.eval_questions(ids[:n])["priority"]
    if ref is None:
        ref = pri
    ok = ~np.isnan(ref[:n])
    assert np.array_equal(np.isnan(pri), np.isnan(ref[:n]))
    assert np.allclose(pri[ok], ref[:n][ok], rtol=2e-12, atol=0), (which, chunk, lanes, n)
eng.set_eval_kernel(0)
eng.eval_questions_detailed(int(ids[0]))
eng.list_top_targets_batch(ids, 10)
eng.list_top_targets(int(ids[1]), 7)
eng.record_quiz_target_batch(ids, np.arange(40) % T)
eng.train([pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(3, 0)], 5, 1.5)
eng.resident_bind(ids); eng.resident_step(); eng.resident_fetch(); eng.flush_l2(); eng.synchronize()
eng.download_kb(); eng.copy_a_targets(2, 1)
eng.release_quiz_batch(ids)
print("sanitize smoke ok")
