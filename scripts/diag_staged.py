"""Diagnostic: which component of the staged kernel deviates from the exact kernel (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probqa_b200 import engine as pqa, synth

Q, K, T, W = 48, 5, 1000, 8
kb = synth.binary_search_kb(Q, K, T, 0.1, 3)
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=1)
eng.upload_kb(*kb)
for b, depth in enumerate((0, 1, 3, 5, 2)):
    quiz = eng.start_quiz()
    for q, a in synth.quiz_prefix(b, depth, Q, T, K):
        eng.set_active_question(quiz, q); eng.record_answer(quiz, a)
    eng.set_eval_kernel(1); ex = eng.eval_questions_detailed(quiz)
    eng.set_eval_kernel(2); st = eng.eval_questions_detailed(quiz)
    ok = ~np.isnan(ex["priority"])
    def rel(name):
        a, b_ = st[name][ok], ex[name][ok]
        return float(np.nanmax(np.abs(a - b_) / np.abs(b_)))
    print("depth", depth, {n: "%.2e" % rel(n) for n in ("W", "H", "V", "lack", "priority")})
    i = int(np.nanargmax(np.where(ok, np.abs(st["priority"] - ex["priority"]) / np.abs(ex["priority"]), 0)))
    print("  worst question", i, "H", st["H"][i], ex["H"][i], "V", st["V"][i], ex["V"][i], "lack", st["lack"][i], ex["lack"][i])
