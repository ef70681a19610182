"""TEST INFRASTRUCTURE. Diagnostic: exact kernel vs oracle at full size, which component differs (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from probqa_b200 import engine as pqa, synth
from oracle import oracle as ora

Q, K, T, W = 1000, 5, 1000, 8
kb = synth.binary_search_kb(Q, K, T, 0.1, 3)
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=1)
eng.upload_kb(*kb)
quiz = eng.start_quiz()
for q, a in synth.quiz_prefix(0, 3, Q, T, K):
    eng.set_active_question(quiz, q); eng.record_answer(quiz, a)
prior = eng.copy_quiz_priors(quiz)
eng.set_eval_kernel(1); ex = eng.eval_questions_detailed(quiz)
bits = lambda a: np.asarray(a, dtype=np.float64).view(np.int64)
nbad = 0
for i in range(Q):
    if np.isnan(ex["priority"][i]): continue
    o = ora.eval_question(kb[0][i], kb[1][i], prior)
    d = abs(int(bits(ex["priority"][i])) - int(bits(o["priority"])))
    if d > 8:
        nbad += 1
        if nbad <= 5:
            print(i, "ulp", d, "W", np.array_equal(bits(ex["W"][i]), bits(o["W"])), "H", np.array_equal(bits(ex["H"][i]), bits(o["H"])),
                  "V", np.array_equal(bits(ex["V"][i]), bits(o["V"])), "lack", bits(ex["lack"][i]) == bits(o["lack"]))
            print("   V", ex["V"][i], o["V"], "H", ex["H"][i], "W", ex["W"][i])
            tw = o["totW"]; 
            avgV = float(np.sum(o["W"] * np.sqrt(o["V"])) / tw); avgH = float(np.sum(o["W"] * o["H"]) / tw)
            print("   avgV", avgV, "avgH", avgH, "pri", ex["priority"][i], o["priority"])
print("nbad", nbad)
