"""Diagnostic (not a test): component-wise relative differences of the staged kernel against the oracle at a big T."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from probqa_b200 import engine as pqa, synth
from oracle import oracle as ora

Q, K, T, W, B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 5, int(sys.argv[2]) if len(sys.argv) > 2 else 10000, 8, 130
ora.build()
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=3,
                                                initial_quiz_capacity=B)
eng.fill_binary_search_kb(3)
quizzes = eng.start_quiz_batch(B)
states = [synth.quiz_prefix(b, (0, 3, 8)[b % 3], Q, T, K) for b in range(B)]
for s in range(8):
    sel = [x for x in range(B) if len(states[x]) > s]
    eng.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
    eng.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
det = eng.eval_questions_detailed_batch(quizzes)
rng = np.random.default_rng(31)
sample_q = np.unique(rng.integers(0, Q, 60))
worst = dict(W=0, H=0, V=0, lack=0, priority=0)
for x in (0, 1, 2, 64, 129):
    prior = eng.copy_quiz_priors(int(quizzes[x]))
    for i in sample_q:
        i = int(i)
        if np.isnan(det["priority"][x, i]):
            continue
        a = np.stack([eng.copy_a_targets(i, k) for k in range(K)]); d = eng.copy_d_targets(i)
        o = ora.eval_question(a, d, prior)
        for key in ("W", "H", "V"):
            worst[key] = max(worst[key], float(np.max(np.abs(det[key][x, i] - o[key]) / np.abs(o[key]))))
        worst["lack"] = max(worst["lack"], abs(det["lack"][x, i] - o["lack"]) / abs(o["lack"]))
        r = abs(det["priority"][x, i] - o["priority"]) / abs(o["priority"])
        if r > worst["priority"]:
            worst["priority"] = r
            wc = (x, i, [float(v) for v in (det["H"][x, i] - o["H"]) / o["H"]], [float(v) for v in (det["V"][x, i] - o["V"]) / o["V"]],
                  float((det["lack"][x, i] - o["lack"]) / o["lack"]), [float(v) for v in o["H"]], [float(v) for v in o["W"]])
print("Q=%d T=%d worst relative differences:" % (Q, T), {k: "%.3g" % v for k, v in worst.items()})
print("worst priority case (quiz, question, relH[k], relV[k], relLack, H[k], W[k]):", wc)
