#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Generates tests/golden/ref_vectors.npz from the REFERENCE'S OWN CODE (oracle/_ref/libpqa_ref.so = the subtask bodies of
/root/reference/ProbQA compiled where they lie by oracle/build_ref.sh). Run in the build container, where /root/reference
exists; the fixtures travel with the repo, so the oracle restatement (CPU tests) and the CUDA path (GPU tests) are checked
against the reference's outputs even where neither /root/reference nor oracle/_ref is present (tests/test_golden.py).

Contents, per case c (small synthetic KBs of probqa_b200/synth.py, emulated worker count W, optional gap masks):
  c/dims = [Q, K, T, W], c/kb_kind, c/tgaps, c/qgaps, c/aqs = the (question, answer) sequence of one quiz
  c/prior_<s>      posterior after s answers (s = 0 is StartQuiz)                      CEQuiz.h:77-122 etc.
  c/run_<s>, c/grand_<s>, c/bounds_<s>   NextQuestion run-lengths / chunk totals        CpuEngine.cpp:337-374
  c/top_<s>_idx, c/top_<s>_prob          ListTopTargets(10)                             CEListTopTargetsAlgorithm.cpp:30-97
  c/resume         ResumeQuiz priors for the whole sequence                             CECreateQuizOperation.cpp:55-83
  c/train_sA, c/train_mD, c/train_vB     KB after RecordQuizTarget(train_aqs, train_target, 0.75)   CETrainOperation.cpp:15-83
plus log2hot_x / log2hot_y (SRVectMath.h:87-135) and the PairSum known answers of SRAccumulatorTest.cpp:20-34."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from probqa_b200 import synth  # noqa: E402

CASES = [
    ("gamma_gaps", "gamma", 24, 5, 103, 4, True),
    ("binary", "binary", 32, 5, 200, 3, False),
    ("uniform", "uniform", 12, 3, 64, 2, False),
    ("gamma_w7", "gamma", 40, 4, 96, 7, False),
]
KB = {"binary": lambda Q, K, T: synth.binary_search_kb(Q, K, T, 0.1, 3), "gamma": lambda Q, K, T: synth.gamma_kb(Q, K, T, 0.1),
      "uniform": lambda Q, K, T: synth.uniform_kb(Q, K, T, 0.1)}


def main():
    ref.build()
    out = {}
    rng = np.random.default_rng(20171016)
    x = np.concatenate([rng.random(2048), np.exp2(rng.uniform(-60, 0, 2048)), [1.0, 0.5, 0.25, 0.0, 0.999, 0.1, 3.0, 1e-300, 5e-324]])
    out["log2hot_x"], out["log2hot_y"] = x, ref.log2hot(x)
    a = np.arange(1, 31, dtype=np.float64).reshape(-1, 2)              # SRAccumulatorTest.cpp:20-34 style inputs
    out["pairsum_in"] = a
    out["pairsum_out"] = np.array(ref.v4_pair_at(a[:, 0], a[:, 1]))
    for name, kind, Q, K, T, W, gaps in CASES:
        sA, mD, vB = KB[kind](Q, K, T)
        tg = qg = None
        if gaps:
            tg = np.zeros(T, dtype=bool); tg[[3, 17, 50, 51, 99]] = True
            qg = np.zeros(Q, dtype=bool); qg[[2, 11]] = True
        eng = ref.RefEngine(sA, mD, vB, W, qgaps=qg, tgaps=tg)
        aqs = [(q, a) for q, a in synth.quiz_prefix(5, 6, Q, T, K) if qg is None or not qg[q]][:5]
        out[name + "/dims"] = np.array([Q, K, T, W])
        out[name + "/kb_kind"] = np.array(kind)
        out[name + "/tgaps"] = np.zeros(T, dtype=bool) if tg is None else tg
        out[name + "/qgaps"] = np.zeros(Q, dtype=bool) if qg is None else qg
        out[name + "/aqs"] = np.array(aqs, dtype=np.int64)
        p = eng.start_quiz()
        asked = np.zeros(Q, dtype=bool)
        for s in range(len(aqs) + 1):
            out["%s/prior_%d" % (name, s)] = p.copy()
            ev = eng.eval_questions(p, asked)
            out["%s/run_%d" % (name, s)], out["%s/grand_%d" % (name, s)], out["%s/bounds_%d" % (name, s)] = ev["runLength"], ev["grand"], ev["bounds"]
            top = eng.list_top_targets(p, 10)
            out["%s/top_%d_idx" % (name, s)] = np.array([t for t, _ in top], dtype=np.int64)
            out["%s/top_%d_prob" % (name, s)] = np.array([pr for _, pr in top], dtype=np.float64)
            if s < len(aqs):
                p = eng.record_answer(p, *aqs[s])
                asked[aqs[s][0]] = True
        out[name + "/resume"] = eng.resume_quiz(aqs)
        target = int(np.flatnonzero(~out[name + "/tgaps"])[7])
        # consecutive pairs: two plain ones, same question + same answer, same question + different answers, a single
        train_aqs = aqs[:4] + [aqs[0], aqs[0]] + [aqs[1], (aqs[1][0], (aqs[1][1] + 1) % K)] + [aqs[2]]
        eng.record_quiz_target(train_aqs, target, 0.75)
        out[name + "/train_aqs"] = np.array(train_aqs, dtype=np.int64)
        out[name + "/train_target"] = np.array(target)
        out[name + "/train_sA"], out[name + "/train_mD"], out[name + "/train_vB"] = eng.read_kb()
        eng.close()
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d arrays, %.1f KB" % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
