"""Golden vectors produced by the REFERENCE'S OWN compiled code (tests/golden/make_golden.py -> tests/golden/ref_vectors.npz):
the CPU oracle must reproduce them bit for bit (-m "not gpu": this is what pins the oracle where neither /root/reference nor
oracle/_ref exists), and the CUDA path, through the C ABI, must meet its parity bars against them (-m gpu)."""
import os

import numpy as np
import pytest

from probqa_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files if "/" in k})
KB = {"binary": lambda Q, K, T: synth.binary_search_kb(Q, K, T, 0.1, 3), "gamma": lambda Q, K, T: synth.gamma_kb(Q, K, T, 0.1),
      "uniform": lambda Q, K, T: synth.uniform_kb(Q, K, T, 0.1)}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def case(name):
    Q, K, T, W = (int(v) for v in GOLD[name + "/dims"])
    kb = KB[str(GOLD[name + "/kb_kind"])](Q, K, T)
    tg, qg = GOLD[name + "/tgaps"], GOLD[name + "/qgaps"]
    aqs = [(int(q), int(a)) for q, a in GOLD[name + "/aqs"]]
    return Q, K, T, W, kb, (tg if tg.any() else None), (qg if qg.any() else None), aqs


def test_golden_file_is_complete():
    assert len(CASES) == 4 and all(name + "/resume" in GOLD.files for name in CASES)


def test_oracle_log2hot_matches_reference_vectors(ora):
    assert np.array_equal(bits(ora.log2hot(GOLD["log2hot_x"])), bits(GOLD["log2hot_y"]))
    assert GOLD["pairsum_out"].tolist() == [225.0, 240.0]      # sums of 1,3,..,29 and 2,4,..,30 (SRAccumulatorTest.cpp:20-34 style)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_vectors(ora, name):
    Q, K, T, W, (sA, mD, vB), tg, qg, aqs = case(name)
    p = ora.start_quiz(vB, W, tgaps=tg)
    asked = np.zeros(Q, dtype=bool)
    for s in range(len(aqs) + 1):
        assert np.array_equal(bits(p), bits(GOLD["%s/prior_%d" % (name, s)]))
        ev = ora.eval_questions(sA, mD, p, W, asked=asked, qgaps=qg, tgaps=tg)
        assert np.array_equal(ev["bounds"], GOLD["%s/bounds_%d" % (name, s)])
        assert np.array_equal(bits(ev["runLength"]), bits(GOLD["%s/run_%d" % (name, s)]))
        assert np.array_equal(bits(ev["grand"]), bits(GOLD["%s/grand_%d" % (name, s)]))
        top = ora.list_top_targets(p, W, 10, tgaps=tg)
        assert [t for t, _ in top] == GOLD["%s/top_%d_idx" % (name, s)].tolist()
        assert np.array_equal(bits([pr for _, pr in top]), bits(GOLD["%s/top_%d_prob" % (name, s)]))
        if s < len(aqs):
            q, a = aqs[s]
            p = ora.record_answer(p, sA[q, a], mD[q], max(1, W - 1), tgaps=tg)
            asked[q] = True
    assert np.array_equal(bits(ora.resume_quiz(sA, mD, vB, aqs, W, tgaps=tg)), bits(GOLD[name + "/resume"]))
    tA, tD, tB = sA.copy(), mD.copy(), vB.copy()
    ora.record_quiz_target(tA, tD, tB, [(int(q), int(a)) for q, a in GOLD[name + "/train_aqs"]], int(GOLD[name + "/train_target"]), 0.75)
    assert np.array_equal(bits(tA), bits(GOLD[name + "/train_sA"])) and np.array_equal(bits(tD), bits(GOLD[name + "/train_mD"]))
    assert np.array_equal(bits(tB), bits(GOLD[name + "/train_vB"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_against_reference_vectors(name):
    """Posteriors, top-10 lists, ResumeQuiz priors and the trained KB bit-exact; run-lengths of the exact kernel within
    1e-13 (device pow/exp2/log vs glibc), of the staged kernels within 2e-12."""
    from probqa_b200 import engine as pqa
    Q, K, T, W, kb, tg, qg, aqs = case(name)
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=1)
    eng.upload_kb(*kb)
    if tg is not None or qg is not None:
        eng.start_maintenance(False)
        if tg is not None:
            eng.remove_targets(np.flatnonzero(tg))
        if qg is not None:
            eng.remove_questions(np.flatnonzero(qg))
        eng.finish_maintenance()
    quiz = eng.start_quiz()
    for s in range(len(aqs) + 1):
        assert np.array_equal(bits(eng.copy_quiz_priors(quiz)), bits(GOLD["%s/prior_%d" % (name, s)]))
        want = GOLD["%s/run_%d" % (name, s)]
        for which, tol in ((1, 1e-13), (2, 2e-12)):
            eng.set_eval_kernel(which)
            got = eng.eval_questions([quiz])
            for g, w in ((got["runLength"][0], want), (got["grand"][0], GOLD["%s/grand_%d" % (name, s)])):
                nz = w != 0                      # a chunk that starts with asked / removed questions runs at exactly 0
                assert np.all(g[~nz] == 0) and np.max(np.abs(g[nz] - w[nz]) / np.abs(w[nz])) < tol, (which, s)
        top = eng.list_top_targets(quiz, 10)
        assert [r.i_target for r in top] == GOLD["%s/top_%d_idx" % (name, s)].tolist()
        assert np.array_equal(bits([r.prob for r in top]), bits(GOLD["%s/top_%d_prob" % (name, s)]))
        if s < len(aqs):
            eng.set_active_question(quiz, aqs[s][0])
            eng.record_answer(quiz, aqs[s][1])
    resumed = eng.resume_quiz([pqa.AnsweredQuestion(q, a) for q, a in aqs])
    assert np.array_equal(bits(eng.copy_quiz_priors(resumed)), bits(GOLD[name + "/resume"]))
    trainee = eng.resume_quiz([pqa.AnsweredQuestion(int(q), int(a)) for q, a in GOLD[name + "/train_aqs"]])
    eng.record_quiz_target(trainee, int(GOLD[name + "/train_target"]), 0.75)
    gA, gD, gB = eng.download_kb()
    live_t = np.ones(T, dtype=bool) if tg is None else ~tg      # cells of removed targets are not touched by either side
    assert np.array_equal(bits(gA[:, :, live_t]), bits(GOLD[name + "/train_sA"][:, :, live_t]))
    assert np.array_equal(bits(gD[:, live_t]), bits(GOLD[name + "/train_mD"][:, live_t]))
    assert np.array_equal(bits(gB[live_t]), bits(GOLD[name + "/train_vB"][live_t]))
