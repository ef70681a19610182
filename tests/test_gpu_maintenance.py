"""GPU tests of maintenance mode: AddQsTs / RemoveQuestions / RemoveTargets / Compact, the gap masks in
the quiz kernels, the permanent <-> compact id maps and KB files with gaps -- through the C ABI, against the CPU oracle
(which takes the same gap masks) and against numpy restatements of the reference's bookkeeping
(CpuEngine.cpp:468-658, BaseEngine.cpp:640-779, GapTracker.h, PermanentIdManager.cpp)."""
import os

import numpy as np
import pytest

from probqa_b200 import synth

pytestmark = pytest.mark.gpu
INIT = 0.1


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.fixture(scope="module")
def pqa():
    from probqa_b200 import engine
    engine.load_library()
    return engine


def make_engine(pqa, Q, K, T, W, kb=None, init=INIT):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=init), emulated_workers=W, rng_seed=3)
    if kb is not None:
        eng.upload_kb(*kb)
    return eng


def test_dimensions_cpu_increase(pqa):
    """PqaCoreTests/Dimensions.cpp:11-89 (Dimensions.CpuIncrease) re-expressed: grow a 2x5x2 engine by 0..4 questions and
    targets per round; every cell must hold the initial amounts; a quiz is started and asked each round and force-closed
    by the next StartMaintenance(true)."""
    init = 1.0
    eng = make_engine(pqa, 2, 5, 2, 4, init=init)
    rng = np.random.default_rng(5)
    for _ in range(120):
        nq, nt = int(rng.integers(0, 5)), int(rng.integers(0, 5))
        before = eng.copy_dims()
        eng.start_maintenance(True)
        got_q, got_t = eng.add_qs_ts([init] * nq, [init] * nt)
        eng.finish_maintenance()
        after = eng.copy_dims()
        assert after.n_questions == before.n_questions + nq and after.n_targets == before.n_targets + nt
        assert got_q == list(range(before.n_questions, after.n_questions))
        assert got_t == list(range(before.n_targets, after.n_targets))
        sA, mD, vB = eng.download_kb()
        assert sA.shape == (after.n_questions, 5, after.n_targets)
        assert np.all(sA == init) and np.all(mD == 5 * init) and np.all(vB == init)
        quiz = eng.start_quiz()
        q = eng.next_question(quiz)
        assert 0 <= q < after.n_questions
    # the per-row copies of the reference test, on the final dimensions
    d = eng.copy_dims()
    for i in range(0, d.n_questions, 37):
        for k in range(5):
            assert np.all(eng.copy_a_targets(i, k) == init)
        assert np.all(eng.copy_d_targets(i) == 5 * init)
    assert np.all(eng.copy_b_targets() == init)


def test_mode_contract(pqa):
    """WrongMode (14) in both directions, QuizzesActive (21), MaintenanceModeAlreadyThis (6), AbsentId (13)."""
    eng = make_engine(pqa, 6, 3, 9, 2)
    quiz = eng.start_quiz()

    def code(err):
        return err.to_string(True)

    for call in (lambda: eng.add_qs_ts([1.0], [], throw=False), lambda: eng.remove_questions([0], throw=False),
                 lambda: eng.remove_targets([0], throw=False)):
        assert "maintenance-only mode operation" in code(call())
    with pytest.raises(pqa.PqaException):
        eng.compact()
    err = eng.start_maintenance(False, throw=False)
    assert err is not None and "nQuizzes=1" in code(err)          # quizzes alive, not forced: stays regular
    assert eng.next_question(quiz) >= 0
    eng.start_maintenance(True)                                    # forced: the quiz is destroyed
    assert eng.start_maintenance(True, throw=False) is not None    # already in maintenance
    for call in (lambda: eng.start_quiz(), lambda: eng.next_question(quiz), lambda: eng.list_top_targets(quiz, 3),
                 lambda: eng.record_quiz_target(quiz, 1), lambda: eng.release_quiz(quiz), lambda: eng.get_active_question_id(quiz),
                 lambda: eng.set_active_question(quiz, 0), lambda: eng.record_answer(quiz, 0)):
        with pytest.raises(pqa.PqaException) as ei:
            call()
        assert "regular-only mode operation" in str(ei.value)
    eng.train([(0, 1), (2, 0)], 4, 1.0)                            # Train is allowed in both modes (CpuEngine.cpp:136)
    eng.remove_targets([3])
    err = eng.remove_targets([3], throw=False)
    assert err is not None and "rather at a gap" in code(err)
    assert eng.remove_questions([17], throw=False) is not None
    eng.finish_maintenance()
    assert eng.finish_maintenance(throw=False) is not None
    with pytest.raises(pqa.PqaException):
        eng.next_question(quiz)                                    # the forced quiz is gone (AbsentId)
    err = eng.train([(0, 1)], 3, 1.0, throw=False)
    assert err is not None and "rather at a gap" in code(err)
    q2 = eng.start_quiz()
    assert q2 == quiz                                              # quiz ids are reused LIFO
    err = eng.record_quiz_target(q2, 3, throw=False)
    assert err is not None and "rather at a gap" in code(err)
    # permanent ids: the re-created quiz has a new permanent id; removed target 3 maps to nothing
    assert eng.quiz_perm_from_comp([q2])[0] == 1 and eng.quiz_comp_from_perm([0])[0] == -1
    assert eng.target_perm_from_comp([2, 3, 4]).tolist() == [2, -1, 4]
    assert eng.target_comp_from_perm([3])[0] == -1


@pytest.mark.parametrize("dims,W,kernel,n_quizzes,chunk", [((40, 5, 203), 6, 2, 5, 0), ((64, 5, 1000), 8, 2, 40, 0), ((24, 4, 96), 3, 1, 5, 0),
                                                           ((30, 5, 700), 5, 2, 70, 64), ((20, 5, 400), 4, 2, 33, 96)])
def test_gaps_parity_then_compact(pqa, ora, dims, W, kernel, n_quizzes, chunk):
    """Remove targets and questions; the quiz path must then equal the oracle run with the same gap masks (priors and
    top-10 bit-exact, priorities within the kernel's bar, gap questions never chosen). Compact; the compacted engine must
    equal the oracle on the numpy-compacted KB, and the old-id arrays must be the reference's (CpuEngine.cpp:585-658)."""
    Q, K, T = dims
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.set_eval_kernel(kernel, chunk_targets=chunk)     # n_quizzes / chunk pick the kernel shape: small, 4 threads per quiz,
    rng = np.random.default_rng(11)                      # 2 threads per quiz (8- and 4-warp CTAs), whole slab or chunked targets
    rm_t = rng.choice(T - 8, size=max(3, T // 9), replace=False)          # keeps the last targets: no trailing gaps
    rm_q = rng.choice(Q - 6, size=max(2, Q // 7), replace=False)
    eng.start_maintenance(False)
    eng.remove_targets(rm_t)
    eng.remove_questions(rm_q)
    eng.finish_maintenance()
    tg = np.zeros(T, dtype=bool); tg[rm_t] = True
    qg = np.zeros(Q, dtype=bool); qg[rm_q] = True
    sA, mD, vB = kb

    def check_session(eng, sA, mD, vB, tg, qg, n_steps=3):
        Qc, Tc = sA.shape[0], sA.shape[2]
        tgm, qgm = (tg if tg.any() else None), (qg if qg.any() else None)
        quizzes = eng.start_quiz_batch(n_quizzes)
        priors = [ora.start_quiz(vB, W, tgaps=tgm) for _ in quizzes]
        asked = [np.zeros(Qc, dtype=bool) for _ in quizzes]
        r2 = np.random.default_rng(12)
        for step in range(n_steps):
            ev = eng.eval_questions(quizzes)
            randoms = r2.integers(0, 2 ** 64, size=len(quizzes), dtype=np.uint64)
            chosen = eng.next_question_batch(quizzes, randoms)
            for x, quiz in enumerate(quizzes):
                got = eng.copy_quiz_priors(int(quiz))
                assert np.array_equal(bits(got), bits(priors[x])), "priors differ with gaps"
                assert np.all(got[tg] == 0)
                want = ora.eval_questions(sA, mD, priors[x], W, asked=asked[x], qgaps=qgm, tgaps=tgm, nThreads=4)
                skip = asked[x] | qg
                assert np.all(np.isnan(ev["priority"][x][skip]))
                rel = np.abs(ev["priority"][x][~skip] - want["priority"][~skip]) / np.abs(want["priority"][~skip])
                assert rel.max() < (1e-13 if kernel == 1 else 2e-12), rel.max()
                # the engine's own run-lengths decide its choice; the oracle's selection on them must agree
                sel = ora.select_question(dict(runLength=ev["runLength"][x], grand=ev["grand"][x], bounds=want["bounds"]), Qc,
                                          int(randoms[x]), asked=asked[x], qgaps=qgm)
                assert chosen[x] == sel and not qg[chosen[x]] and not asked[x][chosen[x]]
                top = [(r.i_target, r.prob) for r in eng.list_top_targets(int(quiz), 10)]
                assert top == ora.list_top_targets(priors[x], W, 10, tgaps=tgm)
                assert all(not tg[t] for t, _ in top)
            answers = [(int(c) * 5 + step) % K for c in chosen]
            eng.record_answer_batch(quizzes, answers)
            for x in range(len(quizzes)):
                q, a = int(chosen[x]), answers[x]
                priors[x] = ora.record_answer(priors[x], sA[q, a], mD[q], max(1, W - 1), tgaps=tgm)
                asked[x][q] = True
        for x, quiz in enumerate(quizzes):
            assert np.array_equal(bits(eng.copy_quiz_priors(int(quiz))), bits(priors[x]))
        eng.release_quiz_batch(quizzes)

    check_session(eng, sA, mD, vB, tg, qg)

    # ---- compaction
    def ref_compact_questions(n, gap):           # CpuEngine.cpp:594-608
        old, first, last = {}, 0, n - 1
        while first <= last:
            if not gap[first]:
                old[first] = first
            else:
                while gap[last] and last > first:
                    last -= 1
                if first == last:
                    break
                old[first] = last
                last -= 1
            first += 1
        return [old[i] for i in range(len(old))]

    def ref_compact_targets(n, gap):             # :616-641 (moves[] filled from both ends)
        old, dests, srcs, first, last = {}, [], [], 0, n - 1
        while first <= last:
            if not gap[first]:
                old[first] = first
            else:
                while gap[last] and last > first:
                    last -= 1
                if first == last:
                    break
                dests.append(first); srcs.append(last)
                last -= 1
            first += 1
        for g, d in enumerate(dests):
            old[d] = srcs[len(dests) - 1 - g]
        return [old[i] for i in range(len(old))]

    perm_q_before = eng.question_perm_from_comp(np.arange(Q))
    perm_t_before = eng.target_perm_from_comp(np.arange(T))
    eng.start_maintenance(True)
    old_q, old_t = eng.compact()
    eng.finish_maintenance()
    assert old_q == ref_compact_questions(Q, qg) and old_t == ref_compact_targets(T, tg)
    assert len(old_q) == Q - len(rm_q) and len(old_t) == T - len(rm_t)
    d = eng.copy_dims()
    assert (d.n_questions, d.n_targets) == (len(old_q), len(old_t))
    cA, cD, cB = sA[old_q][:, :, old_t], mD[old_q][:, old_t], vB[old_t]
    gA, gD, gB = eng.download_kb()
    assert np.array_equal(bits(gA), bits(cA)) and np.array_equal(bits(gD), bits(cD)) and np.array_equal(bits(gB), bits(cB))
    # permanent ids follow their rows (PermanentIdManager::OnCompact)
    assert np.array_equal(eng.question_perm_from_comp(np.arange(len(old_q))), perm_q_before[old_q])
    assert np.array_equal(eng.target_perm_from_comp(np.arange(len(old_t))), perm_t_before[old_t])
    assert np.array_equal(eng.target_comp_from_perm(perm_t_before[old_t]), np.arange(len(old_t)))
    check_session(eng, np.ascontiguousarray(cA), np.ascontiguousarray(cD), np.ascontiguousarray(cB),
                  np.zeros(len(old_t), dtype=bool), np.zeros(len(old_q), dtype=bool), n_steps=2)


def test_add_reuses_gaps_and_initial_amounts(pqa):
    """AddQsTs after removals: removed ids come back LIFO before new ones are appended; cell values follow
    CpuEngine.cpp:496-578 (question rows carry the question's amount in every column, target columns carry the target's
    amount in the rows of the questions that existed before the call; re-added question rows win)."""
    Q, K, T = 7, 3, 10
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, 2, kb)
    sA, mD, vB = [a.copy() for a in kb]
    eng.start_maintenance(False)
    eng.remove_questions([1, 4])
    eng.remove_targets([2, 8])
    got_q, got_t = eng.add_qs_ts([0.5, 0.7, 0.9], [0.3, 0.4, 0.6])       # equal reuse counts: the reference's indexing is well-defined
    eng.finish_maintenance()
    assert got_q == [4, 1, Q] and got_t == [8, 2, T]                      # gaps LIFO, then appended
    nA = np.zeros((Q + 1, K, T + 1)); nD = np.zeros((Q + 1, T + 1)); nB = np.zeros(T + 1)
    nA[:Q, :, :T], nD[:Q, :T], nB[:T] = sA, mD, vB
    for t, amt in zip(got_t, [0.3, 0.4, 0.6]):                            # target columns in the old questions' rows
        nA[:Q, :, t] = amt * amt; nD[:Q, t] = amt * amt * K; nB[t] = amt
    for q, amt in zip(got_q, [0.5, 0.7, 0.9]):                            # question rows, all columns
        nA[q] = amt * amt; nD[q] = amt * amt * K
    gA, gD, gB = eng.download_kb()
    assert np.array_equal(bits(gA), bits(nA)) and np.array_equal(bits(gD), bits(nD)) and np.array_equal(bits(gB), bits(nB))
    # re-added ids get fresh permanent ids after the ones handed out so far
    assert eng.question_perm_from_comp(got_q).tolist() == [Q, Q + 1, Q + 2]
    assert eng.target_perm_from_comp(got_t).tolist() == [T, T + 1, T + 2]
    assert eng.question_comp_from_perm([1, 4]).tolist() == [-1, -1]
    quiz = eng.start_quiz()
    assert 0 <= eng.next_question(quiz) <= Q


def test_kb_file_with_gaps_roundtrip(pqa, ora, tmp_path):
    """SaveKB / LoadCpuEngine with removed questions and targets: the gap lists and id maps are written in the reference's
    layout (BaseEngine.cpp:142-152,323-385; PermanentIdManager.cpp:27-39) and a loaded engine behaves identically."""
    Q, K, T, W = 12, 4, 30, 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.start_maintenance(False)
    eng.remove_targets([5, 17, 3])
    eng.remove_questions([7])
    eng.finish_maintenance()
    path = str(tmp_path / "gaps.kb")
    eng.save_kb(path)
    raw = open(path, "rb").read()
    off = 8 + 24 + 8 + 8 * (Q * K * T + Q * T + T)
    tail = np.frombuffer(raw[off:], dtype=np.int64)
    want_tail = [1, 7, 3, 5, 17, 3]                                        # question gaps {n, ids}, target gaps {n, ids (LIFO order)}
    pq = list(range(Q)); pq[7] = -1
    pt = list(range(T)); pt[5] = pt[17] = pt[3] = -1
    want_tail += [Q, Q] + pq + [T, T] + pt + [0, 0]                        # id maps {nextPerm, nComp, comp2perm}; quizzes saved empty
    assert tail.tolist() == want_tail
    env_w = os.environ.get("PQA_B200_EMULATED_WORKERS")
    os.environ["PQA_B200_EMULATED_WORKERS"] = str(W)
    try:
        eng2, _ = pqa.PqaEngineFactory().load_cpu_engine(path)
    finally:
        if env_w is None:
            os.environ.pop("PQA_B200_EMULATED_WORKERS", None)
        else:
            os.environ["PQA_B200_EMULATED_WORKERS"] = env_w
    tg = np.zeros(T, dtype=bool); tg[[5, 17, 3]] = True
    for e in (eng, eng2):
        quiz = e.start_quiz()
        prior = e.copy_quiz_priors(quiz)
        assert np.array_equal(bits(prior), bits(ora.start_quiz(kb[2], W, tgaps=tg)))
        assert e.target_perm_from_comp([3, 4, 5]).tolist() == [-1, 4, -1]
        err = e.train([(7, 0)], 1, 1.0, throw=False)
        assert err is not None                                             # question 7 is a gap
    eng2.start_maintenance(True)
    got_q, got_t = eng2.add_qs_ts([1.0], [1.0, 1.0])
    assert got_q == [7] and got_t == [3, 17]                               # the loaded gap lists keep their LIFO order


def test_clear_old_quizzes(pqa):
    """BaseEngine::ClearOldQuizzes (BaseEngine.cpp:814-872): age limit first, then the oldest quizzes beyond maxCount."""
    import time
    eng = make_engine(pqa, 6, 3, 9, 2)
    old = [eng.start_quiz() for _ in range(3)]
    time.sleep(2.1)
    young = [eng.start_quiz() for _ in range(4)]
    eng.next_question(old[1])                       # using a quiz refreshes its age
    with pytest.raises(pqa.PqaException):
        eng.clear_old_quizzes(-1, 10.0)
    eng.clear_old_quizzes(100, 1.5)                 # older than 1.5 s: old[0] and old[2]
    for q in (old[0], old[2]):
        with pytest.raises(pqa.PqaException):
            eng.next_question(q)
    for q in [old[1]] + young:
        eng.list_top_targets(q, 2)
    eng.clear_old_quizzes(2, 1e9)                   # keep at most two
    alive = 0
    for q in [old[1]] + young:
        try:
            eng.list_top_targets(q, 2)
            alive += 1
        except pqa.PqaException:
            pass
    assert alive == 2
    eng.start_maintenance(True)
    assert eng.clear_old_quizzes(0, 0.0, throw=False) is None      # a no-op outside regular mode
