"""GPU parity tests: the CUDA path, called through the C ABI of libPqaCore.so, against the CPU oracle (oracle/).

Bars (DESIGN.md "Parity"):
  * priors after StartQuiz / RecordAnswer ........ bit-exact (Kahan order of CpuEngine emulated for worker count W)
  * ListTopTargets ............................... identical indices, order and probability bits (ties included)
  * exact evaluation kernel (which=1) ............ W_k, H_k, V_k, lack bit-exact; priority <= 8 ulp (device libm pow/exp2/log)
  * staged evaluation kernel (which=2, default) .. W_k bit-exact (4-lane Kahan order reproduced); H_k, V_k, lack 1e-12
        relative; priority 2e-12 relative, flat (no conditioning term: posteriors are bit-exact because W_k is)
  * question selection ........................... identical index for identical run-lengths and the same 64-bit draw
  * RecordQuizTarget / Train ..................... sA, mD, vB bit-exact
"""
import numpy as np
import pytest

from probqa_b200 import synth

pytestmark = pytest.mark.gpu

INIT = 0.1


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))   # same-signed finite values only


@pytest.fixture(scope="module")
def pqa():
    from probqa_b200 import engine
    engine.load_library()
    return engine


def make_engine(pqa, Q, K, T, W, kb=None, seed=12345):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT),
                                                    emulated_workers=W, rng_seed=seed)
    if kb is not None:
        eng.upload_kb(*kb)
    return eng


KBS = {
    "binary": lambda Q, K, T: synth.binary_search_kb(Q, K, T, INIT, 3),
    "gamma": lambda Q, K, T: synth.gamma_kb(Q, K, T, INIT),
    "uniform": lambda Q, K, T: synth.uniform_kb(Q, K, T, INIT),
}


def drive_quiz(eng, ora, kb, W, quiz, prefix, check_top=True):
    """Answers `prefix` in quiz, checking priors and top-10 against the oracle after every step. Returns the prior."""
    sA, mD, vB = kb
    prior = ora.start_quiz(vB, W)
    got = eng.copy_quiz_priors(quiz)
    assert np.array_equal(bits(got), bits(prior)), "StartQuiz priors differ"
    for (q, a) in prefix:
        eng.set_active_question(quiz, q)
        eng.record_answer(quiz, a)
        prior = ora.record_answer(prior, sA[q, a], mD[q], max(1, W - 1))
        got = eng.copy_quiz_priors(quiz)
        assert np.array_equal(bits(got), bits(prior)), "RecordAnswer priors differ at question %d" % q
        if check_top:
            want = ora.list_top_targets(prior, W, 10)
            have = eng.list_top_targets(quiz, 10)
            assert [(r.i_target, r.prob) for r in have] == want
    return prior


@pytest.mark.parametrize("kbname", ["binary", "gamma", "uniform"])
@pytest.mark.parametrize("dims,W", [((64, 5, 1000), 8), ((33, 5, 203), 7), ((16, 3, 50), 1), ((20, 5, 1000), 64)])
def test_priors_and_top_targets_bit_exact(pqa, ora, kbname, dims, W):
    Q, K, T = dims
    kb = KBS[kbname](Q, K, T)
    eng = make_engine(pqa, Q, K, T, W, kb)
    for b in range(3):
        quiz = eng.start_quiz()
        want = ora.list_top_targets(ora.start_quiz(kb[2], W), W, 10)
        have = eng.list_top_targets(quiz, 10)
        assert [(r.i_target, r.prob) for r in have] == want   # all-ties case: pure heap mechanics
        drive_quiz(eng, ora, kb, W, quiz, synth.quiz_prefix(b, min(6, Q - 1), Q, T, K))
        eng.release_quiz(quiz)


def test_top_targets_more_than_positive(pqa, ora):
    Q, K, T, W = 8, 5, 40, 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    quiz = eng.start_quiz()
    prior = np.zeros(T)
    prior[[3, 17, 18, 39]] = [0.25, 0.5, 0.125, 0.125]
    eng.set_quiz_priors(quiz, prior)
    have = eng.list_top_targets(quiz, 10)
    want = ora.list_top_targets(prior, W, 10)
    assert len(have) == 4 and [(r.i_target, r.prob) for r in have] == want
    # whole list, with ties
    prior = np.full(T, 1.0 / T)
    eng.set_quiz_priors(quiz, prior)
    assert [(r.i_target, r.prob) for r in eng.list_top_targets(quiz, T)] == ora.list_top_targets(prior, W, T)


def oracle_eval_all(ora, kb, prior, W, asked=None):
    sA, mD, _ = kb
    return ora.eval_questions(sA, mD, prior, W, asked=asked, nThreads=8)


@pytest.mark.parametrize("kbname", ["binary", "gamma", "uniform"])
@pytest.mark.parametrize("dims,W", [((48, 5, 1000), 8), ((21, 5, 203), 3), ((12, 2, 64), 1), ((10, 9, 77), 2)])
def test_exact_kernel_bit_exact(pqa, ora, kbname, dims, W):
    Q, K, T = dims
    kb = KBS[kbname](Q, K, T)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.set_eval_kernel(1)
    for depth in (0, 3):
        quiz = eng.start_quiz()
        prefix = synth.quiz_prefix(depth + 1, min(depth, Q - 1), Q, T, K)
        prior = drive_quiz(eng, ora, kb, W, quiz, prefix, check_top=False)
        asked = np.zeros(Q, dtype=bool)
        asked[[q for q, _ in prefix]] = True
        det = eng.eval_questions_detailed(quiz)
        for i in range(Q):
            if asked[i]:
                assert np.isnan(det["priority"][i])
                continue
            o = ora.eval_question(kb[0][i], kb[1][i], prior)
            assert np.array_equal(bits(det["W"][i]), bits(o["W"])), (i, det["W"][i], o["W"])
            assert np.array_equal(bits(det["H"][i]), bits(o["H"])), i
            assert np.array_equal(bits(det["V"][i]), bits(o["V"])), i
            assert bits(det["lack"][i]) == bits(o["lack"]), i
            assert ulp_diff(det["priority"][i], o["priority"]) <= 8, (i, det["priority"][i], o["priority"])
        # run-lengths, grand totals and selection through the same API
        ev = eng.eval_questions([quiz])
        oev = oracle_eval_all(ora, kb, prior, W, asked)
        assert np.allclose(ev["runLength"][0], oev["runLength"], rtol=1e-14, atol=0)
        assert np.allclose(ev["grand"][0], oev["grand"], rtol=1e-14, atol=0)
        eng.release_quiz(quiz)


TOL_STAGED = 2e-12   # relative, per question priority (DESIGN.md "Parity"); measured max ~1.5e-13


def staged_tolerance(kb, prior):
    """Flat tolerance of the staged kernel. It needs no conditioning term because the kernel reproduces the
    reference's normaliser W_k bit for bit, hence posteriors and (post - prior) bit for bit."""
    return np.full(kb[0].shape[0], TOL_STAGED)


@pytest.mark.parametrize("kbname", ["binary", "gamma", "uniform"])
@pytest.mark.parametrize("dims,W,chunk", [((48, 5, 1000), 8, 0), ((48, 5, 1000), 8, 96), ((21, 5, 203), 3, 0),
                                          ((21, 5, 203), 3, 32), ((12, 2, 64), 1, 0), ((9, 8, 130), 2, 64),
                                          ((10, 9, 77), 2, 0)])
@pytest.mark.parametrize("lanes", [0, 1, 4])   # auto (small-batch kernel here) / four threads per quiz / one thread per quiz
def test_staged_kernel_within_tolerance(pqa, ora, kbname, dims, W, chunk, lanes):
    Q, K, T = dims
    kb = KBS[kbname](Q, K, T)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.set_eval_kernel(2, chunk_targets=chunk, kahan_lanes_per_thread=lanes)
    quizzes, priors, askeds = [], [], []
    for b, depth in enumerate((0, 1, 3, 5, 2)):   # odd batch: exercises the half-filled quiz pair
        quiz = eng.start_quiz()
        prefix = synth.quiz_prefix(b, min(depth, Q - 1), Q, T, K)
        priors.append(drive_quiz(eng, ora, kb, W, quiz, prefix, check_top=False))
        asked = np.zeros(Q, dtype=bool)
        asked[[q for q, _ in prefix]] = True
        askeds.append(asked)
        quizzes.append(quiz)
    ev = eng.eval_questions(quizzes)
    single = eng.eval_questions(quizzes[:1])     # n = 1 path (one quiz per warp)
    for x, quiz in enumerate(quizzes):
        oev = oracle_eval_all(ora, kb, priors[x], W, askeds[x])
        tol = staged_tolerance(kb, priors[x])
        got, want = ev["priority"][x], oev["priority"]
        assert np.array_equal(np.isnan(got), askeds[x])
        ok = ~askeds[x]
        rel = np.abs(got[ok] - want[ok]) / np.abs(want[ok])
        assert np.all(rel <= tol[ok]), (x, float(np.max(rel / tol[ok])), float(rel.max()))
        if x == 0:
            assert np.array_equal(np.isnan(single["priority"][0]), askeds[0])
            rel1 = np.abs(single["priority"][0][ok] - want[ok]) / np.abs(want[ok])
            assert np.all(rel1 <= tol[ok])
        # run-lengths / grand totals inherit the tolerance
        assert np.allclose(ev["runLength"][x], oev["runLength"], rtol=float(tol[ok].max()) * 4, atol=0)
        assert np.allclose(ev["grand"][x], oev["grand"], rtol=float(tol[ok].max()) * 4, atol=0)


def test_staged_matches_detail_outputs(pqa, ora):
    Q, K, T, W = 40, 5, 1000, 8
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = make_engine(pqa, Q, K, T, W, kb)
    quiz = eng.start_quiz()
    prior = drive_quiz(eng, ora, kb, W, quiz, synth.quiz_prefix(7, 3, Q, T, K), check_top=False)
    det = eng.eval_questions_detailed(quiz)   # default kernel = staged
    for i in range(Q):
        if np.isnan(det["priority"][i]):
            continue
        o = ora.eval_question(kb[0][i], kb[1][i], prior)
        assert np.array_equal(bits(det["W"][i]), bits(o["W"]))
        assert np.allclose(det["H"][i], o["H"], rtol=1e-12, atol=0)
        assert np.allclose(det["V"][i], o["V"], rtol=1e-12, atol=0)
        assert abs(det["lack"][i] - o["lack"]) <= 1e-12 * abs(o["lack"])


def test_zero_prior_targets_follow_log2hot_edge(pqa, ora):
    """Posterior exactly 0 contributes invD^2 / Log2Hot(0) = invD^2 / -1023 to lack (SURVEY hard part 5)."""
    Q, K, T, W = 6, 5, 64, 2
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    quiz = eng.start_quiz()
    prior = np.random.default_rng(5).uniform(0.1, 1, T)
    prior[::3] = 0.0
    prior[5] = 1e-320   # subnormal posterior
    prior /= prior.sum()
    eng.set_quiz_priors(quiz, prior)
    for which, tol in ((1, 0.0), (2, 1e-12)):
        eng.set_eval_kernel(which)
        det = eng.eval_questions_detailed(quiz)
        for i in range(Q):
            o = ora.eval_question(kb[0][i], kb[1][i], prior)
            if which == 1:
                assert bits(det["lack"][i]) == bits(o["lack"])
            assert abs(det["lack"][i] - o["lack"]) <= tol * abs(o["lack"])
            assert abs(det["priority"][i] - o["priority"]) <= max(tol, 1e-15) * abs(o["priority"])


@pytest.mark.parametrize("which", [1, 2])
def test_next_question_selection(pqa, ora, which):
    Q, K, T, W = 96, 5, 300, 4
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.set_eval_kernel(which)
    rng = np.random.default_rng(11)
    n = 24
    quizzes = eng.start_quiz_batch(n)
    askeds = np.zeros((n, Q), dtype=bool)
    for x, quiz in enumerate(quizzes):
        for (q, a) in synth.quiz_prefix(x, x % 5, Q, T, K):
            eng.set_active_question(int(quiz), q)
            eng.record_answer(int(quiz), a)
            askeds[x, q] = True
    randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    randoms[0], randoms[1] = 0, 2 ** 64 - 1
    ev = eng.eval_questions(quizzes)
    asked_before = eng.get_total_questions_asked()
    chosen = eng.next_question_batch(quizzes, randoms)
    assert eng.get_total_questions_asked() == asked_before + n
    bounds = ora.calc_split(Q, 8 * W)
    for x in range(n):
        e1 = dict(runLength=ev["runLength"][x], grand=ev["grand"][x], bounds=bounds)
        want = ora.select_question(e1, Q, int(randoms[x]), asked=askeds[x])
        assert chosen[x] == want, (x, chosen[x], want)
        assert not askeds[x, chosen[x]]
        assert eng.get_active_question_id(int(quizzes[x])) == chosen[x]


@pytest.mark.parametrize("dims,W", [((96, 5, 300), 4), ((1000, 5, 1000), 16), ((40, 3, 203), 2)])
def test_one_quiz_fused_launch(pqa, ora, dims, W):
    """The reference ABI's call shape -- PqaEngine_NextQuestion, one quiz per call (and the 2..4 a combiner round may
    collect) -- runs as ONE fused launch (k_eval_few: evaluation from the derived KB, selection by the last CTA, result in
    mapped host memory). Its priorities (W_k bit-exact inside) against the oracle at the staged bar, its chosen question
    against the oracle's selection for the same draw, and the quiz state it leaves."""
    Q, K, T = dims
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = make_engine(pqa, Q, K, T, W, kb)
    rng = np.random.default_rng(17)
    quizzes, priors, askeds = [], [], []
    for b, depth in enumerate((0, 3, 8, 5, 1, 2, 4, 6, 7, 0, 3)):
        quiz = eng.start_quiz()
        prefix = synth.quiz_prefix(b, min(depth, Q - 1), Q, T, K)
        priors.append(drive_quiz(eng, ora, kb, W, quiz, prefix, check_top=False))
        asked = np.zeros(Q, dtype=bool)
        asked[[q for q, _ in prefix]] = True
        askeds.append(asked)
        quizzes.append(quiz)
    bounds = ora.calc_split(Q, 8 * W)
    for n in (1, 2, 4, 5, 11):                            # up to 4: ids in the launch parameters; beyond: device arrays, grid.y tiles
        ev = eng.eval_questions(quizzes[:n])               # the same kernel, evaluation only
        randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        before = eng.get_total_questions_asked()
        chosen = eng.next_question_batch(quizzes[:n], randoms)
        assert eng.get_total_questions_asked() == before + n
        for x in range(n):
            oev = oracle_eval_all(ora, kb, priors[x], W, askeds[x])
            ok = ~askeds[x]
            assert np.array_equal(np.isnan(ev["priority"][x]), askeds[x])
            rel = np.abs(ev["priority"][x][ok] - oev["priority"][ok]) / np.abs(oev["priority"][ok])
            assert rel.max() <= TOL_STAGED, (n, x, float(rel.max()))
            assert np.allclose(ev["runLength"][x], oev["runLength"], rtol=4 * TOL_STAGED, atol=0)
            assert np.allclose(ev["grand"][x], oev["grand"], rtol=4 * TOL_STAGED, atol=0)
            e1 = dict(runLength=ev["runLength"][x], grand=ev["grand"][x], bounds=bounds)
            assert chosen[x] == ora.select_question(e1, Q, int(randoms[x]), asked=askeds[x]), (n, x)
            assert chosen[x] == ora.select_question(oev, Q, int(randoms[x]), asked=askeds[x]), (n, x)
            assert eng.get_active_question_id(int(quizzes[x])) == chosen[x] and not askeds[x][chosen[x]]
    q1 = eng.next_question(int(quizzes[3]))                  # the one-quiz entry point itself
    assert 0 <= q1 < Q and not askeds[3][q1] and eng.get_active_question_id(int(quizzes[3])) == q1
    # priors are untouched by NextQuestion
    assert np.array_equal(bits(eng.copy_quiz_priors(int(quizzes[0]))), bits(priors[0]))


def test_selection_skips_to_nearest_unasked(pqa, ora):
    Q, K, T, W = 130, 5, 64, 1
    kb = synth.uniform_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    quiz = eng.start_quiz()
    asked = np.zeros(Q, dtype=bool)
    for q in list(range(0, 70)) + list(range(71, 129)):   # leaves 70 and 129
        eng.set_active_question(quiz, q)
        eng.record_answer(quiz, 0)
        asked[q] = True
    for rnd in (0, 2 ** 63, 2 ** 64 - 1):
        got = eng.next_question_batch([quiz], np.array([rnd], dtype=np.uint64))[0]
        assert got in (70, 129)
    # exhaust
    for q in (70, 129):
        eng.set_active_question(quiz, q)
        eng.record_answer(quiz, 1)
    with pytest.raises(pqa.PqaException) as ei:
        eng.next_question(quiz)
    assert "Engine has run out of questions" in str(ei.value)


def test_record_quiz_target_and_train_bit_exact(pqa, ora):
    Q, K, T, W = 50, 5, 120, 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    sA, mD, vB = [a.copy() for a in kb]
    eng = make_engine(pqa, Q, K, T, W, kb)
    rng = np.random.default_rng(3)
    # many quizzes sharing targets and questions => colliding cells, applied in array order
    n = 40
    quizzes = eng.start_quiz_batch(n)
    targets = rng.integers(0, 6, size=n)
    amounts = rng.choice([1.0, 0.5, 2.25], size=n)
    all_aqs = []
    for x, quiz in enumerate(quizzes):
        aqs = []
        for q in rng.choice(12, size=int(rng.integers(0, 9)), replace=False):
            a = int(rng.integers(0, K))
            eng.set_active_question(int(quiz), int(q))
            eng.record_answer(int(quiz), a)
            aqs.append((int(q), a))
        all_aqs.append(aqs)
        ora.record_quiz_target(sA, mD, vB, aqs, int(targets[x]), float(amounts[x]))
    eng.record_quiz_target_batch(quizzes, targets, amounts)
    gA, gD, gB = eng.download_kb()
    assert np.array_equal(bits(gA), bits(sA)) and np.array_equal(bits(gD), bits(mD)) and np.array_equal(bits(gB), bits(vB))
    # single-call form
    eng.record_quiz_target(int(quizzes[3]), 7, 1.5)
    ora.record_quiz_target(sA, mD, vB, all_aqs[3], 7, 1.5)
    # Train(): duplicate questions (same and different answers) inside one call
    aqs = [(3, 1), (7, 2), (3, 1), (11, 0), (7, 4), (15, 2), (3, 2), (19, 1), (23, 3)]
    before = eng.get_total_questions_asked()
    eng.train([pqa.AnsweredQuestion(q, a) for q, a in aqs], 9, 0.75)
    assert eng.get_total_questions_asked() == before + len(aqs)
    ora.train(sA, mD, vB, aqs, 9, 0.75, W)
    gA, gD, gB = eng.download_kb()
    assert np.array_equal(bits(gA), bits(sA)) and np.array_equal(bits(gD), bits(mD)) and np.array_equal(bits(gB), bits(vB))


def test_large_record_quiz_target_batch_device_grouped(pqa, ora):
    """A batch big enough (>= 8192 cell operations) to take the device-grouped path (radix sort by cell, pqa_train_sort.cu)
    with heavy cell collisions: the KB must equal the oracle applying the quizzes one by one, bit for bit."""
    Q, K, T, W = 40, 5, 64, 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    sA, mD, vB = [a.copy() for a in kb]
    eng = make_engine(pqa, Q, K, T, W, kb)
    rng = np.random.default_rng(8)
    n, d = 1800, 6
    quizzes = eng.start_quiz_batch(n)
    qs = np.argsort(rng.random((n, Q)), axis=1)[:, :d]
    ans = rng.integers(0, K, size=(n, d))
    for s_ in range(d):
        eng.set_active_question_batch(quizzes, qs[:, s_])
        eng.record_answer_batch(quizzes, ans[:, s_])
    targets = rng.integers(0, T, size=n)
    amounts = rng.choice([1.0, 0.5, 2.25], size=n)
    eng.record_quiz_target_batch(quizzes, targets, amounts)
    for x in range(n):
        ora.record_quiz_target(sA, mD, vB, list(zip(qs[x].tolist(), ans[x].tolist())), int(targets[x]), float(amounts[x]))
    gA, gD, gB = eng.download_kb()
    assert np.array_equal(bits(gA), bits(sA)) and np.array_equal(bits(gD), bits(mD)) and np.array_equal(bits(gB), bits(vB))


def test_initial_kb_and_dims(pqa):
    """Dimensions.CpuIncrease-style check (PqaCoreTests/Dimensions.cpp:62-77): A = init^2, D = K*init^2, B = init."""
    Q, K, T = 7, 5, 203
    eng = make_engine(pqa, Q, K, T, 2)
    d = eng.copy_dims()
    assert (d.n_answers, d.n_questions, d.n_targets) == (K, Q, T)
    assert np.all(eng.copy_a_targets(3, 2) == INIT * INIT)
    assert np.all(eng.copy_d_targets(6) == K * (INIT * INIT))
    assert np.all(eng.copy_b_targets() == INIT)
    sA, mD, vB = eng.download_kb()
    assert np.all(sA == INIT * INIT) and np.all(mD == K * (INIT * INIT)) and np.all(vB == INIT)


@pytest.mark.parametrize("n", [1, 40])
def test_degenerate_priorities_are_counted_not_fatal(pqa, ora, n):
    """CEEvalQsSubtaskConsider.cpp:209-211 and CpuEngine.cpp:368-377 only WARN about a priority <= 0, a non-finite running
    total and a grand total <= 0, and NextQuestion still answers. Cells of 1e300 make 1/mD^2 underflow: lack = 0, every
    priority = 0 (the oracle agrees), the grand total is 0. The engine counts the three conditions (fused launch and batch
    path) and keeps answering."""
    Q, K, T, W = 24, 3, 16, 2
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=1e150), emulated_workers=W,
                                                    rng_seed=9)
    assert np.array_equal(eng.anomaly_counts(), [0, 0, 0])
    quizzes = eng.start_quiz_batch(n)
    sA = np.full((Q, K, T), 1e300); mD = np.full((Q, T), K * 1e300)
    prior = ora.start_quiz(np.full(T, 1e150), W)
    assert np.all(oracle_eval_all(ora, (sA, mD, None), prior, W)["priority"] == 0.0)
    got = eng.eval_questions(quizzes)
    assert np.all(got["priority"] == 0.0)
    before = eng.anomaly_counts()
    chosen = eng.next_question_batch(quizzes, np.arange(n, dtype=np.uint64) * 977)
    assert np.all((chosen >= 0) & (chosen < Q))
    after = eng.anomaly_counts()
    assert after[0] - before[0] == n * Q and after[1] == before[1] and after[2] - before[2] == n
    # a healthy engine counts nothing
    eng2 = make_engine(pqa, 16, 3, 32, 2)
    q2 = eng2.start_quiz_batch(5)
    eng2.next_question_batch(q2, np.arange(5, dtype=np.uint64))
    assert np.array_equal(eng2.anomaly_counts(), [0, 0, 0])


def test_error_contract(pqa):
    Q, K, T = 8, 3, 16
    eng = make_engine(pqa, Q, K, T, 2)
    quiz = eng.start_quiz()
    with pytest.raises(pqa.PqaException) as ei:
        eng.record_answer(quiz, 0)
    assert "[No active question in the quiz]" in str(ei.value) and "answerId=0" in str(ei.value)
    eng.next_question(quiz)
    with pytest.raises(pqa.PqaException) as ei:
        eng.record_answer(quiz, K)
    assert "[Index is out of range]" in str(ei.value) and "subjIndex=3 not in 0...2" in str(ei.value)
    with pytest.raises(pqa.PqaException) as ei:
        eng.next_question(quiz + 5)
    assert "[Index is out of range]" in str(ei.value)
    with pytest.raises(pqa.PqaException) as ei:
        eng.record_quiz_target(quiz, 3, 0.0)
    assert "[The amount is not positive]" in str(ei.value)
    with pytest.raises(pqa.PqaException) as ei:
        eng.record_quiz_target(quiz, T, 1.0)
    assert "[Index is out of range]" in str(ei.value)
    eng.release_quiz(quiz)
    with pytest.raises(pqa.PqaException) as ei:
        eng.next_question(quiz)
    assert "[The ID is absent from KB]" in str(ei.value)
    assert eng.start_quiz() == quiz          # ids are recycled LIFO (GapTracker.h:38-49)
    with pytest.raises(pqa.PqaException) as ei:
        eng.train([pqa.AnsweredQuestion(0, 0)], 0, -1.0)
    assert "[The amount is not positive]" in str(ei.value)
    with pytest.raises(pqa.PqaException):
        pqa.PqaEngineFactory().create_cpu_engine(pqa.EngineDefinition(1, 1, 1))
    # the same quiz twice in one RecordAnswer launch: the second answer has no active question left (CEQuiz.h:78-92); the
    # whole batch is refused before any quiz is touched
    a, b = eng.start_quiz_batch(2)
    eng.next_question_batch([a, b], np.array([1, 2], dtype=np.uint64))
    before = eng.copy_quiz_priors(int(a))
    with pytest.raises(pqa.PqaException) as ei:
        eng.record_answer_batch([a, b, a], [0, 1, 2])
    assert "[No active question in the quiz]" in str(ei.value)
    assert np.array_equal(bits(eng.copy_quiz_priors(int(a))), bits(before)) and eng.get_active_question_id(int(a)) >= 0
    eng.record_answer_batch([a, b], [0, 1])


def test_shutdown_contract(pqa, tmp_path):
    """PqaEngine_Shutdown (BaseEngine.cpp:260-322): the KB is saved when a path is given, then the engine is gone for every
    later call (ObjectShutDown, MaintenanceSwitch.h:138-141), dimensions read 0, and a second Shutdown says so too."""
    Q, K, T = 12, 3, 40
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, 2, kb)
    quiz = eng.start_quiz()
    eng.next_question(quiz)
    path = str(tmp_path / "down.kb")
    eng.shutdown(path)
    again = pqa.PqaEngineFactory().load_b200_engine(path, emulated_workers=2)
    for g, w in zip(again.download_kb(), kb):
        assert np.array_equal(bits(g), bits(w))
    for call in (lambda: eng.start_quiz(), lambda: eng.next_question(quiz), lambda: eng.record_answer(quiz, 0),
                 lambda: eng.list_top_targets(quiz, 3), lambda: eng.train([pqa.AnsweredQuestion(0, 0)], 1),
                 lambda: eng.save_kb(path), lambda: eng.download_kb(), lambda: eng.shutdown()):
        with pytest.raises(pqa.PqaException) as ei:
            call()
        assert "[Object is shut(ting) down]" in str(ei.value)
    d = eng.copy_dims()
    assert (d.n_answers, d.n_questions, d.n_targets) == (0, 0, 0)


@pytest.mark.parametrize("depth", [0, 3, 8])
def test_full_size_staged_vs_exact_and_oracle(pqa, ora, depth):
    """BASELINE config 2 size (1000x5x1000): the staged kernel against the exact kernel for a batch of quizzes, and
    against the oracle for one quiz (the oracle needs ~1 s per quiz at this size)."""
    Q, K, T, W = 1000, 5, 1000, 8
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = make_engine(pqa, Q, K, T, W, kb)
    n = 16
    quizzes = eng.start_quiz_batch(n)
    for x, quiz in enumerate(quizzes):
        for (q, a) in synth.quiz_prefix(x, depth, Q, T, K):
            eng.set_active_question(int(quiz), q)
            eng.record_answer(int(quiz), a)
    eng.set_eval_kernel(2, kahan_lanes_per_thread=1)
    fast = eng.eval_questions(quizzes)["priority"]
    eng.set_eval_kernel(2, kahan_lanes_per_thread=4)
    fast4 = eng.eval_questions(quizzes)["priority"]
    assert np.array_equal(np.isnan(fast), np.isnan(fast4))
    assert np.allclose(fast[~np.isnan(fast)], fast4[~np.isnan(fast)], rtol=TOL_STAGED, atol=0)
    eng.set_eval_kernel(1)
    exact = eng.eval_questions(quizzes)["priority"]
    assert np.array_equal(np.isnan(fast), np.isnan(exact))
    for x, quiz in enumerate(quizzes):
        prior = eng.copy_quiz_priors(int(quiz))
        tol = staged_tolerance(kb, prior)
        ok = ~np.isnan(exact[x])
        rel = np.abs(fast[x][ok] - exact[x][ok]) / np.abs(exact[x][ok])
        assert np.all(rel <= tol[ok]), (x, float(np.max(rel / tol[ok])))
    prior0 = eng.copy_quiz_priors(int(quizzes[0]))
    asked0 = np.isnan(exact[0])
    oev = oracle_eval_all(ora, kb, prior0, W, asked0)
    ok = ~asked0
    assert np.max(ulp_diff(exact[0][ok], oev["priority"][ok])) <= 8
    # top-10 lists of every quiz are bit-identical to the oracle's
    items, counts = eng.list_top_targets_batch(quizzes, 10)
    for x, quiz in enumerate(quizzes):
        want = ora.list_top_targets(eng.copy_quiz_priors(int(quiz)), W, 10)
        assert [(int(t), float(p)) for t, p in items[x][:counts[x]]] == want


def bench_batch(eng, Q, K, T, B, depths=(0, 3, 8)):
    """The batch bench.py times (bench.py quiz_states): quiz b has depth depths[b % 3] and synth.quiz_prefix(b, depth)."""
    quizzes = eng.start_quiz_batch(B)
    states = [synth.quiz_prefix(b, min(depths[b % len(depths)], Q - 1), Q, T, K) for b in range(B)]
    for s in range(max(len(pf) for pf in states)):
        sel = [x for x in range(B) if len(states[x]) > s]
        eng.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
        eng.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
    return quizzes, states


@pytest.mark.parametrize("dims,B,W,sample", [
    ((1000, 5, 1000), 256, 8, (0, 1, 2, 127, 128, 129, 254, 255)),   # BASELINE config 2 as bench.py runs it: two CTA tiles of 128 quizzes
    ((300, 5, 999), 200, 8, (0, 64, 130, 199)),                      # ragged T (padding lanes), second tile partly filled
    ((200, 3, 1400), 65, 4, (0, 31, 32, 64)),                        # smallest batch that takes the 8-warp / two-threads-per-quiz shape
    # medium batches (k_eval_multi: 128 / G questions per CTA, slabs streamed in chunks)
    ((1000, 5, 1000), 40, 8, (0, 1, 2, 39)),                         # G = 64, two questions per CTA
    ((61, 5, 203), 33, 3, (0, 16, 32)),                              # odd question count: the last CTA holds one question; ragged T
    ((45, 5, 999), 20, 3, (0, 10, 19)),                              # G = 32, four questions per CTA, last CTA holds one
    ((130, 4, 600), 64, 4, (0, 33, 63)),                             # every quiz place of G = 64 taken
    ((90, 2, 2100), 17, 2, (0, 8, 16)),                              # two answers, several chunks per question
    ((77, 8, 300), 12, 2, (0, 5, 11)),                               # four threads per quiz, G = 16, four questions per CTA, K = 8
    ((50, 5, 1000), 9, 8, (0, 4, 8)),                                # smallest batch of the medium kernel
])
def test_benched_instantiation_vs_oracle(pqa, ora, dims, B, W, sample):
    """The instantiation behind every BENCH / SCALE number -- default tuning, batch > 64, whole slab in shared memory
    (k_eval_staged<K,2,8>) -- against the oracle (CEEvalQsSubtaskConsider.cpp:41-217): priorities 2e-12 flat, W_k bit for
    bit, H_k / V_k / lack 1e-12, run-lengths and the selected question for the same 64-bit draw."""
    Q, K, T = dims
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = make_engine(pqa, Q, K, T, W, kb)
    eng.set_eval_kernel(0)
    quizzes, states = bench_batch(eng, Q, K, T, B)
    det = eng.eval_questions_detailed_batch(quizzes)
    ev = eng.eval_questions(quizzes)
    assert np.array_equal(bits(det["priority"]), bits(ev["priority"]))   # the same kernel, run twice: deterministic
    rng = np.random.default_rng(5)
    randoms = rng.integers(0, 2 ** 64, size=B, dtype=np.uint64)
    chosen = eng.next_question_batch(quizzes, randoms)
    bounds = ora.calc_split(Q, 8 * W)
    for x in range(B):      # cheap checks on every quiz of the batch
        asked = np.zeros(Q, dtype=bool)
        asked[[q for q, _ in states[x]]] = True
        assert np.array_equal(np.isnan(ev["priority"][x]), asked), x
        e1 = dict(runLength=ev["runLength"][x], grand=ev["grand"][x], bounds=bounds)
        assert chosen[x] == ora.select_question(e1, Q, int(randoms[x]), asked=asked), x
    for x in sample:        # the oracle on a sample covering every depth and both quiz tiles
        prior = eng.copy_quiz_priors(int(quizzes[x]))
        asked = np.isnan(ev["priority"][x])
        ok = ~asked
        oev = oracle_eval_all(ora, kb, prior, W, asked)
        rel = np.abs(ev["priority"][x][ok] - oev["priority"][ok]) / np.abs(oev["priority"][ok])
        assert rel.max() <= TOL_STAGED, (x, float(rel.max()))
        assert np.allclose(ev["runLength"][x], oev["runLength"], rtol=4 * TOL_STAGED, atol=0)
        assert np.allclose(ev["grand"][x], oev["grand"], rtol=4 * TOL_STAGED, atol=0)
        for i in np.flatnonzero(ok)[::max(1, Q // 120)]:
            o = ora.eval_question(kb[0][i], kb[1][i], prior)
            assert np.array_equal(bits(det["W"][x, i]), bits(o["W"])), (x, i)
            assert np.allclose(det["H"][x, i], o["H"], rtol=1e-12, atol=0)
            assert np.allclose(det["V"][x, i], o["V"], rtol=1e-12, atol=0)
            assert abs(det["lack"][x, i] - o["lack"]) <= 1e-12 * abs(o["lack"])


def test_kb_file_roundtrip_in_reference_layout(pqa, tmp_path):
    """SaveKB / LoadCpuEngine in the reference's byte layout (BaseEngine.cpp:323-385, CpuEngine.cpp:664-688,
    PermanentIdManager.cpp:27-39): header, sA rows, mD rows, vB, empty gap lists, identity id maps."""
    Q, K, T = 9, 5, 203     # T not a multiple of 4: device rows are padded, file rows are not
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, 3, kb)
    quiz = eng.start_quiz()
    eng.next_question(quiz)
    eng.record_answer(quiz, 1)
    eng.record_quiz_target(quiz, 17)
    sA, mD, vB = eng.download_kb()
    path = str(tmp_path / "kb.bin")
    eng.save_kb(path)
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:40], dtype=np.uint64)
    assert hdr[0] & 0xF == 3 and (hdr[0] >> 4) & 0xFFFFFFF == 53 and (hdr[0] >> 32) & 0xFFFF == 11
    assert tuple(np.frombuffer(raw[8:32], dtype=np.int64)) == (K, Q, T) and hdr[4] == eng.get_total_questions_asked() == 1
    off = 40
    fA = np.frombuffer(raw, dtype=np.float64, count=Q * K * T, offset=off).reshape(Q, K, T); off += fA.nbytes
    fD = np.frombuffer(raw, dtype=np.float64, count=Q * T, offset=off).reshape(Q, T); off += fD.nbytes
    fB = np.frombuffer(raw, dtype=np.float64, count=T, offset=off); off += fB.nbytes
    assert np.array_equal(bits(fA), bits(sA)) and np.array_equal(bits(fD), bits(mD)) and np.array_equal(bits(fB), bits(vB))
    tail = np.frombuffer(raw, dtype=np.int64, offset=off)
    want_tail = [0, 0, Q, Q] + list(range(Q)) + [T, T] + list(range(T)) + [1, 0]
    assert tail.tolist() == want_tail
    eng2, err = pqa.PqaEngineFactory().load_cpu_engine(path)
    assert err is None and eng2.get_total_questions_asked() == 1
    d = eng2.copy_dims()
    assert (d.n_answers, d.n_questions, d.n_targets) == (K, Q, T)
    lA, lD, lB = eng2.download_kb()
    assert np.array_equal(bits(lA), bits(sA)) and np.array_equal(bits(lD), bits(mD)) and np.array_equal(bits(lB), bits(vB))
    with pytest.raises(pqa.PqaException) as ei:
        pqa.PqaEngineFactory().load_cpu_engine(str(tmp_path / "absent.bin"))
    assert "[Cannot open file]" in str(ei.value)


def test_concurrent_one_quiz_calls_are_combined_correctly(pqa, ora):
    """Many client threads on the reference's one-quiz-per-call ABI at once (like PqaClient.cpp:238-245): the engine
    combines whatever calls are pending into one batch launch. Every quiz must still see exactly its own posterior
    (bit-exact against the oracle for its own answer sequence) and its own errors."""
    import threading
    Q, K, T, W = 60, 5, 300, 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    eng = make_engine(pqa, Q, K, T, W, kb)
    n_threads, n_quizzes, n_steps = 12, 4, 5
    failures, lock = [], threading.Lock()

    def client(tid):
        try:
            for z in range(n_quizzes):
                quiz = eng.start_quiz()
                prior = ora.start_quiz(kb[2], W)
                for step in range(n_steps):
                    q = eng.next_question(quiz)
                    assert 0 <= q < Q and eng.get_active_question_id(quiz) == q
                    a = (q + tid + step) % K
                    eng.record_answer(quiz, a)
                    prior = ora.record_answer(prior, kb[0][q, a], kb[1][q], max(1, W - 1))
                    assert np.array_equal(bits(eng.copy_quiz_priors(quiz)), bits(prior))
                    top = [(r.i_target, r.prob) for r in eng.list_top_targets(quiz, 3 + tid % 4)]
                    assert top == ora.list_top_targets(prior, W, 3 + tid % 4)
                if tid % 3 == 0:   # an invalid call from this thread must not disturb the others
                    with pytest.raises(pqa.PqaException) as ei:
                        eng.record_answer(quiz, 0)
                    assert "[No active question in the quiz]" in str(ei.value)
                    with pytest.raises(pqa.PqaException):
                        eng.next_question(10 ** 6 + tid)
                eng.release_quiz(quiz)
        except BaseException as ex:   # noqa: BLE001 - reported to the main thread
            with lock:
                failures.append((tid, repr(ex)))

    threads = [threading.Thread(target=client, args=(t,)) for t in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not failures, failures[:3]
    assert eng.get_total_questions_asked() == n_threads * n_quizzes * n_steps


@pytest.mark.parametrize("dims,W", [((30, 5, 64), 4), ((25, 5, 203), 3), ((300, 4, 50), 1), ((40, 5, 1000), 8)])
def test_resume_quiz_bit_exact(pqa, ora, dims, W):
    """ResumeQuiz: priors bit-exact against the oracle (itself pinned on the reference's own
    CEUpdatePriorsSubtaskMul / NormalizePriors code), asked bits and answers installed, then the quiz continues normally."""
    Q, K, T = dims
    rng = np.random.default_rng(5)
    sA, mD, vB = synth.gamma_kb(Q, K, T, INIT)
    vB = vB + rng.uniform(0, 3, size=T)
    eng = make_engine(pqa, Q, K, T, W, (sA, mD, vB))
    lists = []
    for n in (1, 2, 7, 0, min(Q - 1, 250)):
        qs = rng.choice(Q, size=n, replace=False)
        lists.append([(int(q), int(rng.integers(0, K))) for q in qs])
    ids = eng.resume_quiz_batch(lists)
    for quiz, aqs in zip(ids, lists):
        want = ora.resume_quiz(sA, mD, vB, aqs, W) if aqs else ora.start_quiz(vB, W)
        assert np.array_equal(bits(eng.copy_quiz_priors(int(quiz))), bits(want)), len(aqs)
        pri = eng.eval_questions([int(quiz)])["priority"][0]
        assert sorted(np.nonzero(np.isnan(pri))[0].tolist()) == sorted(q for q, _ in aqs)   # asked bits
        assert [(r.i_target, r.prob) for r in eng.list_top_targets(int(quiz), 10)] == ora.list_top_targets(want, W, 10)
    # the single-call form, then the quiz goes on: NextQuestion never returns an answered question
    quiz = eng.resume_quiz([pqa.AnsweredQuestion(q, a) for q, a in lists[2]])
    assert np.array_equal(bits(eng.copy_quiz_priors(quiz)), bits(ora.resume_quiz(sA, mD, vB, lists[2], W)))
    nxt = eng.next_question(quiz)
    assert nxt not in [q for q, _ in lists[2]]
    eng.record_answer(quiz, 0)
    eng.record_quiz_target(quiz, 1)     # trains with the resumed answers + the new one
    with pytest.raises(pqa.PqaException) as ei:
        eng.resume_quiz([pqa.AnsweredQuestion(Q, 0)])
    assert "[Aggregate error]" in str(ei.value) and "[Index is out of range]" in str(ei.value)
