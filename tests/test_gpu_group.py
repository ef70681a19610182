"""A sharded engine group (PqaB200_CreateShardedEngine): N shard engines of one process behind ONE engine handle that
serves the reference's C ABI. Run here with all shards on one GPU (on a multi-GPU box pass devices=[0, 1, ...]); the
results must be those of a single engine."""
import json
import os
import subprocess
import threading

import numpy as np
import pytest

from probqa_b200 import synth

pytestmark = pytest.mark.gpu
INIT = 0.1


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def devices(n):
    import torch
    have = torch.cuda.device_count()
    return [r % have for r in range(n)]


@pytest.mark.parametrize("axis,dims,n_shards,exact", [("questions", (40, 5, 203), 3, False), ("targets", (36, 5, 1000), 2, True),
                                                      ("targets", (25, 4, 96), 4, False)])
def test_group_behaves_like_one_engine(axis, dims, n_shards, exact, tmp_path):
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    one = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    grp = fac.create_sharded_engine(edef, axis, n_shards, devices=devices(n_shards), exact_order=exact, max_batch=16,
                                    emulated_workers=W, rng_seed=5)
    assert grp.shard_count() == n_shards and one.shard_count() == 1
    one.upload_kb(*kb); grp.upload_kb(*kb)
    n = 37                                                   # more than max_batch: the group cuts the batch into slices
    ids = one.start_quiz_batch(n)
    assert np.array_equal(ids, grp.start_quiz_batch(n))
    rng = np.random.default_rng(82)
    for step in range(3):
        randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        c_one = one.next_question_batch(ids, randoms)
        c_grp = grp.next_question_batch(ids, randoms)
        if axis == "questions":
            assert np.array_equal(c_one, c_grp)             # bit-identical priorities -> identical choices
        else:
            one.set_active_question_batch(ids, c_grp)       # target shards: tolerance-level priorities, follow the group
        assert grp.get_active_question_id(int(ids[3])) == c_grp[3]
        answers = [(int(c) * 7 + step) % K for c in c_grp]
        one.record_answer_batch(ids, answers)
        grp.record_answer_batch(ids, answers)
        for q in ids[:6]:
            assert np.array_equal(bits(grp.copy_quiz_priors(int(q))), bits(one.copy_quiz_priors(int(q))))
        a, ca = grp.list_top_targets_batch(ids, 10)
        b, cb = one.list_top_targets_batch(ids, 10)
        assert np.array_equal(ca, cb) and a.tobytes() == b.tobytes()
    # the reference's one-quiz entry points on the group handle
    q1 = grp.start_quiz(); q1b = one.start_quiz()
    assert q1 == q1b
    nq = grp.next_question(q1)
    one.next_question(q1b)                                   # (counts as a question asked on both sides)
    one.set_active_question(q1b, nq)
    grp.record_answer(q1, 2); one.record_answer(q1b, 2)
    assert [(r.i_target, r.prob) for r in grp.list_top_targets(q1, 5)] == [(r.i_target, r.prob) for r in one.list_top_targets(q1b, 5)]
    grp.record_quiz_target(q1, 7, 1.5); one.record_quiz_target(q1b, 7, 1.5)
    grp.release_quiz(q1); one.release_quiz(q1b)
    with pytest.raises(pqa.PqaException):
        grp.next_question(q1)
    # training and the KB
    targets = rng.integers(0, T, size=n)
    grp.record_quiz_target_batch(ids, targets); one.record_quiz_target_batch(ids, targets)
    aqs = [pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(3, 0), pqa.AnsweredQuestion(1, 1)]
    grp.train(aqs, 5, 0.75); one.train(aqs, 5, 0.75)
    assert grp.train([pqa.AnsweredQuestion(Q, 0)], 1, 1.0, throw=False) is not None
    for g, w in zip(grp.download_kb(), one.download_kb()):
        assert np.array_equal(bits(g), bits(w))
    assert np.array_equal(bits(grp.copy_a_targets(2, 1)), bits(one.copy_a_targets(2, 1)))
    assert np.array_equal(bits(grp.copy_d_targets(2)), bits(one.copy_d_targets(2)))
    assert grp.get_total_questions_asked() == one.get_total_questions_asked()
    pg, po = str(tmp_path / "g.kb"), str(tmp_path / "o.kb")
    grp.save_kb(pg); one.save_kb(po)
    assert open(pg, "rb").read() == open(po, "rb").read()
    again = fac.load_sharded_engine(pg, axis, n_shards, devices=devices(n_shards), emulated_workers=W)
    for g, w in zip(again.download_kb(), one.download_kb()):
        assert np.array_equal(bits(g), bits(w))
    quiz = again.start_quiz()
    assert 0 <= again.next_question(quiz) < Q


@pytest.mark.parametrize("axis,dims,n_shards", [("questions", (40, 5, 203), 3), ("targets", (36, 5, 1000), 2), ("targets", (25, 4, 96), 4)])
def test_group_resume_quiz_and_clear_old_quizzes(axis, dims, n_shards):
    """ResumeQuiz (CpuEngine.cpp:277-282) and ClearOldQuizzes (BaseEngine.cpp:814-872) on a group handle: priors after a
    resume are bit-identical to one engine's (and hence to the oracle's, test_resume_quiz_bit_exact), on every shard; the
    resumed quizzes then continue like any other; ClearOldQuizzes releases the same quizzes on both sides."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    one = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    grp = fac.create_sharded_engine(edef, axis, n_shards, devices=devices(n_shards), exact_order=True, max_batch=16,
                                    emulated_workers=W, rng_seed=5)
    one.upload_kb(*kb); grp.upload_kb(*kb)
    rng = np.random.default_rng(83)
    lists = []
    for n_ans in (0, 1, 7, min(Q - 2, 30), 3, 0, 12):
        qs = rng.permutation(Q)[:n_ans]
        lists.append([pqa.AnsweredQuestion(int(q), int(rng.integers(0, K))) for q in qs])
    ids_one = one.resume_quiz_batch(lists)
    ids_grp = grp.resume_quiz_batch(lists)
    assert np.array_equal(ids_one, ids_grp) and np.all(ids_grp >= 0)
    for q in ids_grp:
        assert np.array_equal(bits(grp.copy_quiz_priors(int(q))), bits(one.copy_quiz_priors(int(q))))
    single = grp.resume_quiz(lists[2]); single_one = one.resume_quiz(lists[2])
    assert single == single_one
    assert np.array_equal(bits(grp.copy_quiz_priors(single)), bits(one.copy_quiz_priors(single_one)))
    # the resumed quizzes go on: the asked bits were installed on every shard (an asked question is never chosen again)
    randoms = rng.integers(0, 2 ** 64, size=len(ids_grp), dtype=np.uint64)
    chosen = grp.next_question_batch(ids_grp, randoms)
    for x, c in enumerate(chosen):
        assert c not in {aq.i_question for aq in lists[x]}
    one.set_active_question_batch(ids_one, chosen)
    answers = [int(c) % K for c in chosen]
    grp.record_answer_batch(ids_grp, answers); one.record_answer_batch(ids_one, answers)
    for q in ids_grp:
        assert np.array_equal(bits(grp.copy_quiz_priors(int(q))), bits(one.copy_quiz_priors(int(q))))
    a, ca = grp.list_top_targets_batch(ids_grp, 10)
    b, cb = one.list_top_targets_batch(ids_one, 10)
    assert np.array_equal(ca, cb) and a.tobytes() == b.tobytes()
    # ClearOldQuizzes: keep at most 3 quizzes (all have the same age here: the heap order decides, identically on both sides)
    grp.clear_old_quizzes(3, 3600.0); one.clear_old_quizzes(3, 3600.0)
    def alive(eng, q):
        try:
            eng.get_active_question_id(int(q))
            return True
        except pqa.PqaException:
            return False
    alive_g = [int(q) for q in list(ids_grp) + [single] if alive(grp, q)]
    alive_o = [int(q) for q in list(ids_one) + [single_one] if alive(one, q)]
    assert alive_g == alive_o and len(alive_g) == 3
    assert grp.start_quiz() == one.start_quiz()            # released ids are reused in the same (LIFO) order
    grp.clear_old_quizzes(0, -1.0); one.clear_old_quizzes(0, -1.0)
    assert grp.start_quiz() == one.start_quiz()


@pytest.mark.parametrize("axis,dims,n_shards", [("questions", (12, 4, 50), 3), ("targets", (10, 5, 203), 2), ("targets", (9, 3, 64), 4)])
def test_group_maintenance_mode(axis, dims, n_shards, tmp_path):
    """Maintenance mode on a group handle (BaseEngine.cpp:640-779): the KB is gathered into one engine, resized / trimmed /
    compacted there and split over new shards at FinishMaintenance. Every step against one un-sharded engine doing the same:
    dimensions, id maps, the KB bit for bit, the mode contract; afterwards quizzes run on the re-sharded KB with posteriors
    bit-identical to the single engine's."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 4
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    one = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    grp = fac.create_sharded_engine(edef, axis, n_shards, devices=devices(n_shards), exact_order=True, emulated_workers=W, rng_seed=5)
    one.upload_kb(*kb); grp.upload_kb(*kb)
    quiz = grp.start_quiz(); one.start_quiz()
    assert grp.add_qs_ts([1.0], [], throw=False) is not None                       # maintenance-only call in regular mode
    assert grp.start_maintenance(False, throw=False) is not None                   # QuizzesActive unless forced (BaseEngine.cpp:646-655)
    grp.start_maintenance(True); one.start_maintenance(True)                       # destroys the quizzes
    with pytest.raises(pqa.PqaException):
        grp.start_quiz()
    def same_kb():
        for g, w in zip(grp.download_kb(), one.download_kb()):
            assert np.array_equal(bits(g), bits(w))
    same_kb()
    # grow: 3 questions, 5 targets; then remove some of each; ids and cells must follow the single engine
    for e in (grp, one):
        e.add_qs_ts([0.5, 1.5, 2.5], [0.25, 0.5, 0.75, 1.0, 1.25])
    dg, do = grp.copy_dims(), one.copy_dims()
    assert (dg.n_questions, dg.n_targets) == (do.n_questions, do.n_targets) == (Q + 3, T + 5)
    grp._refresh_dims(); one._refresh_dims()
    same_kb()
    for e in (grp, one):
        e.remove_questions([1, Q + 1]); e.remove_targets([0, 7, T + 2])
    assert np.array_equal(grp.question_perm_from_comp(np.arange(Q + 3)), one.question_perm_from_comp(np.arange(Q + 3)))
    assert np.array_equal(grp.target_perm_from_comp(np.arange(T + 5)), one.target_perm_from_comp(np.arange(T + 5)))
    assert grp.finish_maintenance(throw=False) is not None                         # shards hold no gaps: compact first
    cg, co = grp.compact(), one.compact()
    assert np.array_equal(cg[0], co[0]) and np.array_equal(cg[1], co[1])
    grp._refresh_dims(); one._refresh_dims()
    same_kb()
    p = str(tmp_path / "maint.kb")
    grp.save_kb(p)                                                                  # still in maintenance: the gathered KB is saved
    grp.finish_maintenance(); one.finish_maintenance()
    dg = grp.copy_dims()
    assert (dg.n_questions, dg.n_targets) == (Q + 1, T + 2) and grp.shard_count() == n_shards
    same_kb()
    p1 = str(tmp_path / "one.kb")
    one.save_kb(p1)
    assert open(p, "rb").read() == open(p1, "rb").read()
    assert np.array_equal(grp.target_comp_from_perm(np.arange(T + 5)), one.target_comp_from_perm(np.arange(T + 5)))
    # the re-sharded KB serves quizzes like the single engine
    ids_g, ids_o = grp.start_quiz_batch(6), one.start_quiz_batch(6)
    assert np.array_equal(ids_g, ids_o)
    rng = np.random.default_rng(3)
    for step in range(2):
        randoms = rng.integers(0, 2 ** 64, size=6, dtype=np.uint64)
        chosen = grp.next_question_batch(ids_g, randoms)
        one.set_active_question_batch(ids_o, chosen)
        answers = [int(c) % K for c in chosen]
        grp.record_answer_batch(ids_g, answers); one.record_answer_batch(ids_o, answers)
        for q in ids_g:
            assert np.array_equal(bits(grp.copy_quiz_priors(int(q))), bits(one.copy_quiz_priors(int(q))))


def test_group_serves_concurrent_one_quiz_clients():
    """Client threads call the reference's one-quiz entry points on a group handle; the shell's combiner turns them into
    exchanged batch launches on all shards. Every quiz must end with the posterior a single engine computes for the same
    (question, answer) sequence."""
    from probqa_b200 import engine as pqa
    Q, K, T, W = 48, 5, 400, 4
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    grp = fac.create_sharded_engine(edef, "targets", 2, devices=devices(2), exact_order=True, emulated_workers=W, rng_seed=9)
    one = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=9)
    grp.upload_kb(*kb); one.upload_kb(*kb)
    logs, errors = {}, []

    def client(x):
        try:
            quiz = grp.start_quiz()
            seq = []
            for step in range(5):
                q = grp.next_question(quiz)
                a = (q + x + step) % K
                grp.record_answer(quiz, a)
                seq.append((q, a))
                grp.list_top_targets(quiz, 3)
            logs[x] = (quiz, seq)
        except Exception as ex:          # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=client, args=(x,)) for x in range(12)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for x, (quiz, seq) in logs.items():
        ref_quiz = one.start_quiz()
        for q, a in seq:
            one.set_active_question(ref_quiz, q)
            one.record_answer(ref_quiz, a)
        assert np.array_equal(bits(grp.copy_quiz_priors(quiz)), bits(one.copy_quiz_priors(ref_quiz)))


def test_reference_client_loop_on_a_group(tmp_path):
    """The reference's own client program, unmodified, on a two-shard group: PQA_B200_SHARDS makes the reference factory
    call hand out the group."""
    from probqa_b200 import build
    exe = build.build_client()
    env = dict(os.environ, PQA_B200_SHARDS="2", PQA_B200_SHARD_AXIS="targets", PQA_B200_SHARD_DEVICES=",".join(map(str, devices(2))),
               PQA_B200_SHARD_EXACT="1")
    out = subprocess.run([exe, "--trainings", "3000", "--learners", "32", "--report-every", "1024", "--progress", str(tmp_path / "p.txt")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["failed"] is False and line["questions_asked"] > 3000
    print("pqa_client on a 2-shard group:", json.dumps(line))
