"""GPU test of the question-sharded engines: several shard engines in one process on one GPU, driven by the same
orchestrator that torchrun ranks use (probqa_b200/sharded.py), against a single un-sharded engine: all-reduced
priorities, chosen questions, posteriors, top-10 lists and the trained KB must be bit-identical."""
import numpy as np
import pytest

from probqa_b200 import sharded, synth

pytestmark = pytest.mark.gpu
INIT = 0.1


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("dims,n_shards,lanes", [((40, 5, 203), 2, 1), ((64, 5, 1000), 3, 2), ((37, 4, 96), 4, 4)])
def test_sharded_engines_match_single_engine(dims, n_shards, lanes):
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    full = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    full.upload_kb(*kb)
    full.set_eval_kernel(2, kahan_lanes_per_thread=lanes)
    shards = []
    for first, count in sharded.shard_ranges(Q, n_shards):
        e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5, question_shard_first=first, question_shard_count=count)
        e.upload_kb(*kb)
        e.set_eval_kernel(2, kahan_lanes_per_thread=lanes)
        assert e.question_shard() == (first, count)
        shards.append(sharded.B200Shard(e))
    eng = sharded.QuestionShardedEngine(shards)

    n = 9
    ids = eng.start_quiz_batch(n)
    ids_full = full.start_quiz_batch(n)
    assert np.array_equal(ids, ids_full)
    rng = np.random.default_rng(77)
    for step in range(4):
        want_pri = full.eval_questions(ids_full)["priority"]
        got_pri = eng.eval_priorities(ids)
        assert np.array_equal(np.isnan(got_pri), np.isnan(want_pri))
        ok = ~np.isnan(want_pri)
        assert np.array_equal(bits(got_pri[ok]), bits(want_pri[ok])), "all-reduced priorities differ from the single engine"
        randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        chosen = eng.next_question_batch(ids, randoms)
        assert np.array_equal(chosen, full.next_question_batch(ids_full, randoms))
        answers = [(int(c) * 7 + step) % K for c in chosen]
        eng.record_answer_batch(ids, answers)
        full.record_answer_batch(ids_full, answers)
        for s in shards:   # every shard holds the same posterior bits as the single engine
            for q in ids:
                assert np.array_equal(bits(s.copy_quiz_priors(int(q))), bits(full.copy_quiz_priors(int(q))))
        items, counts = eng.list_top_targets_batch(ids, 10)
        items_f, counts_f = full.list_top_targets_batch(ids_full, 10)
        assert np.array_equal(counts, counts_f) and items.tobytes() == items_f.tobytes()
    # training: every shard updates the cells of its own questions, vB everywhere
    targets = rng.integers(0, T, size=n)
    eng.record_quiz_target_batch(ids, targets)
    full.record_quiz_target_batch(ids_full, targets)
    aqs = [pqa.AnsweredQuestion(int(q), int(a)) for q, a in zip(rng.integers(0, Q, 12), rng.integers(0, K, 12))]
    eng.train(aqs, 3, 0.5)
    full.train(aqs, 3, 0.5)
    wA, wD, wB = full.download_kb()
    gA, gD = np.full_like(wA, np.nan), np.full_like(wD, np.nan)
    for s in shards:
        a, d, b = s.engine.download_kb()    # fills this shard's rows only
        f, c = s.first, s.count
        gA[f:f + c], gD[f:f + c] = a[f:f + c], d[f:f + c]
        assert np.array_equal(bits(b), bits(wB))
    assert np.array_equal(bits(gA), bits(wA)) and np.array_equal(bits(gD), bits(wD))
    # a sharded engine refuses the single-engine calls instead of answering from a partial KB
    with pytest.raises(pqa.PqaException):
        shards[0].engine.next_question(int(ids[0]))
