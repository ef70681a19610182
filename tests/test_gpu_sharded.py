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


@pytest.mark.parametrize("dims,n_shards,chunk,n", [((40, 5, 203), 2, 0, 9), ((64, 5, 1000), 3, 64, 70), ((37, 4, 96), 4, 0, 5),
                                                   ((24, 5, 2050), 2, 0, 130)])
def test_target_sharded_engines_match_single_engine(dims, n_shards, chunk, n):
    """Target shards (BASELINE config 4's axis) on one GPU against one un-sharded engine. Posteriors, top-10 lists and the
    trained KB are bit-identical; priorities are tolerance-level (the sums over shards are not CpuEngine's single Kahan
    pass): W_k 1e-14, priority 1e-10 relative on this KB, and every shard holds the same priority bits."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    full = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    full.upload_kb(*kb)
    shards = []
    ranges = sharded.target_shard_ranges(T, n_shards)
    assert sum(c for _, c in ranges) == T and all(f % 4 == 0 for f, _ in ranges)
    for first, count in ranges:
        e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5, target_shard_first=first, target_shard_count=count)
        e.upload_kb(*kb)
        e.set_eval_kernel(2, chunk_targets=chunk)
        assert e.target_shard() == (first, count)
        shards.append(sharded.B200TargetShard(e))
    eng = sharded.TargetShardedEngine(shards)

    ids = eng.start_quiz_batch(n)
    ids_full = full.start_quiz_batch(n)
    assert np.array_equal(ids, ids_full)
    rng = np.random.default_rng(78)
    for step in range(3):
        want_pri = full.eval_questions(ids_full)["priority"]
        got_pri = eng.eval_priorities(ids)      # also asserts that all shards hold identical bits
        assert np.array_equal(np.isnan(got_pri), np.isnan(want_pri))
        ok = ~np.isnan(want_pri)
        rel = np.abs(got_pri[ok] - want_pri[ok]) / np.abs(want_pri[ok])
        assert rel.max() < 1e-10, rel.max()
        # the all-reduced normalisers against the single engine's bit-exact W_k
        det = full.eval_questions_detailed(int(ids_full[0]))
        w_sh = shards[0]._view(2).cpu().numpy().reshape(n, Q, K)[0]
        okw = ~np.isnan(det["W"])
        assert np.max(np.abs(w_sh[okw] - det["W"][okw]) / det["W"][okw]) < 1e-14
        randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        chosen = eng.next_question_batch(ids, randoms)
        assert np.all((chosen >= 0) & (chosen < Q))
        full.set_active_question_batch(ids_full, chosen)     # same state on both sides whatever the draw resolved to
        answers = [(int(c) * 7 + step) % K for c in chosen]
        eng.record_answer_batch(ids, answers)
        full.record_answer_batch(ids_full, answers)
        for s in shards:
            for q in ids[:6]:
                assert np.array_equal(bits(s.copy_quiz_priors(int(q))), bits(full.copy_quiz_priors(int(q))))
        items, counts = eng.list_top_targets_batch(ids, 10)
        items_f, counts_f = full.list_top_targets_batch(ids_full, 10)
        assert np.array_equal(counts, counts_f) and items.tobytes() == items_f.tobytes()
    targets = rng.integers(0, T, size=n)
    eng.record_quiz_target_batch(ids, targets)
    full.record_quiz_target_batch(ids_full, targets)
    aqs = [pqa.AnsweredQuestion(int(q), int(a)) for q, a in zip(rng.integers(0, Q, 12), rng.integers(0, K, 12))]
    eng.train(aqs, 3, 0.5)
    full.train(aqs, 3, 0.5)
    wA, wD, wB = full.download_kb()
    gA, gD = np.full_like(wA, np.nan), np.full_like(wD, np.nan)
    for s in shards:
        a, d, b = s.engine.download_kb()    # fills this shard's columns only
        f, c = s.first, s.count
        gA[:, :, f:f + c], gD[:, f:f + c] = a[:, :, f:f + c], d[:, f:f + c]
        assert np.array_equal(bits(b), bits(wB))
        assert np.array_equal(bits(s.engine.copy_a_targets(1, 2)[f:f + c]), bits(wA[1, 2, f:f + c]))
        assert np.array_equal(bits(s.engine.copy_d_targets(1)[f:f + c]), bits(wD[1, f:f + c]))
    assert np.array_equal(bits(gA), bits(wA)) and np.array_equal(bits(gD), bits(wD))
    with pytest.raises(pqa.PqaException):
        shards[0].engine.next_question(int(ids[0]))


@pytest.mark.parametrize("dims", [(50, 5, 1000), (33, 3, 250), (20, 5, 4100)])
def test_device_filled_binary_search_kb_is_bit_identical(dims):
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    want = synth.binary_search_kb(Q, K, T, INIT, 3)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    full = fac.create_b200_engine(edef, emulated_workers=4)
    full.fill_binary_search_kb(3)
    for g, w in zip(full.download_kb(), want):
        assert np.array_equal(bits(g), bits(w))
    gA, gD = np.full_like(want[0], np.nan), np.full_like(want[1], np.nan)
    for f, c in sharded.target_shard_ranges(T, 3):
        e = fac.create_b200_engine(edef, emulated_workers=4, target_shard_first=f, target_shard_count=c)
        e.fill_binary_search_kb(3)
        a, d, b = e.download_kb()
        gA[:, :, f:f + c], gD[:, f:f + c] = a[:, :, f:f + c], d[:, f:f + c]
        assert np.array_equal(bits(b), bits(want[2]))
    assert np.array_equal(bits(gA), bits(want[0])) and np.array_equal(bits(gD), bits(want[1]))
    f, c = sharded.shard_ranges(Q, 2)[1]
    e = fac.create_b200_engine(edef, emulated_workers=4, question_shard_first=f, question_shard_count=c)
    e.fill_binary_search_kb(3)
    a, d, b = e.download_kb()
    assert np.array_equal(bits(a[f:f + c]), bits(want[0][f:f + c])) and np.array_equal(bits(d[f:f + c]), bits(want[1][f:f + c]))


@pytest.mark.parametrize("axis,dims,n_shards,n", [("questions", (40, 5, 203), 2, 9), ("questions", (64, 5, 1000), 3, 40),
                                                  ("targets", (40, 5, 203), 2, 9), ("targets", (30, 5, 1000), 4, 130)])
def test_peer_memory_exchange_matches_caller_side_exchange(axis, dims, n_shards, n):
    """The same sharded session twice: shards exchanging through the caller (sum of the Shard*/TShard* buffers) and
    shards exchanging over peer memory from their kernels' epilogues with the device-side barrier (P2P entry points;
    several engines of one process on one GPU, so "peer memory" is ordinary device memory here). Chosen questions,
    priorities, posteriors and top-10 lists must agree bit for bit."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)

    def make():
        shards = []
        if axis == "questions":
            for first, count in sharded.shard_ranges(Q, n_shards):
                e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5, question_shard_first=first, question_shard_count=count)
                e.upload_kb(*kb)
                shards.append(sharded.B200Shard(e))
            return sharded.QuestionShardedEngine(shards)
        for first, count in sharded.target_shard_ranges(T, n_shards):
            e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5, target_shard_first=first, target_shard_count=count)
            e.upload_kb(*kb)
            shards.append(sharded.B200TargetShard(e))
        return sharded.TargetShardedEngine(shards)

    host, p2p = make(), make()
    p2p.enable_p2p(n)
    ids = host.start_quiz_batch(n)
    assert np.array_equal(ids, p2p.start_quiz_batch(n))
    rng = np.random.default_rng(79)
    for step in range(4):     # consecutive NextQuestion calls exercise both parities of the inbox
        for rep in range(2):
            randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
            chosen = host.next_question_batch(ids, randoms)
            assert np.array_equal(chosen, p2p.next_question_batch(ids, randoms))
            want = host.shards[0]._view(0).cpu().numpy()
            for s in p2p.shards:
                got = s._view(0).cpu().numpy()[:want.size]
                assert np.array_equal(np.isnan(got), np.isnan(want))
                assert np.array_equal(bits(got[~np.isnan(got)]), bits(want[~np.isnan(want)]))
        answers = [(int(c) * 7 + step) % K for c in chosen]
        host.record_answer_batch(ids, answers)
        p2p.record_answer_batch(ids, answers)
        for s in p2p.shards:
            for q in ids[:5]:
                assert np.array_equal(bits(s.copy_quiz_priors(int(q))), bits(host.copy_quiz_priors(int(q))))
        items, counts = p2p.list_top_targets_batch(ids, 10)
        items_h, counts_h = host.list_top_targets_batch(ids, 10)
        assert np.array_equal(counts, counts_h) and items.tobytes() == items_h.tobytes()
    # a smaller batch than the inbox was sized for, and a call sequence error
    sub = ids[:3]
    randoms = rng.integers(0, 2 ** 64, size=3, dtype=np.uint64)
    assert np.array_equal(host.next_question_batch(sub, randoms), p2p.next_question_batch(sub, randoms))
    for s in p2p.shards:
        s.engine.p2p_next_question_begin(sub, randoms)
    with pytest.raises(pqa.PqaException):
        p2p.shards[0].engine.p2p_next_question_begin(sub, randoms)      # its End is pending
    for s in p2p.shards:
        s.engine.p2p_next_question_end(sub)


def test_peer_memory_exchange_across_processes():
    """One process per GPU, cudaIpc-mapped inboxes over NVLink (scripts/p2p_multiproc_check.py: un-sharded engine vs NCCL
    exchange vs peer-memory exchange, both sharding axes). Needs at least two GPUs; `profiles/r01_multi_gpu_*.log` hold the
    2- and 8-GPU runs of this round."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (ran on 2 and 8 GPUs this round: profiles/r01_multi_gpu_*.log)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)),
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "p2p_multiproc_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "questions: OK" in out.stdout and "targets: OK" in out.stdout


@pytest.mark.parametrize("kbname,dims,n_shards,n,chunk", [("binary", (64, 5, 1000), 2, 9, 0), ("binary", (200, 5, 1000), 4, 70, 0),
                                                          ("gamma", (48, 5, 2050), 3, 130, 128), ("binary", (300, 5, 500), 3, 40, 0)])
def test_target_shards_exact_order_pipeline(kbname, dims, n_shards, n, chunk):
    """Exact-order pipeline (PqaB200_P2PSetExactOrder): the Kahan lanes of pass 1 travel from shard to shard in target
    order, so W_k is the reference's own sum. Priorities must then meet the SINGLE-engine bar (2e-12 flat against the
    un-sharded staged kernel) even on the binary-search KB, whose uninformative questions amplify any other W_k to
    percent-level priority differences; the plain target-sharded exchange is run beside it to show exactly that."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    W = 6
    kb = synth.binary_search_kb(Q, K, T, INIT, 3) if kbname == "binary" else synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    full = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5)
    full.upload_kb(*kb)

    def make(exact):
        shards = []
        for first, count in sharded.target_shard_ranges(T, n_shards):
            e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=5, target_shard_first=first, target_shard_count=count)
            e.upload_kb(*kb)
            e.set_eval_kernel(2, chunk_targets=chunk)
            shards.append(sharded.B200TargetShard(e))
        se = sharded.TargetShardedEngine(shards)
        se.enable_p2p(n, exact_order=exact)
        return se

    exact, plain = make(True), make(False)
    ids = full.start_quiz_batch(n)
    assert np.array_equal(ids, exact.start_quiz_batch(n)) and np.array_equal(ids, plain.start_quiz_batch(n))
    rng = np.random.default_rng(81)
    worst_exact = worst_plain = 0.0
    for step in range(4):
        for rep in range(2):                                   # both parities of the inbox
            randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
            want = full.eval_questions(ids)["priority"].ravel()
            chosen = exact.next_question_batch(ids, randoms)
            plain.next_question_batch(ids, randoms)
            ok = ~np.isnan(want)
            got = [s._view(0).cpu().numpy()[:want.size] for s in exact.shards]
            for g in got[1:]:
                assert np.array_equal(bits(g[ok]), bits(got[0][ok])), "shards disagree"
            assert np.array_equal(np.isnan(got[0]), ~ok)
            worst_exact = max(worst_exact, float(np.max(np.abs(got[0][ok] - want[ok]) / np.abs(want[ok]))))
            gp = plain.shards[0]._view(0).cpu().numpy()[:want.size]
            worst_plain = max(worst_plain, float(np.max(np.abs(gp[ok] - want[ok]) / np.abs(want[ok]))))
        for e in (full, plain):
            e.set_active_question_batch(ids, chosen)
        answers = [(int(c) * 7 + step) % K for c in chosen]
        for e in (full, exact, plain):
            e.record_answer_batch(ids, answers)
    assert worst_exact < 2e-12, worst_exact
    print("max relative priority difference vs the single engine: exact-order %.3g, summed partials %.3g" % (worst_exact, worst_plain))


@pytest.mark.parametrize("axis,dims,n_shards", [("questions", (23, 4, 101), 3), ("targets", (17, 5, 203), 3), ("targets", (9, 3, 64), 2)])
def test_sharded_kb_file_save_and_load(axis, dims, n_shards, tmp_path):
    """Sharded engines save into ONE file in the reference's layout (frame writer first, then every shard in place) and
    load their own rows / columns back out of a file; both must equal what a single engine writes / holds, byte for byte."""
    from probqa_b200 import engine as pqa
    Q, K, T = dims
    kb = synth.gamma_kb(Q, K, T, INIT)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    full = fac.create_b200_engine(edef, emulated_workers=4, rng_seed=5)
    full.upload_kb(*kb)
    ranges = sharded.shard_ranges(Q, n_shards) if axis == "questions" else sharded.target_shard_ranges(T, n_shards)
    key = "question_shard" if axis == "questions" else "target_shard"
    shards = []
    for f, c in ranges:
        e = fac.create_b200_engine(edef, emulated_workers=4, rng_seed=5, **{key + "_first": f, key + "_count": c})
        e.upload_kb(*kb)
        shards.append(sharded.B200Shard(e) if axis == "questions" else sharded.B200TargetShard(e))
    se = (sharded.QuestionShardedEngine if axis == "questions" else sharded.TargetShardedEngine)(shards)
    # train a little so that the file is not the uploaded KB, identically on both sides
    aqs = [pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(3, 0), pqa.AnsweredQuestion(1, 1)]
    full.train(aqs, 5, 0.75)
    se.train(aqs, 5, 0.75)
    p_full, p_shards = str(tmp_path / "full.kb"), str(tmp_path / "shards.kb")
    full.save_kb(p_full)
    se.save_kb(p_shards)
    assert open(p_full, "rb").read() == open(p_shards, "rb").read()
    want = full.download_kb()
    gA, gD = np.full_like(want[0], np.nan), np.full_like(want[1], np.nan)
    for f, c in ranges:
        e = fac.load_b200_engine(p_shards, emulated_workers=4, **{key + "_first": f, key + "_count": c})
        a, d, b = e.download_kb()
        if axis == "questions":
            gA[f:f + c], gD[f:f + c] = a[f:f + c], d[f:f + c]
        else:
            gA[:, :, f:f + c], gD[:, f:f + c] = a[:, :, f:f + c], d[:, f:f + c]
        assert np.array_equal(bits(b), bits(want[2])) and e.get_total_questions_asked() == full.get_total_questions_asked()
    assert np.array_equal(bits(gA), bits(want[0])) and np.array_equal(bits(gD), bits(want[1]))
    one = fac.load_b200_engine(p_shards, emulated_workers=4)               # and as a single engine, streamed
    for g, w in zip(one.download_kb(), want):
        assert np.array_equal(bits(g), bits(w))
