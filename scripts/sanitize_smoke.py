"""Tiny run that touches every kernel (all evaluation variants, chunked and single-chunk), meant to be run under
compute-sanitizer --tool memcheck / racecheck / initcheck on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from probqa_b200 import engine as pqa, synth

Q, K, T, W = 12, 5, 70, 3
kb = synth.gamma_kb(Q, K, T, 0.1)
eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W, rng_seed=9)
eng.upload_kb(*kb)
ids = eng.start_quiz_batch(40)
for step in range(2):
    qs = eng.next_question_batch(ids)
    eng.record_answer_batch(ids, (qs + step) % K)
ref = None
for which, chunk, lanes, n in ((1, 0, 0, 40), (2, 0, 0, 3), (2, 0, 1, 40), (2, 0, 2, 40), (2, 0, 4, 40), (2, 32, 1, 5), (2, 32, 2, 40), (2, 32, 4, 40)):
    eng.set_eval_kernel(which, chunk, 0, lanes)
    pri = eng.eval_questions(ids[:n])["priority"]
    if ref is None:
        ref = pri
    ok = ~np.isnan(ref[:n])
    assert np.array_equal(np.isnan(pri), np.isnan(ref[:n]))
    assert np.allclose(pri[ok], ref[:n][ok], rtol=2e-12, atol=0), (which, chunk, lanes, n)
eng.set_eval_kernel(0)
eng.eval_questions_detailed(int(ids[0]))
eng.list_top_targets_batch(ids, 10)
eng.list_top_targets(int(ids[1]), 7)
eng.record_quiz_target_batch(ids, np.arange(40) % T)
eng.train([pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(1, 2), pqa.AnsweredQuestion(3, 0)], 5, 1.5)
eng.resident_bind(ids); eng.resident_step(); eng.resident_fetch(); eng.flush_l2(); eng.synchronize()
eng.download_kb(); eng.copy_a_targets(2, 1)
eng.release_quiz_batch(ids)
print("sanitize smoke ok")

# ---- kernels added later in round 1: target / question shards with both exchanges, exact-order pipeline, maintenance,
# device-grouped training, resume
from probqa_b200 import sharded
fac = pqa.PqaEngineFactory()
edef = pqa.EngineDefinition(K, Q, T, init_amount=0.1)
for axis in ("questions", "targets", "targets-exact"):
    shards = []
    ranges = sharded.shard_ranges(Q, 2) if axis == "questions" else sharded.target_shard_ranges(T, 2)
    for f, c in ranges:
        kw = dict(question_shard_first=f, question_shard_count=c) if axis == "questions" else dict(target_shard_first=f, target_shard_count=c)
        e = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=9, **kw)
        e.upload_kb(*kb)
        e.set_eval_kernel(2, 32 if axis != "questions" else 0)
        shards.append(sharded.B200Shard(e) if axis == "questions" else sharded.B200TargetShard(e))
    for p2p in (False, True):
        if axis == "targets-exact" and not p2p:
            continue
        se = (sharded.QuestionShardedEngine if axis == "questions" else sharded.TargetShardedEngine)(shards)
        if p2p:
            se.enable_p2p(40, exact_order=axis == "targets-exact")
        sid = se.start_quiz_batch(40 if axis != "targets-exact" else 7)
        for step in range(2):
            ch = se.next_question_batch(sid, np.arange(sid.size, dtype=np.uint64) * 977 + step)
            se.record_answer_batch(sid, (ch + step) % K)
        se.list_top_targets_batch(sid, 5)
        se.record_quiz_target_batch(sid, np.arange(sid.size) % T)
        se.release_quiz_batch(sid)
        if p2p:
            break      # an engine gets one inbox
    del se, shards
m = fac.create_b200_engine(edef, emulated_workers=W, rng_seed=9)
m.upload_kb(*kb)
m.fill_binary_search_kb(2)
m.start_maintenance(True)
m.remove_targets([3, 9, 40]); m.remove_questions([2])
m.finish_maintenance()
mq = m.start_quiz_batch(6)
m.record_answer_batch(mq, m.next_question_batch(mq) % K)
m.eval_questions(mq); m.list_top_targets_batch(mq, 4)
m.start_maintenance(True)
m.add_qs_ts([0.3, 0.4], [0.5, 0.6, 0.7, 0.8, 0.9])
m.remove_targets([1]); m.compact()
m.finish_maintenance()
mq = m.start_quiz_batch(3)
m.next_question_batch(mq)
rq = m.resume_quiz_batch([[(0, 1), (3, 2)], [(1, 0)]])
big = m.start_quiz_batch(1500)
for s_ in range(6):
    m.set_active_question_batch(big, (np.arange(1500) + s_) % m.n_questions)
    m.record_answer_batch(big, (np.arange(1500) * 3 + s_) % K)
m.record_quiz_target_batch(big, np.arange(1500) % m.n_targets)      # >= 8192 cell operations: device-grouped path
m.clear_old_quizzes(10, 1e9)
print("sanitize smoke (round-1 additions) ok")
