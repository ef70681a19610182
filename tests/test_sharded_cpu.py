"""CPU test of the multi-GPU (question-sharded) host logic with world_size = 2 over gloo: the orchestrator
(probqa_b200/sharded.py) drives one shard port per process; here the shard port is backed by the CPU oracle, so the
exchange protocol (zero-padded all-reduce of priorities and of updated priors, identical draws on every rank) is checked
end to end against a single un-sharded run, bit for bit. The CUDA shard port is covered by tests/test_gpu_sharded.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from probqa_b200 import sharded, synth  # noqa: E402

Q, K, T, W = 22, 5, 64, 2


def kahan_prefix(values):
    s = c = 0.0
    out = []
    for v in values:
        y = v - c
        t = s + y
        c = (t - s) - y
        s = t
        out.append(s - c)
    return out


class OracleShard:
    """Shard port (same methods as sharded.B200Shard) computing with the CPU oracle on host tensors."""

    def __init__(self, kb, first, count):
        from oracle import oracle as ora
        self.ora = ora
        self.sA, self.mD, self.vB = kb[0][first:first + count], kb[1][first:first + count], kb[2]
        self.first, self.count = first, count
        self.quizzes = {}
        self._pri = self._rows = None

    def start_quiz_batch(self, n):
        ids = []
        for _ in range(n):
            qid = len(self.quizzes)
            self.quizzes[qid] = dict(prior=self.ora.start_quiz(self.vB, W), asked=np.zeros(Q, dtype=bool), active=-1)
            ids.append(qid)
        return np.array(ids, dtype=np.int64)

    def eval(self, quiz_ids):
        buf = torch.zeros(len(quiz_ids) * Q, dtype=torch.float64)
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            for i in range(self.first, self.first + self.count):
                buf[x * Q + i] = float("nan") if z["asked"][i] else self.ora.eval_question(
                    self.sA[i - self.first], self.mD[i - self.first], z["prior"])["priority"]
        self._pri = buf
        return buf

    def select(self, quiz_ids, randoms):
        bounds = self.ora.calc_split(Q, 8 * W)
        out = []
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            pri = self._pri[x * Q:(x + 1) * Q].numpy()
            run, first = np.empty(Q), 0
            for lim in bounds:
                vals = kahan_prefix([0.0 if z["asked"][i] else pri[i] for i in range(first, lim)])
                run[first:lim] = vals
                first = lim
            grand = np.array(kahan_prefix([run[b - 1] for b in bounds]))
            q = self.ora.select_question(dict(runLength=run, grand=grand, bounds=bounds), Q, int(randoms[x]), asked=z["asked"])
            z["active"] = q
            out.append(q)
        return np.array(out, dtype=np.int64)

    def record_answer_begin(self, quiz_ids, answers):
        rows = torch.zeros(len(quiz_ids) * T, dtype=torch.float64)
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            q = z["active"]
            if self.first <= q < self.first + self.count:
                new = self.ora.record_answer(z["prior"], self.sA[q - self.first, int(answers[x])], self.mD[q - self.first], max(1, W - 1))
                rows[x * T:(x + 1) * T] = torch.from_numpy(new)
            z["asked"][q] = True
            z["active"] = -1
        self._rows = rows
        return rows

    def record_answer_end(self, quiz_ids):
        for x, qid in enumerate(quiz_ids):
            self.quizzes[int(qid)]["prior"] = self._rows[x * T:(x + 1) * T].numpy().copy()

    def set_active_question_batch(self, quiz_ids, questions):
        for qid, q in zip(quiz_ids, questions):
            self.quizzes[int(qid)]["active"] = int(q)

    def list_top_targets_batch(self, quiz_ids, max_count):
        return [self.ora.list_top_targets(self.quizzes[int(q)]["prior"], W, max_count) for q in quiz_ids]

    def copy_quiz_priors(self, quiz):
        return self.quizzes[int(quiz)]["prior"].copy()


class OracleTargetShard(OracleShard):
    """Target-shard port (same methods as sharded.B200TargetShard): holds columns [first, first+count) of every row and
    full-length priors; partial sums in numpy, the bit-exact pieces (normalisation, selection, top-k) from the oracle."""

    def __init__(self, kb, first, count):
        super().__init__(kb, 0, Q)
        self.t0, self.t1 = first, first + count
        self.sA, self.mD = kb[0][:, :, first:first + count], kb[1][:, first:first + count]
        self._w = self._hvl = None

    def _lik(self, i, prior):
        invD = 1.0 / self.mD[i]
        return (self.sA[i] * invD[None, :]) * prior[None, self.t0:self.t1], invD

    def eval_w(self, quiz_ids):
        buf = torch.zeros(len(quiz_ids) * Q * K, dtype=torch.float64)
        w = buf.numpy().reshape(len(quiz_ids), Q, K)
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            for i in range(Q):
                if not z["asked"][i]:
                    w[x, i] = self._lik(i, z["prior"])[0].sum(axis=1)
        self._w = buf
        return buf

    def eval_hvl(self, quiz_ids):
        NV = 2 * K + 1
        buf = torch.zeros(len(quiz_ids) * Q * NV, dtype=torch.float64)
        hvl = buf.numpy().reshape(len(quiz_ids), Q, NV)
        w = self._w.numpy().reshape(len(quiz_ids), Q, K)      # all-reduced by now
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            for i in range(Q):
                if z["asked"][i]:
                    continue
                lik, invD = self._lik(i, z["prior"])
                post = lik * (1.0 / w[x, i])[:, None]
                l2 = np.log2(post)
                hvl[x, i, :K] = (post * l2).sum(axis=1)
                hvl[x, i, K:2 * K] = ((post - z["prior"][None, self.t0:self.t1]) ** 2).sum(axis=1)
                hvl[x, i, 2 * K] = ((invD * invD)[None, :] / l2).sum()
        self._hvl = buf
        return buf

    def priority(self, quiz_ids):
        NV = 2 * K + 1
        w = self._w.numpy().reshape(len(quiz_ids), Q, K)
        hvl = self._hvl.numpy().reshape(len(quiz_ids), Q, NV)
        buf = torch.full((len(quiz_ids) * Q,), float("nan"), dtype=torch.float64)
        ln_sqrt2 = 0.34657359027997265
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            for i in range(Q):
                if z["asked"][i]:
                    continue
                tot = w[x, i].sum()
                avg_h = (w[x, i] * -hvl[x, i, :K]).sum() / tot          # CEEvalQsSubtaskConsider.cpp:134-177
                avg_v = (w[x, i] * np.sqrt(hvl[x, i, K:2 * K])).sum() / tot
                v_comp = 1.0 / (ln_sqrt2 - np.log(avg_v) + ln_sqrt2 / float((T + 1) * (T + 1)))   # :24-33
                buf[x * Q + i] = -hvl[x, i, 2 * K] * v_comp ** 9 * np.exp2(avg_h) ** -2           # :201-207
        self._pri = buf
        return buf

    def record_answer_begin(self, quiz_ids, answers):
        rows = torch.zeros(len(quiz_ids) * T, dtype=torch.float64)
        for x, qid in enumerate(quiz_ids):
            z = self.quizzes[int(qid)]
            q = z["active"]
            m = z["prior"][self.t0:self.t1] * (self.sA[q, int(answers[x])] / self.mD[q])    # divide first, then multiply
            rows[x * T + self.t0:x * T + self.t1] = torch.from_numpy(m)
            z["asked"][q] = True
            z["active"] = -1
        self._rows = rows
        return rows

    def record_answer_end(self, quiz_ids):
        one = np.ones(T)
        for x, qid in enumerate(quiz_ids):     # the complete un-normalised row, normalised in the reference's order
            m = self._rows[x * T:(x + 1) * T].numpy().copy()
            self.quizzes[int(qid)]["prior"] = self.ora.record_answer(m, one, one, max(1, W - 1))


def _tworker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kb = synth.gamma_kb(Q, K, T, 0.1)
        first, count = sharded.target_shard_ranges(T, world)[rank]
        eng = sharded.TargetShardedEngine([OracleTargetShard(kb, first, count)], group=dist.group.WORLD)
        ids = eng.start_quiz_batch(2)
        pri = eng.eval_priorities(ids)
        want = np.array([[eng.shards[0].ora.eval_question(kb[0][i], kb[1][i], eng.shards[0].quizzes[int(q)]["prior"])["priority"]
                          for i in range(Q)] for q in ids])
        ok = bool(np.max(np.abs(pri - want) / want) < 1e-9)
        got = run_quizzes(eng)
        # posteriors of a target-sharded run are bit-exact given the same (question, answer) sequence
        ref = sharded.QuestionShardedEngine([OracleShard(kb, 0, Q)])
        rids = ref.start_quiz_batch(2)      # mirror of the eval_priorities quizzes
        rids = ref.start_quiz_batch(3)
        for step, (chosen, priors, tops) in enumerate(got):
            ref.set_active_question_batch(rids, chosen)
            ref.record_answer_batch(rids, [(int(c) + step) % K for c in chosen])
            ok &= all(np.array_equal(a.view(np.uint64), ref.copy_quiz_priors(i).view(np.uint64)) for a, i in zip(priors, rids))
            ok &= tops == ref.list_top_targets_batch(rids, 5)
        results.put((rank, bool(ok), [c.tolist() for c, _, _ in got]))
    finally:
        dist.destroy_process_group()


def run_quizzes(eng, n_quizzes=3, n_steps=4):
    """A deterministic little session; returns everything observable."""
    rng = np.random.default_rng(2024)
    ids = eng.start_quiz_batch(n_quizzes)
    log = []
    for step in range(n_steps):
        randoms = rng.integers(0, 2 ** 64, size=n_quizzes, dtype=np.uint64)
        chosen = eng.next_question_batch(ids, randoms)
        answers = [(int(c) + step) % K for c in chosen]
        eng.record_answer_batch(ids, answers)
        log.append((chosen.copy(), [eng.copy_quiz_priors(i).copy() for i in ids], eng.list_top_targets_batch(ids, 5)))
    return log


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kb = synth.gamma_kb(Q, K, T, 0.1)
        first, count = sharded.shard_ranges(Q, world)[rank]
        eng = sharded.QuestionShardedEngine([OracleShard(kb, first, count)], group=dist.group.WORLD)
        got = run_quizzes(eng)
        want = run_quizzes(sharded.QuestionShardedEngine([OracleShard(kb, 0, Q)]))
        ok = True
        for (c1, p1, t1), (c2, p2, t2) in zip(got, want):
            ok &= np.array_equal(c1, c2) and t1 == t2
            ok &= all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(p1, p2))
        results.put((rank, bool(ok), [c.tolist() for c, _, _ in got]))
    finally:
        dist.destroy_process_group()


def test_shard_ranges():
    assert sharded.shard_ranges(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sharded.shard_ranges(1000, 8) == [(125 * s, 125) for s in range(8)]
    assert sum(c for _, c in sharded.shard_ranges(10007, 8)) == 10007


def test_question_sharded_protocol_world2_gloo(ora):
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    out = [results.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
    assert out[0][2] == out[1][2]            # both ranks chose the same questions


def test_question_sharded_in_process_three_shards(ora):
    kb = synth.gamma_kb(Q, K, T, 0.1)
    eng = sharded.QuestionShardedEngine([OracleShard(kb, f, c) for f, c in sharded.shard_ranges(Q, 3)])
    got = run_quizzes(eng)
    want = run_quizzes(sharded.QuestionShardedEngine([OracleShard(kb, 0, Q)]))
    for (c1, p1, t1), (c2, p2, t2) in zip(got, want):
        assert np.array_equal(c1, c2) and t1 == t2
        assert all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(p1, p2))


def test_target_shard_ranges():
    assert sharded.target_shard_ranges(1000, 8) == [(0, 128), (128, 128), (256, 124), (380, 124), (504, 124), (628, 124),
                                                    (752, 124), (876, 124)]
    for T_, n in [(1000, 8), (100000, 8), (203, 2), (96, 4), (10, 3)]:
        r = sharded.target_shard_ranges(T_, n)
        assert sum(c for _, c in r) == T_ and all(f % 4 == 0 for f, _ in r)
        assert all(c % 4 == 0 and c > 0 for _, c in r[:-1]) and r[0][0] == 0 and r[-1][1] > 0
        assert all(r[x][0] + r[x][1] == r[x + 1][0] for x in range(n - 1))
    with pytest.raises(ValueError):
        sharded.target_shard_ranges(7, 3)


def test_target_sharded_protocol_world2_gloo(ora):
    ctx = mp.get_context("spawn")
    results = ctx.Queue()
    port = 29810 + os.getpid() % 150
    procs = [ctx.Process(target=_tworker, args=(r, 2, port, results)) for r in range(2)]
    for p in procs:
        p.start()
    out = [results.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
    assert out[0][2] == out[1][2]            # both ranks chose the same questions
