"""Question-sharded multi-GPU engine: one B200 engine per GPU, each holding the sA/mD rows of a contiguous slice of the
questions (SURVEY.md 8e "Questions" row; the reference itself parallelises NextQuestion over questions,
CpuEngine.cpp:355-360). Quiz state (priors, asked bits) is replicated on every shard.

  NextQuestion : every shard evaluates its questions for all quizzes of the batch into a zero-initialised [n][Q] buffer
                 -> all-reduce(sum) of the buffers (NCCL over NVLink through torch.distributed) -> every shard runs the
                 same selection with the same 64-bit draws. Adding the other shards' +0.0 is exact, so priorities,
                 run-lengths and the chosen questions are those of a single engine.
  RecordAnswer : the shard that owns the answered question computes the new posterior, the others contribute zero rows
                 -> all-reduce(sum) of the [n][Tp] rows -> every shard stores the row (bit-identical to a single engine).
  StartQuiz / ListTopTargets / SetActiveQuestion are local; RecordQuizTarget / Train touch only the owner's cells (vB is
  replicated and updated everywhere).

`shards` is the list of shard ports driven by THIS process: one per process under torchrun (pass the process group), or
several in one process (tests on a single GPU; group=None). A shard port is anything with the methods of B200Shard.
"""
from typing import List, Optional, Sequence

import numpy as np


def shard_ranges(n_questions: int, n_shards: int):
    """[(first, count)]: contiguous, the first n_questions % n_shards shards get one more (like CalcSplit,
    SRPoolRunner.h:96-110)."""
    quot, rem = divmod(n_questions, n_shards)
    out, first = [], 0
    for s in range(n_shards):
        cnt = quot + (1 if s < rem else 0)
        out.append((first, cnt))
        first += cnt
    return out


def target_shard_ranges(n_targets: int, n_shards: int):
    """[(first, count)] over targets in units of 4-target vectors (a target keeps its Kahan lane j % 4 inside a shard;
    only the last shard can end off a multiple of 4, where the global padding lanes are)."""
    out = []
    if n_shards > (n_targets + 3) // 4:
        raise ValueError("more target shards (%d) than 4-target vectors (%d)" % (n_shards, (n_targets + 3) // 4))
    for first_v, cnt_v in shard_ranges((n_targets + 3) // 4, n_shards):
        first = 4 * first_v
        out.append((first, max(0, min(4 * cnt_v, n_targets - first))))
    return out


class _CudaView:
    """Exposes engine-owned device memory to torch without a copy (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class B200Shard:
    """Shard port over a PqaEngine created with a question shard (probqa_b200.engine)."""

    def __init__(self, engine):
        self.engine = engine
        self.first, self.count = engine.question_shard()

    def _view(self, which):
        import torch
        ptr, cnt = self.engine.shard_buffer(which)
        return torch.as_tensor(_CudaView(ptr, cnt), device="cuda")

    def start_quiz_batch(self, n):
        return self.engine.start_quiz_batch(n)

    def eval(self, quiz_ids):
        self.engine.shard_eval(quiz_ids)
        return self._view(0)

    def select(self, quiz_ids, randoms):
        return self.engine.shard_select(quiz_ids, randoms)

    def record_answer_begin(self, quiz_ids, answers):
        self.engine.shard_record_answer_begin(quiz_ids, answers)
        return self._view(1)

    def record_answer_end(self, quiz_ids):
        self.engine.shard_record_answer_end(quiz_ids)

    def set_active_question_batch(self, quiz_ids, questions):
        self.engine.set_active_question_batch(quiz_ids, questions)

    def list_top_targets_batch(self, quiz_ids, max_count):
        return self.engine.list_top_targets_batch(quiz_ids, max_count)

    def record_quiz_target_batch(self, quiz_ids, targets, amounts=None):
        self.engine.record_quiz_target_batch(quiz_ids, targets, amounts)

    def train(self, answered_questions, i_target, amount=1.0):
        self.engine.train(answered_questions, i_target, amount)

    def release_quiz_batch(self, quiz_ids):
        self.engine.release_quiz_batch(quiz_ids)

    def copy_quiz_priors(self, quiz):
        return self.engine.copy_quiz_priors(quiz)


class QuestionShardedEngine:
    def __init__(self, shards: Sequence, group=None, seed: int = 0x5EED):
        assert len(shards) >= 1
        self.shards = list(shards)
        self.group = group
        self._rng = np.random.default_rng(seed)   # same seed on every rank => same draws on every shard
        self._p2p = False

    # ---- exchange over peer memory instead of the caller-side all-reduce (PqaB200Ext.h "P2P" entry points)
    def enable_p2p(self, max_quizzes: int, exact_order: bool = False):
        """Gives every shard engine an inbox and connects them: directly when all shards live in this process
        (group=None), through cudaIpc handles all-gathered over the process group when there is one shard per process.
        From then on next_question_batch / record_answer_batch run without any host-side exchange."""
        engines = [s.engine for s in self.shards]
        if self.group is None:
            bases = [e.p2p_init(r, len(engines), max_quizzes)[0] for r, e in enumerate(engines)]
            for e in engines:
                e.p2p_connect(bases)
        else:
            import torch.distributed as dist
            assert len(engines) == 1, "one shard per process under a process group"
            rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
            base, _ = engines[0].p2p_init(rank, world, max_quizzes)
            handles = [None] * world
            dist.all_gather_object(handles, engines[0].p2p_export_handle(), group=self.group)
            bases = [base if r == rank else engines[0].p2p_open_handle(handles[r]) for r in range(world)]
            engines[0].p2p_connect(bases)
            dist.barrier(group=self.group)
        if exact_order:      # target shards: W_k summed in the reference's own order across the shards (PqaB200Ext.h)
            for e in engines:
                e.p2p_set_exact_order(True)
        self._p2p = True

    def _p2p_next_question(self, quiz_ids, randoms):
        for s in self.shards:
            s.engine.p2p_next_question_begin(quiz_ids, randoms)
        return self._same([s.engine.p2p_next_question_end(quiz_ids) for s in self.shards])

    def _p2p_record_answer(self, quiz_ids, answers):
        for s in self.shards:
            s.engine.p2p_record_answer_begin(quiz_ids, answers)
        for s in self.shards:
            s.engine.p2p_record_answer_end()

    # ---- the exchange step
    def _all_reduce(self, tensors: List):
        """Sum over the local shards' tensors and over the process group; every local tensor receives the total."""
        total = tensors[0]
        for t in tensors[1:]:
            total.add_(t)
        if self.group is not None:
            import torch.distributed as dist
            if total.is_cuda:
                import torch
                torch.cuda.synchronize()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        if total.is_cuda:
            import torch
            for t in tensors[1:]:
                t.copy_(total)
            torch.cuda.synchronize()   # the engines consume the buffers on their own streams
        else:
            for t in tensors[1:]:
                t.copy_(total)

    def _same(self, results):
        first = results[0]
        for r in results[1:]:
            assert np.array_equal(np.asarray(first), np.asarray(r)), "shards disagree"
        return first

    # ---- quiz path
    def start_quiz_batch(self, n: int) -> np.ndarray:
        return self._same([s.start_quiz_batch(n) for s in self.shards])

    def next_question_batch(self, quiz_ids, randoms: Optional[np.ndarray] = None) -> np.ndarray:
        quiz_ids = np.ascontiguousarray(quiz_ids, dtype=np.int64)
        if randoms is None:
            randoms = self._rng.integers(0, 2 ** 64, size=quiz_ids.size, dtype=np.uint64)
        if self._p2p:
            return self._p2p_next_question(quiz_ids, randoms)
        self._all_reduce([s.eval(quiz_ids) for s in self.shards])
        return self._same([s.select(quiz_ids, randoms) for s in self.shards])

    def eval_priorities(self, quiz_ids):
        """All-reduced priorities [n, Q] (NaN where asked) as a host array, without selecting."""
        quiz_ids = np.ascontiguousarray(quiz_ids, dtype=np.int64)
        bufs = [s.eval(quiz_ids) for s in self.shards]
        self._all_reduce(bufs)
        return bufs[0].detach().cpu().numpy().reshape(quiz_ids.size, -1).copy()

    def record_answer_batch(self, quiz_ids, answers):
        quiz_ids = np.ascontiguousarray(quiz_ids, dtype=np.int64)
        if self._p2p:
            return self._p2p_record_answer(quiz_ids, answers)
        self._all_reduce([s.record_answer_begin(quiz_ids, answers) for s in self.shards])
        for s in self.shards:
            s.record_answer_end(quiz_ids)

    def set_active_question_batch(self, quiz_ids, questions):
        for s in self.shards:
            s.set_active_question_batch(quiz_ids, questions)

    def list_top_targets_batch(self, quiz_ids, max_count: int):
        return self.shards[0].list_top_targets_batch(quiz_ids, max_count)   # replicated state: any shard answers

    def record_quiz_target_batch(self, quiz_ids, targets, amounts=None):
        for s in self.shards:
            s.record_quiz_target_batch(quiz_ids, targets, amounts)

    def train(self, answered_questions, i_target, amount=1.0):
        for s in self.shards:
            s.train(answered_questions, i_target, amount)

    def release_quiz_batch(self, quiz_ids):
        for s in self.shards:
            s.release_quiz_batch(quiz_ids)

    def copy_quiz_priors(self, quiz):
        return self.shards[0].copy_quiz_priors(quiz)

    def save_kb(self, file_path: str):
        """One KB file in the reference's layout out of all shards: the first shard writes the frame (header, vB, id maps)
        and its cells, then every other shard writes its cells in place. Load it back per shard with
        PqaEngineFactory.load_b200_engine(path, question_shard_... / target_shard_...)."""
        rank = 0
        if self.group is not None:
            import torch.distributed as dist
            rank = dist.get_rank(self.group)
        if rank == 0:
            self.shards[0].engine.save_kb_shard(file_path, True)
        if self.group is not None:
            dist.barrier(group=self.group)
        for x, s in enumerate(self.shards):
            if not (rank == 0 and x == 0):
                s.engine.save_kb_shard(file_path, False)
        if self.group is not None:
            dist.barrier(group=self.group)


class B200TargetShard:
    """Shard port over a PqaEngine created with a target shard (probqa_b200.engine): the columns
    [first, first + count) of every sA/mD row live on this device, quiz state is replicated."""

    def __init__(self, engine):
        self.engine = engine
        self.first, self.count = engine.target_shard()

    def _view(self, which):
        import torch
        ptr, cnt = self.engine.shard_buffer(which)
        return torch.as_tensor(_CudaView(ptr, cnt), device="cuda")

    def start_quiz_batch(self, n):
        return self.engine.start_quiz_batch(n)

    def eval_w(self, quiz_ids):
        self.engine.tshard_eval_w(quiz_ids)
        return self._view(2)

    def eval_hvl(self, quiz_ids):
        self.engine.tshard_eval_hvl(quiz_ids)
        return self._view(3)

    def priority(self, quiz_ids):
        self.engine.tshard_priority(quiz_ids)
        return self._view(0)

    def select(self, quiz_ids, randoms):
        return self.engine.shard_select(quiz_ids, randoms)

    def record_answer_begin(self, quiz_ids, answers):
        self.engine.shard_record_answer_begin(quiz_ids, answers)
        return self._view(1)

    def record_answer_end(self, quiz_ids):
        self.engine.shard_record_answer_end(quiz_ids)

    def set_active_question_batch(self, quiz_ids, questions):
        self.engine.set_active_question_batch(quiz_ids, questions)

    def list_top_targets_batch(self, quiz_ids, max_count):
        return self.engine.list_top_targets_batch(quiz_ids, max_count)

    def record_quiz_target_batch(self, quiz_ids, targets, amounts=None):
        self.engine.record_quiz_target_batch(quiz_ids, targets, amounts)

    def train(self, answered_questions, i_target, amount=1.0):
        self.engine.train(answered_questions, i_target, amount)

    def release_quiz_batch(self, quiz_ids):
        self.engine.release_quiz_batch(quiz_ids)

    def copy_quiz_priors(self, quiz):
        return self.engine.copy_quiz_priors(quiz)


class TargetShardedEngine(QuestionShardedEngine):
    """Target-sharded multi-GPU engine (SURVEY.md 8e "Targets" row, BASELINE config 4): every shard holds a slice of the
    targets of every row. NextQuestion is two-phase with an all-reduce after each phase:

      phase 1: partial W_k[n][Q][K] over the local targets            -> all-reduce(sum)
      phase 2: partial sum post*log2 post, sum (post-prior)^2, lack   -> all-reduce(sum) -> priority epilogue -> selection

    The sums over shards are not in CpuEngine's order, so priorities are tolerance-level here (identical on all shards:
    every shard receives the same reduced bits), unlike the question-sharded engine. Posteriors stay bit-exact:
    RecordAnswer all-reduces the shards' disjoint column slices of the un-normalised row (x + 0 is exact) and every shard
    normalises the complete row in the reference's order. The remaining calls are inherited."""

    def _priorities(self, quiz_ids):
        self._all_reduce([s.eval_w(quiz_ids) for s in self.shards])
        self._all_reduce([s.eval_hvl(quiz_ids) for s in self.shards])
        return [s.priority(quiz_ids) for s in self.shards]

    def next_question_batch(self, quiz_ids, randoms: Optional[np.ndarray] = None) -> np.ndarray:
        quiz_ids = np.ascontiguousarray(quiz_ids, dtype=np.int64)
        if randoms is None:
            randoms = self._rng.integers(0, 2 ** 64, size=quiz_ids.size, dtype=np.uint64)
        if self._p2p:
            return self._p2p_next_question(quiz_ids, randoms)
        self._priorities(quiz_ids)
        return self._same([s.select(quiz_ids, randoms) for s in self.shards])

    def eval_priorities(self, quiz_ids):
        quiz_ids = np.ascontiguousarray(quiz_ids, dtype=np.int64)
        bufs = self._priorities(quiz_ids)
        out = [b.detach().cpu().numpy().reshape(quiz_ids.size, -1).copy() for b in bufs]
        for o in out[1:]:
            assert np.array_equal(out[0].view(np.uint64), o.view(np.uint64)), "shards computed different priorities"
        return out[0]
