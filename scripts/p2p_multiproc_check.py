"""Multi-process check of the shard exchange over peer memory (one process per GPU, cudaIpc-mapped inboxes, NVLink):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/p2p_multiproc_check.py
Every rank runs the same quiz session three ways -- an un-sharded engine on its own GPU, shards exchanging through NCCL
all-reduce, shards exchanging over peer memory -- for question shards (everything bit-identical to the single engine) and
target shards (posteriors / top-10 bit-identical, priorities 1e-10, P2P == NCCL choice of questions). Prints one OK line
per axis from rank 0; any mismatch raises."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probqa_b200 import engine as pqa, sharded, synth  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def session(axis, rank, world, local_rank):
    Q, K, T, n, W = 96, 5, 2000, 70, 6
    kb = synth.gamma_kb(Q, K, T, 0.1)
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=0.1)
    full = fac.create_b200_engine(edef, device=local_rank, emulated_workers=W, rng_seed=5)
    full.upload_kb(*kb)

    def make(p2p, exact=False):
        if axis == "questions":
            f, c = sharded.shard_ranges(Q, world)[rank]
            e = fac.create_b200_engine(edef, device=local_rank, emulated_workers=W, rng_seed=5, question_shard_first=f, question_shard_count=c)
            e.upload_kb(*kb)
            se = sharded.QuestionShardedEngine([sharded.B200Shard(e)], group=dist.group.WORLD)
        else:
            f, c = sharded.target_shard_ranges(T, world)[rank]
            e = fac.create_b200_engine(edef, device=local_rank, emulated_workers=W, rng_seed=5, target_shard_first=f, target_shard_count=c)
            e.upload_kb(*kb)
            se = sharded.TargetShardedEngine([sharded.B200TargetShard(e)], group=dist.group.WORLD)
        if p2p:
            se.enable_p2p(n, exact_order=exact)
        return se

    nccl, p2p = make(False), make(True)
    exact = make(True, exact=True) if axis == "targets" else None     # Kahan lanes handed from shard to shard
    ids = full.start_quiz_batch(n)
    assert np.array_equal(ids, nccl.start_quiz_batch(n)) and np.array_equal(ids, p2p.start_quiz_batch(n))
    if exact is not None:
        assert np.array_equal(ids, exact.start_quiz_batch(n))
    worst_exact = 0.0
    rng = np.random.default_rng(80)
    for step in range(5):
        for rep in range(2):
            randoms = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
            c_n = nccl.next_question_batch(ids, randoms)
            c_p = p2p.next_question_batch(ids, randoms)
            want = full.eval_questions(ids)["priority"].ravel()
            got_p = p2p.shards[0]._view(0).cpu().numpy()[:want.size]
            got_n = nccl.shards[0]._view(0).cpu().numpy()[:want.size]
            ok = ~np.isnan(want)
            assert np.array_equal(np.isnan(got_p), ~ok) and np.array_equal(np.isnan(got_n), ~ok)
            if axis == "questions":
                c_f = full.next_question_batch(ids, randoms)
                assert np.array_equal(c_n, c_f) and np.array_equal(c_p, c_f), "chosen questions differ from the single engine"
                assert np.array_equal(bits(got_p[ok]), bits(want[ok])) and np.array_equal(bits(got_n[ok]), bits(want[ok]))
            else:
                rel = max(np.max(np.abs(got_p[ok] - want[ok]) / want[ok]), np.max(np.abs(got_n[ok] - want[ok]) / want[ok]))
                assert rel < 1e-10, rel
                if world == 2:      # a + b == b + a: NCCL's order cannot differ from the rank order
                    assert np.array_equal(bits(got_p[ok]), bits(got_n[ok])) and np.array_equal(c_n, c_p)
                exact.next_question_batch(ids, randoms)
                got_e = exact.shards[0]._view(0).cpu().numpy()[:want.size]
                worst_exact = max(worst_exact, float(np.max(np.abs(got_e[ok] - want[ok]) / want[ok])))
                nccl.set_active_question_batch(ids, c_p)
                full.set_active_question_batch(ids, c_p)
                exact.set_active_question_batch(ids, c_p)
        answers = [(int(c) * 7 + step) % K for c in c_p]
        for e in (full, nccl, p2p) + ((exact,) if exact is not None else ()):
            e.record_answer_batch(ids, answers)
        for q in ids[:8]:
            w = bits(full.copy_quiz_priors(int(q)))
            assert np.array_equal(bits(nccl.copy_quiz_priors(int(q))), w) and np.array_equal(bits(p2p.copy_quiz_priors(int(q))), w)
        it_f, cn_f = full.list_top_targets_batch(ids, 10)
        it_p, cn_p = p2p.list_top_targets_batch(ids, 10)
        assert np.array_equal(cn_f, cn_p) and it_f.tobytes() == it_p.tobytes()
    dist.barrier()
    if exact is not None:
        assert worst_exact < 2e-12, worst_exact          # the single-engine bar, with W_k summed in the reference's order
    if rank == 0:
        print("p2p_multiproc_check %s: OK (world %d)%s" % (axis, world, "" if exact is None else
              "; exact-order pipeline max rel priority diff vs single engine %.3g" % worst_exact), flush=True)


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
    for axis in ("questions", "targets"):
        session(axis, rank, world, local_rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
