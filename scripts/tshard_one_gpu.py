"""Two target-shard engines on one GPU running NextQuestion over peer-memory exchange (profiling harness for
k_eval_tshard: 1000 x 5 x 1000, 256 quizzes, each shard holds 500 targets)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probqa_b200 import engine as pqa, sharded  # noqa: E402

Q, K, T, B = 1000, 5, 1000, 256
fac = pqa.PqaEngineFactory()
edef = pqa.EngineDefinition(K, Q, T, init_amount=0.1)
shards = []
for f, c in sharded.target_shard_ranges(T, 2):
    e = fac.create_b200_engine(edef, emulated_workers=16, rng_seed=5, target_shard_first=f, target_shard_count=c, initial_quiz_capacity=B)
    e.fill_binary_search_kb(3)
    shards.append(sharded.B200TargetShard(e))
eng = sharded.TargetShardedEngine(shards)
eng.enable_p2p(B)
ids = eng.start_quiz_batch(B)
rnd = np.random.default_rng(1).integers(0, 2 ** 64, size=B, dtype=np.uint64)
for _ in range(4):
    chosen = eng.next_question_batch(ids, rnd)
print("chosen", chosen[:8])
