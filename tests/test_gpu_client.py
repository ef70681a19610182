"""The reference's PqaClient learner loop (ProbQA/PqaClient/PqaClient.cpp:69-245) re-expressed in C++ over the C ABI
(clients/pqa_client.cpp, reference symbols only), run against the B200 engine: it must complete, write progress lines in
the reference's format, learn (top-1 precision rises), and its questions/s column -- the only throughput the reference
publishes (mean 301 questions/s) -- is reported."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_pqa_client_learner_loop(tmp_path):
    from probqa_b200 import build
    exe = build.build_client()
    progress = str(tmp_path / "progress.txt")
    out = subprocess.run([exe, "--trainings", "6000", "--learners", "48", "--report-every", "1024", "--progress", progress,
                          "--kb-dir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["failed"] is False and line["questions_asked"] > 6000
    rows = [r.split("\t") for r in open(progress).read().strip().splitlines()]
    assert len(rows) == 5 and all(len(r) == 6 for r in rows)              # trainings 1024 .. 5120
    assert [int(r[0]) for r in rows] == [1024 * (x + 1) for x in range(5)]
    precision = [float(r[2]) for r in rows]
    assert precision[-1] > precision[0]                                    # it learns
    assert float(rows[-1][5]) > 301.0                                      # reference-published questions/s (unrecorded CPU)
    assert os.path.exists(str(tmp_path / "dichotomy001024.kb"))           # KB snapshots like the reference's (PqaClient.cpp:92-98)
    print("pqa_client:", json.dumps(line))


def test_cpp_shard_launcher(tmp_path):
    """clients/pqa_shard_launcher.cpp: one process per GPU over the C ABI only (fork, cudaIpc inbox handles over UNIX socket
    pairs, lockstep P2P calls) -- the multi-process form without Python or torch. One GPU: the launcher's single-rank path
    against the same workload through the Python binding; two or more GPUs: all ranks must choose the same questions, and
    the same ones a single engine chooses (exact-order pipeline: priorities at the single-engine bar)."""
    import json
    import subprocess
    import numpy as np
    import torch
    from probqa_b200 import build, engine as pqa, synth
    exe = build.build_launcher()
    Q, K, T, B = 300, 5, 4000, 70
    def run(n, extra=()):
        out = subprocess.run([exe, "--gpus", str(n), "--questions", str(Q), "--answers", str(K), "--targets", str(T), "--batch", str(B),
                              "--steps", "3", "--warmup", "2", *extra], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])
    one = run(1)
    assert one["all_ranks_chose_the_same_questions"] and one["value"] > 0
    # the same batch through the Python binding on one engine
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), rng_seed=1234, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    ids = eng.start_quiz_batch(B)
    states = [synth.quiz_prefix(b, (0, 3, 8)[b % 3], Q, T, K) for b in range(B)]
    for s in range(8):
        sel = [x for x in range(B) if len(states[x]) > s]
        eng.set_active_question_batch(ids[sel], [states[x][s][0] for x in sel])
        eng.record_answer_batch(ids[sel], [states[x][s][1] for x in sel])
    z, randoms = 0x9E3779B97F4A7C15, []
    for _ in range(B):
        z ^= (z << 13) & 0xFFFFFFFFFFFFFFFF; z ^= z >> 7; z ^= (z << 17) & 0xFFFFFFFFFFFFFFFF
        randoms.append(z)
    chosen = eng.next_question_batch(ids, np.array(randoms, dtype=np.uint64))
    assert one["chosen_checksum"] == int(sum(int(c) * (b + 1) for b, c in enumerate(chosen)) % 1000000007)
    n = min(torch.cuda.device_count(), 8)
    if n >= 2:
        for axis, extra in (("targets", ("--exact-order",)), ("questions", ())):
            many = run(n, ("--axis", axis, *extra))
            assert many["all_ranks_chose_the_same_questions"] and many["chosen_checksum"] == one["chosen_checksum"], (axis, many)
