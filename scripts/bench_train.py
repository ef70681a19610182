#!/usr/bin/env python
"""BASELINE config 5: 10^6 RecordQuizTarget cell updates = 125 000 quizzes x 8 answered questions on 1000Q x 5A x 1000T,
B200 engine (one PqaEngine_RecordQuizTargetBatch call through the C ABI, host buffers) vs the CPU oracle applying the
same quizzes one RecordQuizTarget at a time (the CPU-baseline leg of this bench: the only place it touches oracle/); the final
sA/mD/vB are compared bit for bit. Prints one JSON line."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probqa_b200 import engine as pqa, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quizzes", type=int, default=125000)
    ap.add_argument("--answers", type=int, default=8)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    Q, K, T, W = 1000, 5, 1000, 8
    kb = synth.binary_search_kb(Q, K, T, 0.1, 3)
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=0.1), emulated_workers=W,
                                                    rng_seed=3, initial_quiz_capacity=args.quizzes)
    eng.upload_kb(*kb)
    n, d = args.quizzes, args.answers
    rng = np.random.default_rng(20171126)
    targets = rng.integers(0, T, size=n)
    # questions: d distinct questions per quiz (fixed-seed stream with realistic collisions across quizzes)
    qs = np.argsort(rng.random((n, 64)), axis=1)[:, :d] + rng.integers(0, Q - 64, size=(n, 1))
    ans = synth.answer_rule(Q, T, K)[qs, targets[:, None]]
    t0 = time.perf_counter()
    quizzes = eng.start_quiz_batch(n)
    for s in range(d):
        eng.set_active_question_batch(quizzes, qs[:, s])
        eng.record_answer_batch(quizzes, ans[:, s])
    t_setup = time.perf_counter() - t0
    eng.synchronize()
    t0 = time.perf_counter()
    eng.record_quiz_target_batch(quizzes, targets)
    eng.synchronize()
    t_gpu = time.perf_counter() - t0
    out = {"metric": "RecordQuizTarget cell updates/s", "updates": n * d, "quizzes": n, "b200_s": t_gpu,
           "b200_updates_per_s": n * d / t_gpu, "setup_s": t_setup, "api": "PqaEngine_RecordQuizTargetBatch (host buffers, one call)"}
    if not args.no_oracle:
        from oracle import oracle as ora
        sA, mD, vB = [a.copy() for a in kb]
        t0 = time.perf_counter()
        for x in range(n):
            ora.record_quiz_target(sA, mD, vB, list(zip(qs[x].tolist(), ans[x].tolist())), int(targets[x]), 1.0)
        t_cpu = time.perf_counter() - t0
        gA, gD, gB = eng.download_kb()
        same = bool(np.array_equal(gA.view(np.uint64), sA.view(np.uint64)) and np.array_equal(gD.view(np.uint64), mD.view(np.uint64))
                    and np.array_equal(gB.view(np.uint64), vB.view(np.uint64)))
        out.update(cpu_oracle_s=t_cpu, cpu_oracle_updates_per_s=n * d / t_cpu, cpu_kind="port (oracle/pqa_oracle.c, 1 thread, one quiz per call)",
                   bit_identical_kb=same)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
