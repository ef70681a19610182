#!/usr/bin/env python
"""bench.py -- questions-evaluated/s of NextQuestion (BASELINE.json metric) on synthetic Q x A x T knowledge bases.

A "step" is one NextQuestion pass (question evaluation + selection) over one batch of concurrent quizzes.
  value        : whole-job questions-evaluated/s with everything resident in HBM (device-resident stepping, CUDA events
                 on the engine's stream, L2 flushed between steps, max over ranks)
  e2e          : the same metric through the C-ABI batch call PqaEngine_NextQuestionBatch with HOST buffers (quiz ids and
                 random draws copied H2D, chosen questions copied D2H inside the timed region)
  roofline     : algorithmic bytes of the question-evaluation kernel ((K+1)*T*8 per evaluated question)
                 / its device time (events around that kernel inside the timed steps) vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference's own AVX2 CpuEngine code (oracle/_ref, all host cores) on a bounded sample
`--impl reference` times only that CPU arm. N>1: one process per GPU (torchrun), the batch of quizzes is sharded over
the ranks (quizzes are independent; the KB is replicated), no data-path collective; weak scaling.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from probqa_b200 import synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: 1000Q x 5A x 1000T, batch = 256 concurrent quizzes, 1 x B200
    "1000x5x1000_b256": dict(Q=1000, K=5, T=1000, B=256),
    # configs[2]: 10000Q x 5A x 10000T, batch = 1024 (4.8 GB KB)
    "10000x5x10000_b1024": dict(Q=10000, K=5, T=10000, B=1024),
    # configs[3]: 10000Q x 5A x 100000T (48 GB KB), sharded across the GPUs of one box (--shard targets / questions)
    "10000x5x100000_b64": dict(Q=10000, K=5, T=100000, B=64),
    "10000x5x100000_b256": dict(Q=10000, K=5, T=100000, B=256),
    "2000x5x20000_b64": dict(Q=2000, K=5, T=20000, B=64),
    "1000x5x1000_b8": dict(Q=1000, K=5, T=1000, B=8),
    "1000x5x1000_b16": dict(Q=1000, K=5, T=1000, B=16),
    "1000x5x1000_b32": dict(Q=1000, K=5, T=1000, B=32),
    "1000x5x1000_b64": dict(Q=1000, K=5, T=1000, B=64),
    "1000x5x1000_b128": dict(Q=1000, K=5, T=1000, B=128),
    # small smoke size
    "200x5x500_b32": dict(Q=200, K=5, T=500, B=32),
}
DEPTHS = (0, 3, 8)   # answered questions per quiz, cycled over the batch (SURVEY 8d "Quizzes")
INIT = 0.1
METRIC = "questions-evaluated/sec (NextQuestion)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def quiz_states(cfg, first_quiz, n):
    """[(depth, prefix)] for quizzes first_quiz .. first_quiz+n-1."""
    out = []
    for b in range(first_quiz, first_quiz + n):
        d = DEPTHS[b % len(DEPTHS)]
        out.append(synth.quiz_prefix(b, d, cfg["Q"], cfg["T"], cfg["K"]))
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def run_reference(args, cfg, wl_name, rank, world):
    """The reference's own CpuEngine code on the host cores (oracle/_ref), else the C port (oracle/)."""
    if rank != 0:
        return
    cores = host_cores()
    kb = synth.binary_search_kb(cfg["Q"], cfg["K"], cfg["T"], INIT, 3)
    from oracle import ref as R
    kind = "reference" if R.available() else "port"
    if kind == "reference":
        eng = R.RefEngine(kb[0], kb[1], kb[2], nWorkers=cores, osThreads=cores)
    else:
        from oracle import oracle as O
    n_sample = args.ref_quizzes
    states = quiz_states(cfg, 0, n_sample)
    priors, askeds = [], []
    for prefix in states:
        if kind == "reference":
            p = eng.start_quiz()
            for q, a in prefix:
                p = eng.record_answer(p, q, a)
        else:
            p = O.start_quiz(kb[2], cores)
            for q, a in prefix:
                p = O.record_answer(p, kb[0][q, a], kb[1][q], max(1, cores - 1))
        asked = np.zeros(cfg["Q"], dtype=bool)
        asked[[q for q, _ in prefix]] = True
        priors.append(p); askeds.append(asked)
    qevals_step = int(sum(cfg["Q"] - len(pf) for pf in states))

    def step():
        for p, a in zip(priors, askeds):
            if kind == "reference":
                eng.eval_questions(p, a)
            else:
                O.eval_questions(kb[0], kb[1], p, cores, asked=a, nThreads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = qevals_step * args.steps / dt
    sample = "%d quizzes (depths %s) x all unasked questions per step, one NextQuestion evaluation at a time, %d threads" % (
        n_sample, "/".join(map(str, DEPTHS)), cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "questions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "Q": cfg["Q"], "A": cfg["K"], "T": cfg["T"], "kb": "binary_search_kb(init=0.1, rounds=3)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "questions/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(cfg, budget_s=12.0):
    cores = host_cores()
    kb = synth.binary_search_kb(cfg["Q"], cfg["K"], cfg["T"], INIT, 3)
    from oracle import ref as R
    kind = "reference" if R.available() else "port"
    states = quiz_states(cfg, 0, 6)
    if kind == "reference":
        eng = R.RefEngine(kb[0], kb[1], kb[2], nWorkers=cores, osThreads=cores)
        prep = []
        for prefix in states:
            p = eng.start_quiz()
            for q, a in prefix:
                p = eng.record_answer(p, q, a)
            asked = np.zeros(cfg["Q"], dtype=bool); asked[[q for q, _ in prefix]] = True
            prep.append((p, asked))
        run = lambda p, a: eng.eval_questions(p, a)
    else:
        from oracle import oracle as O
        prep = []
        for prefix in states:
            p = O.start_quiz(kb[2], cores)
            for q, a in prefix:
                p = O.record_answer(p, kb[0][q, a], kb[1][q], max(1, cores - 1))
            asked = np.zeros(cfg["Q"], dtype=bool); asked[[q for q, _ in prefix]] = True
            prep.append((p, asked))
        run = lambda p, a: O.eval_questions(kb[0], kb[1], p, cores, asked=a, nThreads=cores)
    run(*prep[0])  # warm-up
    n_calls, qevals = 0, 0
    t0 = time.perf_counter()
    while True:
        p, a = prep[n_calls % len(prep)]
        run(p, a)
        qevals += int(cfg["Q"] - a.sum()); n_calls += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n_calls >= 4096:
            break
    return {"value": qevals / dt, "unit": "questions/s", "cores": cores, "kind": kind,
            "sample": "%d NextQuestion evaluations (quiz depths %s, all unasked questions each) in %.1f s, %d threads" % (
                n_calls, "/".join(map(str, DEPTHS)), dt, cores)}


def bench_sharded(args, cfg, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks):
    """Strong scaling of ONE batch over N GPUs. --shard questions: every rank holds Q/N question rows of the KB, the
    ranks' priority columns are exchanged, every rank selects (bit-identical to one engine). --shard targets (BASELINE
    config 4's axis): every rank holds T/N target columns, two-phase evaluation with an exchange of the W_k partials and
    of the H/V/lack partials. --exchange nccl: torch.distributed all-reduce on the engine's buffers (host sync on both
    sides); --exchange p2p: the kernels store into the peers' inboxes over NVLink and a device-side barrier orders the
    phases. The whole step goes through the public sharded API with host buffers, so value == e2e here."""
    import torch
    from probqa_b200 import sharded
    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    if args.shard == "questions":
        first, count = sharded.shard_ranges(Q, world)[rank]
        eng = fac.create_b200_engine(edef, device=local_rank, emulated_workers=cores, rng_seed=1234, initial_quiz_capacity=B,
                                     question_shard_first=first, question_shard_count=count)
        shard_bytes = count * (K + 1) * T * 8
    else:
        first, count = sharded.target_shard_ranges(T, world)[rank]
        eng = fac.create_b200_engine(edef, device=local_rank, emulated_workers=cores, rng_seed=1234, initial_quiz_capacity=B,
                                     target_shard_first=first, target_shard_count=count)
        shard_bytes = Q * (K + 1) * count * 8
    eng.fill_binary_search_kb(3)       # this rank's shard of synth.binary_search_kb, written on the device
    eng.set_eval_kernel(args.kernel, args.chunk_targets, args.quizzes_per_cta, args.lanes)
    group = dist.group.WORLD if dist is not None else None
    if args.shard == "questions":
        se = sharded.QuestionShardedEngine([sharded.B200Shard(eng)], group=group)
    else:
        se = sharded.TargetShardedEngine([sharded.B200TargetShard(eng)], group=group)
    if args.exchange == "p2p":
        se.enable_p2p(B, exact_order=args.exact_order and args.shard == "targets")
    states = quiz_states(cfg, 0, B)
    quizzes = se.start_quiz_batch(B)
    for s in range(max(DEPTHS)):
        sel = [x for x in range(B) if len(states[x]) > s]
        if not sel:
            break
        se.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
        se.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
    qevals_step = int(sum(Q - len(pf) for pf in states))
    randoms = np.random.default_rng(99).integers(0, 2 ** 64, size=B, dtype=np.uint64)
    for _ in range(args.warmup):
        chosen = se.next_question_batch(quizzes, randoms)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chosen = se.next_question_batch(quizzes, randoms)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    launches = eng.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    assert np.all((chosen >= 0) & (chosen < Q))
    if rank != 0:
        return
    value = qevals_step * args.steps / dt
    peak, peak_src = measured_peak()
    if args.shard == "questions":
        par = "questions sharded over %d GPUs (Q/N rows of sA/mD each), exchange of the [B][Q] priority columns" % world
        xbytes = int(B * Q * 8)
    else:
        par = ("targets sharded over %d GPUs (T/N columns of every sA/mD row each), two-phase evaluation: exchange of the "
               "[B][Q][K] W_k partials, then of the [B][Q][2K+1] H/V/lack partials" % world)
        xbytes = int(B * Q * (3 * K + 1) * 8)
    per_gpu = qevals_step * (K + 1) * T * 8 / (dt / args.steps) / 1e9 / world
    line = {
        "metric": METRIC, "value": value, "unit": "questions/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "Q": Q, "A": K, "T": T, "batch_total": B, "quiz_depths": list(DEPTHS),
                   "kb": "binary_search_kb(init=0.1, rounds=3), filled on the device", "parallelism": par,
                   "exchange": ("NCCL all-reduce via torch.distributed (host sync on both sides)" if args.exchange == "nccl" else
                                "peer-memory stores from the kernel epilogues + device-side barrier (no host round trip)" +
                                ("; exact-order pipeline of the Kahan lanes (W_k bit-exact across shards)"
                                 if args.exact_order and args.shard == "targets" else "")),
                   "l2": "inputs re-read every step; KB shard %.1f MB per GPU" % (shard_bytes / 1e6),
                   "timing": "wall clock around the public sharded API (result D2H + host sync inside every step), max over ranks",
                   "chosen_checksum": int(np.sum(chosen * (np.arange(B) + 1)) % 1000000007)},
        "e2e": {"value": value, "unit": "questions/s", "h2d_bytes_per_step": int(B * 16), "d2h_bytes_per_step": int(B * 8),
                "api": "sharded.%s.next_question_batch" % type(se).__name__, "exchanged_bytes_per_step_per_gpu": xbytes},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak, "traffic": None,
                     "peak_source": peak_src, "note": "per GPU, whole step (evaluation + exchange + selection), algorithmic bytes"},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def bench_group(args, cfg, pqa, cores):
    """One process, one engine handle over --group GPUs (ShardGroup): the same NextQuestion batch through the ordinary
    PqaEngine_NextQuestionBatch call. Strong scaling; value == e2e (host buffers, wall clock)."""
    import torch
    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    axis = args.shard if args.shard in ("questions", "targets") else "targets"
    have = torch.cuda.device_count()
    eng = pqa.PqaEngineFactory().create_sharded_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT), axis, args.group,
                                                       devices=[r % have for r in range(args.group)], exact_order=args.exact_order,
                                                       max_batch=B, emulated_workers=cores, rng_seed=1234, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    states = quiz_states(cfg, 0, B)
    quizzes = eng.start_quiz_batch(B)
    for s in range(max(DEPTHS)):
        sel = [x for x in range(B) if len(states[x]) > s]
        if not sel:
            break
        eng.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
        eng.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
    qevals_step = int(sum(Q - len(pf) for pf in states))
    randoms = np.random.default_rng(99).integers(0, 2 ** 64, size=B, dtype=np.uint64)
    for _ in range(args.warmup):
        chosen = eng.next_question_batch(quizzes, randoms)
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = eng.kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        chosen = eng.next_question_batch(quizzes, randoms)
    dt = time.perf_counter() - t0
    launches = eng.kernel_launch_count() - launches0
    clocks = sampler.stop()
    assert np.all((chosen >= 0) & (chosen < Q))
    value = qevals_step * args.steps / dt
    peak, peak_src = measured_peak()
    per_gpu = qevals_step * (K + 1) * T * 8 / (dt / args.steps) / 1e9 / min(args.group, have)
    line = {
        "metric": METRIC, "value": value, "unit": "questions/s", "n_gpus": min(args.group, have), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "Q": Q, "A": K, "T": T, "batch_total": B, "quiz_depths": list(DEPTHS),
                   "kb": "binary_search_kb(init=0.1, rounds=3), filled on the device",
                   "parallelism": "ONE process, one engine handle (ShardGroup) over %d shards on %d GPU(s), %s sharded, peer-memory exchange%s"
                                  % (args.group, min(args.group, have), axis, ", exact-order pipeline" if args.exact_order and axis == "targets" else ""),
                   "timing": "wall clock around PqaEngine_NextQuestionBatch on the group handle (host buffers)",
                   "chosen_checksum": int(np.sum(chosen * (np.arange(B) + 1)) % 1000000007)},
        "e2e": {"value": value, "unit": "questions/s", "h2d_bytes_per_step": int(B * 16 * args.group), "d2h_bytes_per_step": int(B * 8 * args.group),
                "api": "PqaEngine_NextQuestionBatch (group handle)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak, "traffic": None,
                     "peak_source": peak_src, "note": "per GPU, whole step, algorithmic bytes"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1000x5x1000_b256", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", type=int, default=0, help="0 auto (staged), 1 exact, 2 staged")
    ap.add_argument("--quizzes-per-cta", type=int, default=0)
    ap.add_argument("--chunk-targets", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="Kahan lanes per thread of the staged kernel: 0 auto, 1, 4")
    ap.add_argument("--ref-quizzes", type=int, default=8, help="quizzes per step of the reference arm")
    ap.add_argument("--group", type=int, default=0,
                    help="single process: one sharded engine group over this many GPUs (PqaB200_CreateShardedEngine), axis = --shard")
    ap.add_argument("--exact-order", action="store_true",
                    help="--shard targets --exchange p2p: hand the Kahan lanes from shard to shard (W_k bit-exact across shards)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="sharded modes: how the shards exchange partial results (peer memory from the kernels, or NCCL)")
    ap.add_argument("--shard", default="quizzes", choices=["quizzes", "questions", "targets"],
                    help="N>1: quizzes = KB replicated, batch sharded, no collective; questions = each rank holds Q/N "
                         "questions, NCCL all-reduce of the per-question priorities (probqa_b200/sharded.py)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, args.workload, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from probqa_b200 import engine as pqa
    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    cores = host_cores()
    if args.group > 1:
        return bench_group(args, cfg, pqa, cores)
    if args.shard in ("questions", "targets") and world > 1:
        return bench_sharded(args, cfg, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks)
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT), device=local_rank,
                                                    emulated_workers=cores, rng_seed=1234 + rank, initial_quiz_capacity=B)
    if Q * K * T * 8 > (1 << 30):
        eng.fill_binary_search_kb(3)     # the same KB bit for bit (tests/test_gpu_sharded.py), written on the device
    else:
        eng.upload_kb(*synth.binary_search_kb(Q, K, T, INIT, 3))
    eng.set_eval_kernel(args.kernel, args.chunk_targets, args.quizzes_per_cta, args.lanes)
    # this rank's shard of the batch: quizzes rank*B .. rank*B+B-1 (weak scaling: B per GPU)
    states = quiz_states(cfg, rank * B, B)
    quizzes = eng.start_quiz_batch(B)
    for s in range(max(DEPTHS)):
        sel = [x for x in range(B) if len(states[x]) > s]
        if not sel:
            break
        eng.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
        eng.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
    qevals_step = int(sum(Q - len(pf) for pf in states))
    rng = np.random.default_rng(99 + rank)
    randoms = rng.integers(0, 2 ** 64, size=B, dtype=np.uint64)

    # ---------------- device-resident leg (value, roofline)
    eng.resident_bind(quizzes, randoms)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # nvidia-smi needs ~100 ms per sample: it runs from the warm-up to the end of the timed steps
    for _ in range(args.warmup):
        if not args.no_flush:
            eng.flush_l2()
        eng.resident_step()
    eng.synchronize()
    launches0 = eng.kernel_launch_count()
    ev = [(pqa.DeviceEvent(), pqa.DeviceEvent()) for _ in range(args.steps)]
    eval_ms = []
    barrier()
    for k in range(args.steps):
        if not args.no_flush:
            eng.flush_l2()
        ev[k][0].record(eng)
        eng.resident_step()
        ev[k][1].record(eng)
        eval_ms.append(eng.resident_last_eval_ms())
    eng.synchronize()
    barrier()
    launches = eng.kernel_launch_count() - launches0
    step_ms = [a.elapsed_ms(b) for a, b in ev]
    total_ms = max_over_ranks(sum(step_ms))
    clocks = sampler.stop() if rank == 0 else None
    chosen = eng.resident_fetch()
    assert np.all((chosen >= 0) & (chosen < Q))
    total_qevals = sum_over_ranks(qevals_step)
    value = total_qevals * args.steps / (total_ms * 1e-3)
    eval_ms_avg = max_over_ranks(sum(eval_ms) / len(eval_ms))

    # ---------------- end-to-end leg through the C-ABI batch call with host buffers
    for _ in range(args.warmup):
        eng.next_question_batch(quizzes, randoms)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = eng.next_question_batch(quizzes, randoms)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert np.array_equal(out, chosen), "resident and host-buffer paths chose different questions"
    e2e_value = total_qevals * args.steps / t_e2e

    # ---------------------------------------------------------------- BASELINE configs[0]: one quiz through the reference's one-quiz-per-call entry point
    single = None
    if rank == 0 and world == 1 and Q * K * T * 8 <= (1 << 30):
        n_calls = 200
        q1 = int(quizzes[0])
        for _ in range(20):
            eng.next_question(q1)
        t0 = time.perf_counter()
        for _ in range(n_calls):
            eng.next_question(q1)
        dt1 = time.perf_counter() - t0
        single = {"api": "PqaEngine_NextQuestion (one quiz per call)", "calls_per_s": n_calls / dt1, "us_per_call": 1e6 * dt1 / n_calls,
                  "questions_per_s": (Q - len(states[0])) * n_calls / dt1}
        # the same entry point from many client threads at once (the engine combines pending calls into one launch)
        import threading
        n_thr, per_thr = min(64, B), 100
        def client(x):
            qx = int(quizzes[x])
            for _ in range(per_thr):
                eng.next_question(qx)
        thr = [threading.Thread(target=client, args=(x,)) for x in range(n_thr)]
        t0 = time.perf_counter()
        for t in thr:
            t.start()
        for t in thr:
            t.join()
        dtm = time.perf_counter() - t0
        single.update(threads=n_thr, threaded_calls_per_s=n_thr * per_thr / dtm,
                      threaded_questions_per_s=sum(Q - len(states[x]) for x in range(n_thr)) * per_thr / dtm)

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    traffic = None   # physical DRAM bytes per launch from the committed ncu --set full capture of this workload
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_eval_traffic.json")))
        if tj["workload"] == args.workload and args.kernel != 1:
            traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    except Exception:
        pass
    alg_bytes = qevals_step * (K + 1) * T * 8           # per launch of the evaluation kernel on one GPU
    achieved = alg_bytes / (eval_ms_avg * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "questions/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "Q": Q, "A": K, "T": T, "batch_per_gpu": B, "quiz_depths": list(DEPTHS),
                   "kb": "binary_search_kb(init=0.1, rounds=3)", "parallelism": "quizzes sharded over %d GPU(s), KB replicated" % world,
                   "l2": "not flushed" if args.no_flush else "flushed between steps (256 MB write outside the per-step events)",
                   "timing": "per-step CUDA events on the engine stream, summed; max over ranks",
                   "kernel": {0: "staged", 1: "exact", 2: "staged"}[args.kernel],
                   "emulated_workers": cores, "next_question_calls_per_s": value / (qevals_step / B)},
        "e2e": {"value": e2e_value, "unit": "questions/s", "h2d_bytes_per_step": int(B * 16), "d2h_bytes_per_step": int(B * 8),
                "api": "PqaEngine_NextQuestionBatch (host buffers)", "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_eval_staged" if args.kernel != 1 else "k_eval_exact",
                     "kernel_ms": eval_ms_avg, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "note": "algorithmic bytes are counted per quiz ((K+1)*T*8 per evaluated question, SURVEY 8d); the slab is "
                             "staged once per CTA and shared by its quizzes, so physical DRAM traffic is far lower and the "
                             "kernel is bound by fp64 issue, see DESIGN.md"},
    }
    if single is not None:
        line["single_quiz"] = single
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(cfg)
        except Exception as ex:  # the checker missing must not lose the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "questions/s", "cores": cores, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
