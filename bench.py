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
    "1000x5x1000_b4": dict(Q=1000, K=5, T=1000, B=4),
    "1000x5x1000_b6": dict(Q=1000, K=5, T=1000, B=6),
    "1000x5x1000_b8": dict(Q=1000, K=5, T=1000, B=8),
    "1000x5x1000_b12": dict(Q=1000, K=5, T=1000, B=12),
    "1000x5x1000_b24": dict(Q=1000, K=5, T=1000, B=24),
    "1000x5x1000_b48": dict(Q=1000, K=5, T=1000, B=48),
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


def make_batch(eng, cfg, first_quiz, B):
    """B quizzes in bench.py's shape on `eng` (anything with the batch API): depths DEPTHS cycled, prefixes from synth."""
    states = quiz_states(cfg, first_quiz, B)
    quizzes = eng.start_quiz_batch(B)
    for s in range(max(DEPTHS)):
        sel = [x for x in range(B) if len(states[x]) > s]
        if not sel:
            break
        eng.set_active_question_batch(quizzes[sel], [states[x][s][0] for x in sel])
        eng.record_answer_batch(quizzes[sel], [states[x][s][1] for x in sel])
    return quizzes, states, int(sum(cfg["Q"] - len(pf) for pf in states))


def resident_leg(pqa, eng, steps, warmup, flush, barrier=lambda: None):
    """Device-resident stepping: per-step CUDA events on the engine's stream, L2 flushed between steps outside the events."""
    for _ in range(warmup):
        if flush:
            eng.flush_l2()
        eng.resident_step()
    eng.synchronize()
    launches0 = eng.kernel_launch_count()
    ev = [(pqa.DeviceEvent(), pqa.DeviceEvent()) for _ in range(steps)]
    eval_ms = []
    barrier()
    for k in range(steps):
        if flush:
            eng.flush_l2()
        ev[k][0].record(eng)
        eng.resident_step()
        ev[k][1].record(eng)
        eval_ms.append(eng.resident_last_eval_ms())
    eng.synchronize()
    barrier()
    return dict(total_ms=sum(a.elapsed_ms(b) for a, b in ev), eval_ms=sum(eval_ms) / len(eval_ms),
                launches=eng.kernel_launch_count() - launches0)


def e2e_leg(eng, quizzes, randoms, steps, warmup, barrier=lambda: None):
    for _ in range(warmup):
        eng.next_question_batch(quizzes, randoms)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = eng.next_question_batch(quizzes, randoms)
    return time.perf_counter() - t0, out


def fp64_model(K):
    """fp64 instructions per (quiz, question, answer, target) element of the staged kernel (pqa_eval_staged.cu): pass 1 =
    1 DMUL + 4 Kahan adds; pass 2 = 2 DMUL (post) + 2 DADD (log2 post) + 1 DFMA (H) + DADD + DFMA (V) per element, plus per
    target 2(K-1) (fraction) + 3 (reciprocal) + 2 (lack)."""
    return 5.0 + 7.0 + (2.0 * (K - 1) + 5.0) / K


def fp64_roofline(qevals, K, T, kernel_ms, sm_mhz, sm_count=148, measured_warp_instr=None):
    per_elem = fp64_model(K)
    warp_instr = measured_warp_instr if measured_warp_instr else qevals * K * T * per_elem / 32.0
    peak = 0.5 * 4 * sm_count * (sm_mhz or 1965.0) * 1e6        # one fp64 warp instruction per 2 cycles per sub-partition
    ach = warp_instr / (kernel_ms * 1e-3)
    return {"instr_per_element": per_elem, "warp_instr_per_launch": warp_instr,
            "warp_instr_source": "ncu sm__inst_executed_pipe_fp64.sum" if measured_warp_instr else "model (bench.py fp64_model)",
            "achieved_warp_instr_per_s": ach, "peak_warp_instr_per_s": peak, "frac": ach / peak,
            "peak_source": "0.5 warp-instr/cycle/sub-partition (scripts/microbench/fp64_pipe.cu) x 4 x %d SMs x %.0f MHz" % (sm_count, sm_mhz or 1965.0)}


def measure_traffic(args):
    """DRAM bytes and fp64 warp instructions of ONE launch of the evaluation kernel, from an ncu replay of this same
    workload in a child process (`--traffic-child`: warm-up launches, then the profiled one). None when ncu is missing."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,gpu__time_duration.sum",
           "--clock-control", "none", "-k", "regex:k_eval_staged", "-s", "3", "-c", "1", "--csv",
           sys.executable, os.path.abspath(__file__), "--traffic-child", "--workload", args.workload, "--kernel", str(args.kernel)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=180, cwd=ROOT).stdout
    except Exception:
        return None
    import csv
    vals = {}
    for row in csv.reader(out.splitlines()):
        if len(row) >= 3 and row[-3].startswith(("dram__", "sm__inst", "gpu__time")):
            try:
                v = float(row[-1].replace(",", ""))
            except ValueError:
                continue
            unit = row[-2].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
            vals[row[-3]] = v * scale
    if "dram__bytes_read.sum" not in vals:
        return None
    return {"dram_bytes": vals["dram__bytes_read.sum"] + vals.get("dram__bytes_write.sum", 0.0),
            "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals.get("dram__bytes_write.sum"),
            "fp64_warp_instr": vals.get("sm__inst_executed_pipe_fp64.sum"),
            "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum of one k_eval_staged "
                   "launch of this workload (child process, after 3 warm-up launches; not a timed run)"}


def traffic_child(args, cfg, pqa):
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(cfg["K"], cfg["Q"], cfg["T"], init_amount=INIT),
                                                    emulated_workers=host_cores(), rng_seed=1234, initial_quiz_capacity=cfg["B"])
    eng.fill_binary_search_kb(3)
    eng.set_eval_kernel(args.kernel, args.chunk_targets, args.quizzes_per_cta, args.lanes)
    quizzes, _, _ = make_batch(eng, cfg, 0, cfg["B"])
    eng.resident_bind(quizzes, np.arange(cfg["B"], dtype=np.uint64))
    for _ in range(5):
        eng.resident_step()
    eng.synchronize()


def extra_workload(pqa, name, cores, steps, warmup, device=0):
    """One more BASELINE workload on one GPU, device-resident and through the host-buffer call; returns a dict for the line."""
    cfg = WORKLOADS[name]
    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    t_build = time.perf_counter()
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT), device=device,
                                                    emulated_workers=cores, rng_seed=1234, initial_quiz_capacity=B)
    eng.fill_binary_search_kb(3)
    quizzes, states, qevals = make_batch(eng, cfg, 0, B)
    randoms = np.random.default_rng(99).integers(0, 2 ** 64, size=B, dtype=np.uint64)
    eng.resident_bind(quizzes, randoms)
    t_build = time.perf_counter() - t_build
    r = resident_leg(pqa, eng, steps, warmup, flush=False)     # the KB is far larger than L2
    t_e2e, chosen = e2e_leg(eng, quizzes, randoms, max(2, steps // 2), 1)
    peak, _ = measured_peak()
    alg = qevals * (K + 1) * T * 8
    out = {"workload": name, "Q": Q, "A": K, "T": T, "batch": B, "kb_gb": Q * (K + 1) * T * 8 / 1e9,
           "value": qevals * steps / (r["total_ms"] * 1e-3), "unit": "questions/s", "ms_per_step": r["total_ms"] / steps,
           "steps": steps, "warmup": warmup, "kernel_ms": r["eval_ms"],
           "e2e_value": qevals * max(2, steps // 2) / t_e2e,
           "hbm_algorithmic_gbs": alg / (r["eval_ms"] * 1e-3) / 1e9, "hbm_frac": alg / (r["eval_ms"] * 1e-3) / 1e9 / peak,
           "fp64_frac": fp64_roofline(qevals, K, T, r["eval_ms"], None)["frac"], "setup_s": t_build,
           "l2": "inputs (KB %.1f GB) larger than L2" % (Q * (K + 1) * T * 8 / 1e9),
           "chosen_checksum": int(np.sum(chosen * (np.arange(B) + 1)) % 1000000007)}
    eng.close()
    return out


def train_workload(pqa, cores, cpu_sample_quizzes=4000):
    """BASELINE config 5: 10^6 RecordQuizTarget cell updates = 125 000 quizzes x 8 answered questions on 1000x5x1000 in one
    PqaEngine_RecordQuizTargetBatch call (host buffers), and the CPU port on a bounded sample of the same stream."""
    Q, K, T = 1000, 5, 1000
    n, d = 125000, 8
    kb = synth.binary_search_kb(Q, K, T, INIT, 3)
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT), emulated_workers=cores,
                                                    rng_seed=3, initial_quiz_capacity=n)
    eng.upload_kb(*kb)
    rng = np.random.default_rng(20171126)
    targets = rng.integers(0, T, size=n)
    qs = np.argsort(rng.random((n, 64)), axis=1)[:, :d] + rng.integers(0, Q - 64, size=(n, 1))
    ans = synth.answer_rule(Q, T, K)[qs, targets[:, None]]
    quizzes = eng.start_quiz_batch(n)
    for s in range(d):
        eng.set_active_question_batch(quizzes, qs[:, s])
        eng.record_answer_batch(quizzes, ans[:, s])
    eng.synchronize()
    # the first call of this size also allocates the engine's pinned operation list (40 MB) and sort scratch: it is timed
    # and reported, the value is the second call (same quizzes, same 10^6 updates; RecordQuizTarget leaves the quizzes alive)
    t0 = time.perf_counter()
    eng.record_quiz_target_batch(quizzes, targets)
    eng.synchronize()
    t_first = time.perf_counter() - t0
    launches0 = eng.kernel_launch_count()
    t0 = time.perf_counter()
    eng.record_quiz_target_batch(quizzes, targets)
    eng.synchronize()
    t_gpu = time.perf_counter() - t0
    out = {"workload": "train_1e6_updates_1000x5x1000", "updates": n * d, "quizzes": n, "value": n * d / t_gpu, "unit": "cell updates/s",
           "seconds": t_gpu, "first_call_seconds": t_first,
           "api": "PqaEngine_RecordQuizTargetBatch (host buffers, one call; second call of the size, the first one allocates)",
           "gpu_launches": int(eng.kernel_launch_count() - launches0)}
    try:
        from oracle import oracle as ora
        m = min(cpu_sample_quizzes, n)
        sA, mD, vB = [a.copy() for a in kb]
        t0 = time.perf_counter()
        for x in range(m):
            ora.record_quiz_target(sA, mD, vB, list(zip(qs[x].tolist(), ans[x].tolist())), int(targets[x]), 1.0)
        t_cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m * d / t_cpu, "unit": "cell updates/s", "cores": 1, "kind": "port",
                               "sample": "first %d quizzes of the same stream, one RecordQuizTarget per call (as the reference does)" % m}
    except Exception as ex:
        out["cpu_baseline"] = {"value": None, "kind": "unavailable", "sample": repr(ex)}
    eng.close()
    return out


def sharded_leg(args, wl_name, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks, axis, exchange,
                exact_modes, steps, warmup):
    """Strong scaling of ONE batch over the N GPUs of the job (one process per GPU). axis "targets" (BASELINE config 4's
    partition): every rank holds T/N target columns of every sA/mD row; two-phase evaluation with an exchange of the
    [B][Q][K] W_k partials and of the [B][Q][2K+1] H/V/lack partials. axis "questions": every rank holds Q/N rows, the
    [B][Q] priority columns are exchanged. exchange "p2p": the kernels store into the peers' inboxes over NVLink and a
    device-side barrier orders the phases; "nccl": torch.distributed all-reduce on the engine's buffers. Returns one dict
    per entry of exact_modes (target shards with p2p: False = summed partials, True = exact-order pipeline). Timing: wall
    clock around the public sharded API (host buffers, result D2H inside every step), max over ranks."""
    import torch
    from probqa_b200 import sharded
    cfg = WORKLOADS[wl_name]
    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    fac = pqa.PqaEngineFactory()
    edef = pqa.EngineDefinition(K, Q, T, init_amount=INIT)
    if axis == "questions":
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
    se = (sharded.QuestionShardedEngine([sharded.B200Shard(eng)], group=group) if axis == "questions"
          else sharded.TargetShardedEngine([sharded.B200TargetShard(eng)], group=group))
    if exchange == "p2p":
        se.enable_p2p(B, exact_order=False)
    quizzes, states, qevals_step = make_batch(se, cfg, 0, B)
    randoms = np.random.default_rng(99).integers(0, 2 ** 64, size=B, dtype=np.uint64)
    results = []
    for exact in exact_modes:
        if exchange == "p2p" and axis == "targets":
            eng.p2p_set_exact_order(bool(exact))
        for _ in range(warmup):
            chosen = se.next_question_batch(quizzes, randoms)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = eng.kernel_launch_count()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            chosen = se.next_question_batch(quizzes, randoms)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        launches = eng.kernel_launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        assert np.all((chosen >= 0) & (chosen < Q))
        phases = None
        if exchange == "p2p" and axis == "targets":
            ph = eng.p2p_last_phase_ms()
            names = ["phase1_W_partials", "barrier1", "phase2_HVL_partials", "barrier2", "epilogue_select"]
            phases = {n_: max_over_ranks(float(v)) for n_, v in zip(names, ph)}
            phases["rank0"] = {n_: float(v) for n_, v in zip(names, ph)}
            phases["note"] = ("device ms of the last timed step, CUDA events on each rank's stream, max over ranks; with the "
                              "exact-order pipeline phase 2 includes waiting for the last shard's W_k tiles and there is no barrier1")
        value = qevals_step * steps / dt
        peak, peak_src = measured_peak()
        xbytes = int(B * Q * 8) if axis == "questions" else int(B * Q * (3 * K + 1) * 8)
        if exact and axis == "targets":
            xbytes = int(B * Q * (K * 8 * 8 + K * 8 + (2 * K + 1) * 8))       # Kahan lanes hand-over + complete W_k + H/V/lack partials
        per_gpu = qevals_step * (K + 1) * T * 8 / (dt / steps) / 1e9 / world
        results.append({
            "workload": wl_name, "Q": Q, "A": K, "T": T, "batch_total": B, "n_gpus": world, "axis": axis, "scaling": "strong",
            "exchange": ("NCCL all-reduce via torch.distributed" if exchange == "nccl" else
                         "peer-memory stores from the kernel epilogues over NVLink + device-side barrier (no host round trip)"),
            "exact_order_pipeline": bool(exact and axis == "targets" and exchange == "p2p"),
            "value": value, "unit": "questions/s", "ms_per_step": 1e3 * dt / steps, "steps": steps, "warmup": warmup,
            "exchange_bytes_per_step_per_gpu": xbytes, "kb_shard_gb": shard_bytes / 1e9,
            "hbm_algorithmic_gbs_per_gpu": per_gpu, "hbm_frac_per_gpu": per_gpu / peak, "phases_ms": phases,
            "gpu_launches_per_step_per_gpu": launches / steps, "clocks": clocks,
            "timing": "wall clock around sharded.%s.next_question_batch (host buffers; result D2H + host sync inside every step), max over ranks"
                      % type(se).__name__,
            "chosen_checksum": int(np.sum(chosen * (np.arange(B) + 1)) % 1000000007)})
    eng.close()
    return results


def bench_sharded(args, cfg, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks):
    """`--shard questions|targets` as the job's main line (scripts/run_multi_gpu.sh): value == e2e."""
    res = sharded_leg(args, args.workload, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks, args.shard,
                      args.exchange, [args.exact_order], args.steps, args.warmup)[0]
    if rank == 0:
        peak, peak_src = measured_peak()
        line = {"metric": METRIC, "value": res["value"], "unit": "questions/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "Q": cfg["Q"], "A": cfg["K"], "T": cfg["T"], "batch_total": cfg["B"],
                           "quiz_depths": list(DEPTHS), "kb": "binary_search_kb(init=0.1, rounds=3), filled on the device",
                           "parallelism": "%s sharded over %d GPUs" % (res["axis"], world), "exchange": res["exchange"],
                           "exact_order_pipeline": res["exact_order_pipeline"], "timing": res["timing"],
                           "chosen_checksum": res["chosen_checksum"]},
                "e2e": {"value": res["value"], "unit": "questions/s", "h2d_bytes_per_step": int(cfg["B"] * 16),
                        "d2h_bytes_per_step": int(cfg["B"] * 8), "exchanged_bytes_per_step_per_gpu": res["exchange_bytes_per_step_per_gpu"]},
                "gpu_launches": int(res["gpu_launches_per_step_per_gpu"] * args.steps), "clocks": res["clocks"],
                "phases_ms": res["phases_ms"],
                "roofline": {"bound": "hbm", "achieved": res["hbm_algorithmic_gbs_per_gpu"], "peak": peak, "unit": "GB/s",
                             "frac": res["hbm_frac_per_gpu"], "traffic": None, "peak_source": peak_src,
                             "note": "per GPU, whole step (evaluation + exchange + selection), algorithmic bytes"}}
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
    quizzes, states, qevals_step = make_batch(eng, cfg, 0, B)
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
                    help="N>1 main line: quizzes = KB replicated, batch sharded, no collective (default; the sharded BASELINE "
                         "config 4 leg is then embedded under \"sharded\"); questions / targets = that split as the main line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra legs of the default line (BASELINE configs 3, 4, 5, ncu traffic, sharded config 4)")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
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
    from probqa_b200 import engine as pqa
    if args.traffic_child:
        traffic_child(args, cfg, pqa)
        return
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

    Q, K, T, B = cfg["Q"], cfg["K"], cfg["T"], cfg["B"]
    cores = host_cores()
    if args.group > 1:
        return bench_group(args, cfg, pqa, cores)
    if args.shard in ("questions", "targets") and world > 1:
        return bench_sharded(args, cfg, pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks)
    extras = not args.no_extras and args.workload == "1000x5x1000_b256" and args.kernel != 1
    eng = pqa.PqaEngineFactory().create_b200_engine(pqa.EngineDefinition(K, Q, T, init_amount=INIT), device=local_rank,
                                                    emulated_workers=cores, rng_seed=1234 + rank, initial_quiz_capacity=B)
    if Q * K * T * 8 > (1 << 30):
        eng.fill_binary_search_kb(3)     # the same KB bit for bit (tests/test_gpu_sharded.py), written on the device
    else:
        eng.upload_kb(*synth.binary_search_kb(Q, K, T, INIT, 3))
    eng.set_eval_kernel(args.kernel, args.chunk_targets, args.quizzes_per_cta, args.lanes)
    # this rank's shard of the batch: quizzes rank*B .. rank*B+B-1 (weak scaling: B per GPU)
    quizzes, states, qevals_step = make_batch(eng, cfg, rank * B, B)
    rng = np.random.default_rng(99 + rank)
    randoms = rng.integers(0, 2 ** 64, size=B, dtype=np.uint64)

    # ---------------- device-resident leg (value, roofline)
    eng.resident_bind(quizzes, randoms)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # nvidia-smi needs ~100 ms per sample: it runs from the warm-up to the end of the timed steps
        time.sleep(0.4)
    r = resident_leg(pqa, eng, args.steps, args.warmup, not args.no_flush, barrier)
    total_ms = max_over_ranks(r["total_ms"])
    launches = r["launches"]
    chosen = eng.resident_fetch()
    assert np.all((chosen >= 0) & (chosen < Q))
    total_qevals = sum_over_ranks(qevals_step)
    value = total_qevals * args.steps / (total_ms * 1e-3)
    eval_ms_avg = max_over_ranks(r["eval_ms"])

    # ---------------- end-to-end leg through the C-ABI batch call with host buffers
    t_e2e, out = e2e_leg(eng, quizzes, randoms, args.steps, args.warmup, barrier)
    t_e2e = max_over_ranks(t_e2e)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
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
        peak1, _ = measured_peak()
        single = {"api": "PqaEngine_NextQuestion (one quiz per call)", "calls_per_s": n_calls / dt1, "us_per_call": 1e6 * dt1 / n_calls,
                  "questions_per_s": (Q - len(states[0])) * n_calls / dt1,
                  "hbm_frac": (Q - len(states[0])) * (K + 1) * T * 8 / (dt1 / n_calls) / 1e9 / peak1,
                  "hbm_frac_note": "algorithmic bytes of one call ((K+1)*T*8 per unasked question) / wall time of the whole call / HBM peak"}
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
    eng.close()

    # ---------------------------------------------------------------- extra legs (default line only)
    sharded_res = config4_1gpu = None
    if extras and world > 1:
        # BASELINE config 4 as north_star partitions it: targets sharded over the job's GPUs, peer-memory exchange; the
        # exact-order pipeline (single-engine parity bar) first, the summed-partials exchange beside it
        sharded_res = sharded_leg(args, "10000x5x100000_b64", pqa, dist, rank, world, local_rank, cores, barrier, max_over_ranks,
                                  "targets", "p2p", [True, False], 10, 5)
    if extras and rank == 0:
        # the same workload on ONE GPU (the whole 48 GB KB + its derived form on one B200): the strong-scaling base, measured
        # in the same job on the same box
        config4_1gpu = extra_workload(pqa, "10000x5x100000_b64", cores, 5, 3, device=local_rank)
        if sharded_res:
            for s_ in sharded_res:
                s_["one_gpu_same_job"] = {"value": config4_1gpu["value"], "ms_per_step": config4_1gpu["ms_per_step"],
                                          "e2e_value": config4_1gpu["e2e_value"]}
                s_["speedup_vs_1gpu"] = s_["value"] / config4_1gpu["e2e_value"]
                s_["speedup_note"] = "sharded value (host-buffer API, wall clock) / one-GPU e2e_value (host-buffer API) of the same workload, same job"
    barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    traffic = measure_traffic(args) if (extras and world == 1) else None
    alg_bytes = qevals_step * (K + 1) * T * 8           # per launch of the evaluation kernel on one GPU
    achieved = alg_bytes / (eval_ms_avg * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz")
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
                     "traffic": traffic["dram_bytes"] if traffic else None,
                     "traffic_detail": traffic,
                     "kernel": "k_eval_staged" if args.kernel != 1 else "k_eval_exact",
                     "kernel_ms": eval_ms_avg, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "fp64": fp64_roofline(qevals_step, K, T, eval_ms_avg, sm_mhz,
                                           measured_warp_instr=(traffic or {}).get("fp64_warp_instr")) if args.kernel != 1 else None,
                     "note": "algorithmic bytes are counted per quiz ((K+1)*T*8 per evaluated question, SURVEY 8d); the slab is "
                             "staged once per CTA and shared by its quizzes, so physical DRAM traffic (`traffic`, one ncu replay "
                             "of this workload) is far lower and the kernel is bound by fp64 issue: `fp64` is the binding roofline"},
    }
    if single is not None:
        line["single_quiz"] = single
    if sharded_res is not None:
        line["sharded"] = sharded_res[0]
        line["sharded_summed_partials"] = sharded_res[1]
    if config4_1gpu is not None:
        line["config4_1gpu"] = config4_1gpu
    if extras and world == 1:
        try:
            line["config3"] = extra_workload(pqa, "10000x5x10000_b1024", cores, 5, 3)
        except Exception as ex:
            line["config3"] = {"error": repr(ex)}
        try:
            line["config5_train"] = train_workload(pqa, cores)
        except Exception as ex:
            line["config5_train"] = {"error": repr(ex)}
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
