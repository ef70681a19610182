"""Seeded synthetic knowledge bases and quiz states for the parity tests and bench.py.

Three KBs, all in the reference's file layout (CpuEngine.cpp:664-688): sA[Q,K,T] squared counts, mD[Q,T] = sum_k sA,
vB[T].
  * binary_search_kb: closed form, mirrors the answer rule of the reference's DichotomyTest
    (ProbQA/PqaCoreTests/DichotomyTest.cpp:50-67) after `rounds` rounds of training with amount 1;
  * gamma_kb: sA = (init + Gamma(0.5, 1))^2 from a fixed seed, vB = init + U(0, 10);
  * uniform_kb: the untrained cube (all init^2) -- every prior ties, the ListTopTargets stress case.
Nothing here touches the GPU or the oracle; numpy only.
"""
import numpy as np


def answer_rule(Q, T, K=5):
    """ans[i, j] in 0..4: the answer a perfect user gives to question i when the hidden target is j."""
    w = max(1, (32 * T) // 1000)
    piv = (np.arange(Q, dtype=np.int64) * T) // Q
    j = np.arange(T, dtype=np.int64)[None, :]
    p = piv[:, None]
    ans = np.where(j < p - w, 0, np.where(j < p, 1, np.where(j == p, 2, np.where(j <= p + w, 3, 4))))
    return np.minimum(ans, K - 1).astype(np.int64)


def binary_search_kb(Q, K, T, init=0.1, rounds=3):
    ans = answer_rule(Q, T, K)
    cnt = np.full((Q, K, T), init, dtype=np.float64)
    k = np.arange(K, dtype=np.int64)[None, :, None]
    cnt = cnt + rounds * (k == ans[:, None, :])
    sA = cnt * cnt
    mD = np.zeros((Q, T), dtype=np.float64)
    for kk in range(K):
        mD += sA[:, kk, :]
    vB = np.full(T, init + rounds, dtype=np.float64)
    return np.ascontiguousarray(sA), mD, vB


def gamma_kb(Q, K, T, init=0.1, seed=20171126):
    rng = np.random.Generator(np.random.MT19937(seed))
    cnt = init + rng.gamma(0.5, 1.0, size=(Q, K, T))
    sA = cnt * cnt
    mD = np.zeros((Q, T), dtype=np.float64)
    for kk in range(K):
        mD += sA[:, kk, :]
    vB = init + rng.uniform(0.0, 10.0, size=T)
    return np.ascontiguousarray(sA), mD, vB


def uniform_kb(Q, K, T, init=0.1):
    sA = np.full((Q, K, T), init * init, dtype=np.float64)
    mD = np.full((Q, T), K * (init * init), dtype=np.float64)
    vB = np.full(T, init, dtype=np.float64)
    return sA, mD, vB


def hidden_target(b, T):
    return (int(b) * 2654435761) % T


def quiz_prefix(b, depth, Q, T, K=5):
    """The first `depth` (question, answer) pairs of synthetic quiz b: questions from a fixed LCG (no repeats),
    answers from answer_rule for the quiz' hidden target."""
    t = hidden_target(b, T)
    w = max(1, (32 * T) // 1000)
    out, seen = [], set()
    x = (1103515245 * (int(b) + 12345) + 12345) & 0x7FFFFFFF
    while len(out) < depth:
        x = (1103515245 * x + 12345) & 0x7FFFFFFF
        q = x % Q
        if q in seen:
            continue
        seen.add(q)
        p = (q * T) // Q
        a = 0 if t < p - w else 1 if t < p else 2 if t == p else 3 if t <= p + w else 4
        out.append((q, min(a, K - 1)))
    return out
