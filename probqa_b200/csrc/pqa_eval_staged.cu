// probqa_b200: the throughput question-evaluation kernel (sm_100a).
//
// Computes what CEEvalQsSubtaskConsider<SRDoubleNumber>::Run (CEEvalQsSubtaskConsider.cpp:41-217) computes for
// every (quiz, question) pair of a batch of concurrent quizzes, restructured for the GPU:
//
//  * One CTA owns one question i and a tile of quizzes. The question's slab -- K rows of sA[i][k][.] and the row
//    mD[i][.] -- is staged into shared memory with 1-D bulk-async copies (TMA engine, cp.async.bulk + mbarrier) and
//    transformed in place once per CTA: invD = 1/mD, r[k][j] = sA*invD (the reference's per-target likelihood factor,
//    :72-81), lr[k][j] = log2 r, id2[j] = invD^2 (the numerator of the "lack" term, :116-117). Every quiz of the tile
//    then reuses the staged slab, so HBM/L2 sees the slab once per CTA instead of once per quiz.
//  * The reference sums with a 4-lane AVX2 Kahan accumulator: target j goes to lane j%4, lanes advance in vector
//    order (SRAccumVectDbl256.h:40-46). A GPU thread here owns KL of those 4 Kahan lanes of ONE quiz (KL = 4: one
//    thread per quiz, 32 quizzes per warp, used for big batches; KL = 1: four threads per quiz, 8 quizzes per warp,
//    used for small batches) and walks the 4-target vectors in order, so pass 1 reproduces the reference's normaliser
//    W_k BIT FOR BIT (4-lane Kahan sum + PreciseSum, :81-88), hence its posteriors post = lik * (1/W_k) (:91,:97)
//    and the differences post - prior of the velocity term (:119) bit for bit. This matters: for a question that
//    is uninformative under the current posterior, sum (post - prior)^2 is pure rounding noise and the reference's
//    priority depends on it. All threads of a warp read the same r/lr/id2 vector: shared-memory loads are 128-bit
//    warp broadcasts. A thread carries KL*K independent dependency chains, which is what keeps the fp64 pipe busy.
//  * Pass 2 needs log2(posterior) per element (:106): instead of the reference's Log2Hot (one IEEE divide + series)
//    it uses log2(post) = lr[k][j] + log2(prior[j]) - log2(W_k) (two adds; log2 prior is kept per quiz), and falls
//    back to the bit-faithful Log2Hot for the elements where that split would lose accuracy or where Log2Hot's edge
//    semantics matter: post >= 0.5 (cancellation; also Log2Hot(1) = -6.56e-20 != 0) and post < 2^-1022 (Log2Hot(0) =
//    -1023, subnormals). The lack term's divide is a MUFU seed + 3 DFMA reciprocal. The entropy / lack / velocity
//    sums are plain sums (the reference uses Kahan sums): their terms are same-signed, so this costs ~1e-14
//    relative. Net: W_k bit-exact, H_k / V_k / lack / priority within the tolerance stated in DESIGN.md and enforced
//    by tests/test_gpu_parity.py; the fully bit-level path is k_eval_exact in pqa_kernels.cu.
//  * When a slab does not fit in shared memory (large T) the targets are processed in chunks; pass 1 runs over all
//    chunks (the Kahan state lives in registers across chunks), then pass 2 re-stages them (sA/mD are then read twice
//    per CTA, from L2 when resident).
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

#include <atomic>
#include <math.h>
#include <stdio.h>
#include <stdexcept>

namespace pqa {
void count_launch();

struct StagedParams {
  DeviceKB kb;
  QuizPool qp;
  int64_t n;
  const int64_t *slots;
  double *priority;
  EvalDetail det;
  int64_t Jc;             // targets per shared-memory chunk (multiple of 4)
  int64_t nChunks;
  int64_t quizzesPerCta;  // == quizzes per pass of one CTA when nChunks > 1
  PeerBufs mirror;        // EvalConfig::mirror
};

__device__ __forceinline__ void store_priority(const StagedParams &P, int64_t o, double v) {
  P.priority[o] = v;
  for (int r = 0; r < P.mirror.n; r++) P.mirror.p[r][o] = v;   // other shards' inboxes (peer memory)
}

template <int KL> struct VecD { double v[KL]; };

// KL consecutive doubles starting at a KL*8-byte aligned address
template <int KL> __device__ __forceinline__ VecD<KL> lds_vec(const double *p) {
  VecD<KL> o;
  if (KL == 4) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    o.v[0] = a.x; o.v[1 % KL] = a.y; o.v[2 % KL] = b.x; o.v[3 % KL] = b.y;
  } else if (KL == 2) {
    const double2 a = *reinterpret_cast<const double2 *>(p);
    o.v[0] = a.x; o.v[1 % KL] = a.y;
  } else {
    o.v[0] = *p;
  }
  return o;
}
template <int KL> __device__ __forceinline__ VecD<KL> ldg_vec(const double *p) {
  VecD<KL> o;
  if (KL == 4) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p + 2));
    o.v[0] = a.x; o.v[1 % KL] = a.y; o.v[2 % KL] = b.x; o.v[3 % KL] = b.y;
  } else if (KL == 2) {
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
    o.v[0] = a.x; o.v[1 % KL] = a.y;
  } else {
    o.v[0] = __ldg(p);
  }
  return o;
}

template <int K, int THREADS>
__device__ __forceinline__ void stage_chunk(const StagedParams &P, int64_t iLocal, int64_t c, bool withLog, double *sR,
                                            double *sLR, double *sID2, uint64_t *bar, uint32_t &parity) {
  const int64_t Jc = P.Jc, Tp = P.kb.Tp, T = P.kb.T;
  const int64_t j0 = c * Jc;
  const int64_t cnt = (Tp - j0 < Jc) ? (Tp - j0) : Jc;  // multiple of 4 doubles = 32 bytes
  if (threadIdx.x == 0) {
    fence_proxy_async_smem();  // earlier generic-proxy accesses of the buffers (all threads, ordered by the CTA
                               // barrier before this call) precede the async-proxy writes below
    const uint32_t bytes = (uint32_t)(cnt * sizeof(double));
    mbar_arrive_expect_tx(bar, bytes * (K + 1));
#pragma unroll
    for (int k = 0; k < K; k++) bulk_g2s(sR + k * Jc, P.kb.sA + (iLocal * K + k) * Tp + j0, bytes, bar);
    bulk_g2s(sID2, P.kb.mD + iLocal * Tp + j0, bytes, bar);
  }
  mbar_wait(bar, parity);
  parity ^= 1u;
  for (int64_t j = threadIdx.x; j < cnt; j += THREADS) {
    const int64_t gj = j0 + j;
    const bool gap = gj >= T || bit32(P.kb.tgaps, gj);
    const double invD = __ddiv_rn(1.0, sID2[j]);                         // :72-76
    sID2[j] = gap ? 0.0 : __dmul_rn(invD, invD);
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double r = gap ? 0.0 : __dmul_rn(sR[k * Jc + j], invD);     // :81
      sR[k * Jc + j] = r;
      if (withLog) sLR[k * Jc + j] = log2(r);
    }
  }
  __syncthreads();
}

// Pass 1 over one chunk: thread owns Kahan lanes l0 .. l0+KL-1 of its quiz. Padding / gap lanes hold prior = +0 and
// r = 0, which add +0 exactly as the reference's masked lanes do.
template <int K, int KL>
__device__ __forceinline__ void pass1_chunk(const double *__restrict__ sR, int64_t Jc, int nVects, int64_t j0,
                                            const double *__restrict__ pr, int l0, Kahan (&kw)[KL][K]) {
  const double *prc = pr + j0 + l0;
  // the priors come from L2 (every quiz has its own row: no reuse in L1) and are fetched PF vectors ahead. Measured at
  // 1000x5x1000, B=256: PF = 1 1.50e8 q-evals/s, PF = 3 1.45e8 (and -18 % on the chunked 4-warp shape): deeper prefetch
  // costs more than the long-scoreboard stalls it removes.
  constexpr int PF = 1;
  VecD<KL> pq[PF];
#pragma unroll
  for (int d = 0; d < PF; d++) pq[d] = ldg_vec<KL>(prc + 4 * (d < nVects ? d : 0));
  for (int v = 0; v < nVects; v++) {
    const VecD<KL> p = pq[0];
#pragma unroll
    for (int d = 0; d + 1 < PF; d++) pq[d] = pq[d + 1];
    if (v + PF < nVects) pq[PF - 1] = ldg_vec<KL>(prc + 4 * (v + PF));
    const int j = 4 * v + l0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const VecD<KL> r = lds_vec<KL>(sR + k * Jc + j);
#pragma unroll
      for (int e = 0; e < KL; e++) kw[e][k].add(__dmul_rn(r.v[e], p.v[e]));   // :81-86
    }
  }
}

// The reference's Log2Hot and an IEEE reciprocal for an element outside the fast range (post >= 0.5, or zero /
// subnormal post). Out of line: it is the rare path and must not bloat the vector step.
__device__ __noinline__ double slow_log2(double post, const double *__restrict__ tbl, double *rl2) {
  const double l2 = log2hot(post, tbl);                                 // :106, reference semantics
  *rl2 = __ddiv_rn(1.0, l2);
  return l2;
}

// high word of post in [0x00100000, 0x3FE00000) <=> 2^-1022 <= post < 0.5 (positive, normal): the split log is accurate
__device__ __forceinline__ unsigned fast_range_key(double post) { return (unsigned)__double2hiint(post) - 0x00100000u; }
constexpr unsigned kFastRangeLimit = 0x3FE00000u - 0x00100000u;

template <int K, int KL>
__device__ __forceinline__ void pass2_chunk(const double *__restrict__ sR, const double *__restrict__ sLR,
                                            const double *__restrict__ sID2, int64_t Jc, int nVects, int64_t j0,
                                            const double *__restrict__ pr, const double *__restrict__ lpr,
                                            const double *__restrict__ tbl, int l0, const double (&iW)[K],
                                            const double (&lW)[K], double (&H)[K], double (&V)[K], double (&L)[KL]) {
  const double *prc = pr + j0 + l0, *lprc = lpr + j0 + l0;
  VecD<KL> pn = ldg_vec<KL>(prc), lpn = ldg_vec<KL>(lprc);
  for (int v = 0; v < nVects; v++) {
    const VecD<KL> p = pn, lp = lpn;
    if (v + 1 < nVects) { pn = ldg_vec<KL>(prc + 4 * (v + 1)); lpn = ldg_vec<KL>(lprc + 4 * (v + 1)); }
    const int j = 4 * v + l0;
    const VecD<KL> id2 = lds_vec<KL>(sID2 + j);
    // posteriors of the K*KL elements of this vector step, and whether all of them are in the fast range
    double post[K][KL];
    unsigned worst = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const VecD<KL> r = lds_vec<KL>(sR + k * Jc + j);
#pragma unroll
      for (int e = 0; e < KL; e++) {
        post[k][e] = __dmul_rn(__dmul_rn(r.v[e], p.v[e]), iW[k]);       // :81-82, :97
        worst = max(worst, fast_range_key(post[k][e]));
      }
    }
    if (worst < kFastRangeLimit) {
      // common case: no masks, no selects
#pragma unroll
      for (int k = 0; k < K; k++) {
        const VecD<KL> lr = lds_vec<KL>(sLR + k * Jc + j);
#pragma unroll
        for (int e = 0; e < KL; e++) {
          const double l2 = __dsub_rn(__dadd_rn(lr.v[e], lp.v[e]), lW[k]);
          H[k] = __fma_rn(post[k][e], l2, H[k]);                        // :113-114
          L[e] = __fma_rn(id2.v[e], fast_rcp46(l2), L[e]);              // :116-117 (id2 = 0 on gap / padding lanes)
          const double d = __dsub_rn(post[k][e], p.v[e]);               // :119
          V[k] = __fma_rn(d, d, V[k]);                                  // :126-127
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; k++) {
        const VecD<KL> lr = lds_vec<KL>(sLR + k * Jc + j);
#pragma unroll
        for (int e = 0; e < KL; e++) {
          double l2, rl2;
          if (fast_range_key(post[k][e]) < kFastRangeLimit) {
            l2 = __dsub_rn(__dadd_rn(lr.v[e], lp.v[e]), lW[k]);
            rl2 = fast_rcp(l2);
          } else {
            l2 = slow_log2(post[k][e], tbl, &rl2);
          }
          H[k] = __fma_rn(post[k][e], l2, H[k]);
          L[e] = __fma_rn(id2.v[e], rl2, L[e]);
          const double d = __dsub_rn(post[k][e], p.v[e]);
          V[k] = __fma_rn(d, d, V[k]);
        }
      }
    }
  }
}

// PreciseSum of the four Kahan lanes of a quiz (SRAccumVectDbl256.h:83-91); the lanes live in 4/KL threads.
template <int K, int KL>
__device__ __forceinline__ void finish_pass1(const Kahan (&kw)[KL][K], double (&W)[K], double (&iW)[K], double (&lW)[K]) {
  constexpr int LPQ = 4 / KL;   // threads per quiz
  const unsigned mask = (LPQ == 4) ? (0xFu << (threadIdx.x & 28u)) : (LPQ == 2) ? (0x3u << (threadIdx.x & 30u)) : 0u;
#pragma unroll
  for (int k = 0; k < K; k++) {
    double s[4], c[4];
#pragma unroll
    for (int ln = 0; ln < 4; ln++) {
      if (LPQ == 1) { s[ln] = kw[ln % KL][k].s; c[ln] = kw[ln % KL][k].c; }
      else { s[ln] = __shfl_sync(mask, kw[ln % KL][k].s, ln / KL, LPQ); c[ln] = __shfl_sync(mask, kw[ln % KL][k].c, ln / KL, LPQ); }
    }
    W[k] = precise_sum4(s[0], s[1], s[2], s[3], c[0], c[1], c[2], c[3]);   // :88
    iW[k] = __ddiv_rn(1.0, W[k]);                                          // :91
    lW[k] = log2(W[k]);
  }
}

// Per (quiz, question) epilogue: CEEvalQsSubtaskConsider.cpp:134-207. H holds sum post*log2 post (negative), L the
// lack sum (negative), V the squared distances.
template <int K>
__device__ __forceinline__ void write_priority(const StagedParams &P, int64_t i, int64_t b, const double (&W)[K],
                                               const double (&H)[K], const double (&V)[K], double L) {
  const int64_t o = b * P.kb.Q + i;
  double totW = 0.0, sumH = 0.0, sumV = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    totW += W[k];                                                       // :89,:134
    sumH = __fma_rn(W[k], -H[k], sumH);                                 // :148-172
    sumV = __fma_rn(W[k], sqrt(V[k]), sumV);
    if (P.det.W) P.det.W[o * K + k] = W[k];
    if (P.det.H) P.det.H[o * K + k] = -H[k];
    if (P.det.V) P.det.V[o * K + k] = V[k];
  }
  const double avgH = sumH / totW, avgV = sumV / totW;                  // :176-177
  const double nExp = exp2(avgH);                                       // :181
  const double cLnMaxV = 0.34657359027997265470861606072909;            // SRMath::_cLnSqrt2
  const double lnV = (avgV == 0) ? -746.0 : log(avgV);                  // :27-29
  const double n1 = (double)(P.kb.nValidTargets + 1);
  const double vComp = 1.0 / (cLnMaxV - lnV + cLnMaxV / (n1 * n1));     // :30-33
  const double lack = -L;                                               // :201
  store_priority(P, o, lack * pow(vComp, 9.0) * pow(nExp, -2.0));       // :207
  if (P.det.lack) P.det.lack[o] = lack;
}

template <int KL> __device__ __forceinline__ double quiz_sum(double v) {
  constexpr int LPQ = 4 / KL;
  if (LPQ == 1) return v;
  const unsigned mask = (LPQ == 4) ? (0xFu << (threadIdx.x & 28u)) : (0x3u << (threadIdx.x & 30u));
#pragma unroll
  for (int o = 1; o < LPQ; o <<= 1) v = __dadd_rn(v, __shfl_xor_sync(mask, v, o, LPQ));
  return v;
}

template <int K, int KL>
__device__ __forceinline__ void finish_pass2(const StagedParams &P, int64_t i, int64_t b, int l0, const double (&W)[K],
                                             double (&H)[K], double (&V)[K], const double (&Lp)[KL]) {
  double L = Lp[0];
#pragma unroll
  for (int e = 1; e < KL; e++) L = __dadd_rn(L, Lp[e]);
#pragma unroll
  for (int k = 0; k < K; k++) { H[k] = quiz_sum<KL>(H[k]); V[k] = quiz_sum<KL>(V[k]); }
  L = quiz_sum<KL>(L);
  if (l0 != 0) return;
  write_priority<K>(P, i, b, W, H, V, L);
}

// WARPS = 8: 256 threads, two CTAs per SM (slab budget 100 KB); WARPS = 4: 128 threads, four CTAs per SM (50 KB), used
// when the batch has at most 64 quizzes so that no warp of a CTA pass is idle.
template <int K, int KL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 2) k_eval_staged(const StagedParams P) {
  constexpr int THREADS = WARPS * 32;
  constexpr int LPQ = 4 / KL;              // threads per quiz
  constexpr int QPW = 32 / LPQ;            // quizzes per warp
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bar;
  double *sR = (double *)smRaw;        // [K][Jc]  sA, then r = sA/mD
  double *sLR = sR + K * P.Jc;         // [K][Jc]  log2 r
  double *sID2 = sLR + K * P.Jc;       // [Jc]     mD, then 1/mD^2

  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, Tp = P.kb.Tp;
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qw = lane / LPQ, l0 = (lane % LPQ) * KL;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);

  if (bit32(P.kb.qgaps, i)) {          // CEEvalQsSubtaskConsider.cpp:54-58
    for (int64_t b = tileFirst + threadIdx.x; b < tileLimit; b += THREADS) store_priority(P, b * Q + i, qnan);
    return;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t parity = 0;
  const double *__restrict__ tbl = P.kb.log2tbl;

  if (P.nChunks == 1) {
    stage_chunk<K, THREADS>(P, iLocal, 0, true, sR, sLR, sID2, &bar, parity);
    const int nVects = (int)(Tp >> 2);
    for (int64_t g0 = tileFirst + (int64_t)warp * QPW; g0 < tileLimit; g0 += (int64_t)WARPS * QPW) {
      const int64_t b = g0 + qw;
      bool live = b < tileLimit;
      const int64_t slot = P.slots[live ? b : tileLimit - 1];
      if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
        if (l0 == 0) store_priority(P, b * Q + i, qnan);
        live = false;
      }
      if (!live) continue;             // all threads of this quiz leave together
      const double *pr = P.qp.priors + slot * Tp, *lpr = P.qp.logPriors + slot * Tp;
      double W[K], iW[K], lW[K], H[K], V[K], L[KL];
      {
        Kahan kw[KL][K];
#pragma unroll
        for (int e = 0; e < KL; e++)
#pragma unroll
          for (int k = 0; k < K; k++) kw[e][k].init();
        pass1_chunk<K, KL>(sR, P.Jc, nVects, 0, pr, l0, kw);
        finish_pass1<K, KL>(kw, W, iW, lW);
      }
#pragma unroll
      for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
#pragma unroll
      for (int e = 0; e < KL; e++) L[e] = 0.0;
      pass2_chunk<K, KL>(sR, sLR, sID2, P.Jc, nVects, 0, pr, lpr, tbl, l0, iW, lW, H, V, L);
      finish_pass2<K, KL>(P, i, b, l0, W, H, V, L);
    }
  } else {
    // chunked targets: this thread keeps its quiz for the whole question
    const int64_t b = tileFirst + (int64_t)warp * QPW + qw;
    bool live = b < tileLimit;
    const int64_t slot = P.slots[live ? b : tileLimit - 1];
    if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
      if (l0 == 0) store_priority(P, b * Q + i, qnan);
      live = false;
    }
    const double *pr = P.qp.priors + slot * Tp, *lpr = P.qp.logPriors + slot * Tp;
    double W[K], iW[K], lW[K], H[K], V[K], L[KL];
    Kahan kw[KL][K];
#pragma unroll
    for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; W[k] = 0.0; iW[k] = 0.0; lW[k] = 0.0; }
#pragma unroll
    for (int e = 0; e < KL; e++) {
      L[e] = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) kw[e][k].init();
    }
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K, THREADS>(P, iLocal, c, false, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int nVects = (int)(((Tp - j0 < P.Jc) ? (Tp - j0) : P.Jc) >> 2);
      if (live) pass1_chunk<K, KL>(sR, P.Jc, nVects, j0, pr, l0, kw);
      __syncthreads();  // everyone is done with the buffers before the next stage overwrites them
    }
    if (live) finish_pass1<K, KL>(kw, W, iW, lW);
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K, THREADS>(P, iLocal, c, true, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int nVects = (int)(((Tp - j0 < P.Jc) ? (Tp - j0) : P.Jc) >> 2);
      if (live) pass2_chunk<K, KL>(sR, sLR, sID2, P.Jc, nVects, j0, pr, lpr, tbl, l0, iW, lW, H, V, L);
      __syncthreads();
    }
    if (live) finish_pass2<K, KL>(P, i, b, l0, W, H, V, L);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Small batches (fewer quizzes than one warp of the kernel above would hold): latency matters more than throughput.
// One CTA per question handles 8 quizzes per round. Pass 1 cannot be spread over targets (the reference's Kahan order is
// a serial dependency per (answer, Kahan lane)): warp w runs the 4K chains of quiz w on lanes 4k + l -- W_k stays
// bit-exact. Pass 2 has no such dependency: for every quiz of the round ALL threads of the CTA stride over the targets,
// warp butterflies leave per-warp partial sums in shared memory, and warp w finishes quiz w. Single chunk only.
// With few quizzes per slab, staging log2 r would cost as much as it saves, so pass 2 evaluates the reference's
// Log2Hot and an IEEE divide for every element (entropy and lack TERMS are then the reference's bits; only the
// summation order differs) and the slab is just (K+1) rows: 48 KB at 1000x5x1000, four CTAs per SM.
// Two shapes. <K, 4, false>: one or two quizzes -- (K+1)-row slab, four CTAs per SM, pass 2 with the reference's Log2Hot and
// IEEE divide per element as described above. <K, 8, true>: 3 .. 31 quizzes -- the slab also carries log2 r (staged once per
// CTA, worth it from the third quiz on) and pass 2 uses the throughput kernel's split logarithm and reciprocal (15 instead
// of ~70 fp64 instructions per element); eight warps = eight quizzes per round, one CTA per question takes all quizzes.
template <int K, int SW, bool LOGSPLIT>
__global__ void __launch_bounds__(SW * 32, LOGSPLIT ? 2 : 4) k_eval_small(const StagedParams P) {
  constexpr int kSmallWarps = SW;
  constexpr int THREADS = kSmallWarps * 32;
  constexpr int NV = 2 * K + 1;                       // H_k, V_k, L
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bar;
  __shared__ double sWk[kSmallWarps][3][K];           // per quiz of the round: W_k, 1/W_k, log2 W_k
  __shared__ double sPart[kSmallWarps][kSmallWarps][NV];   // [quiz][warp][value] partial sums of pass 2
  __shared__ int64_t sSlot[kSmallWarps];              // slot of the quiz, or -1 when absent / already asked
  double *sR = (double *)smRaw, *sLR = sR + K * P.Jc, *sID2 = sR + (LOGSPLIT ? 2 * K : K) * P.Jc;
  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, Tp = P.kb.Tp, T = P.kb.T, Jc = P.Jc;
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  if (bit32(P.kb.qgaps, i)) {
    for (int64_t b = tileFirst + threadIdx.x; b < tileLimit; b += THREADS) store_priority(P, b * Q + i, qnan);
    return;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t parity = 0;
  stage_chunk<K, THREADS>(P, iLocal, 0, LOGSPLIT, sR, sLR, sID2, &bar, parity);
  const double *__restrict__ tbl = P.kb.log2tbl;
  const int nVects = (int)(Tp >> 2);
  for (int64_t b0 = tileFirst; b0 < tileLimit; b0 += kSmallWarps) {      // CTA-uniform
    // ---- pass 1: warp w <-> quiz b0 + w
    {
      const int64_t b = b0 + warp;
      int64_t slot = -1;
      if (b < tileLimit) {
        slot = P.slots[b];
        if (bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
          if (lane == 0) store_priority(P, b * Q + i, qnan);
          slot = -1;
        }
      }
      if (lane == 0) sSlot[warp] = slot;
      if (slot >= 0 && lane < 4 * K) {
        const int k = lane >> 2, l = lane & 3;
        const double *rk = sR + k * Jc + l;
        const double *__restrict__ prl = P.qp.priors + slot * Tp + l;
        Kahan kw; kw.init();
#pragma unroll 4
        for (int v = 0; v < nVects; v++) kw.add(__dmul_rn(rk[4 * v], __ldg(prl + 4 * v)));   // :81-86
        const double w = group_precise_sum(kw);                          // :88
        if (l == 0) { sWk[warp][0][k] = w; sWk[warp][1][k] = __ddiv_rn(1.0, w); sWk[warp][2][k] = log2(w); }
      }
    }
    __syncthreads();
    // ---- pass 2: every quiz of the round, all threads over the targets
    for (int q = 0; q < kSmallWarps; q++) {
      const int64_t slot = sSlot[q];
      if (slot < 0) continue;                                            // CTA-uniform
      const double *__restrict__ pr = P.qp.priors + slot * Tp;
      const double *__restrict__ lpr = P.qp.logPriors + slot * Tp;
      double iW[K], lW[K], H[K], V[K], L = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { iW[k] = sWk[q][1][k]; lW[k] = sWk[q][2][k]; H[k] = 0.0; V[k] = 0.0; }
      for (int j = threadIdx.x; j < (int)T; j += THREADS) {
        const double p = __ldg(pr + j), id2 = sID2[j];
        const double lp = LOGSPLIT ? __ldg(lpr + j) : 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double post = __dmul_rn(__dmul_rn(sR[k * Jc + j], p), iW[k]);   // :81-82, :97
          if (LOGSPLIT) {
            double l2, rl2;
            if (fast_range_key(post) < kFastRangeLimit) {               // the throughput kernel's split logarithm
              l2 = __dsub_rn(__dadd_rn(sLR[k * Jc + j], lp), lW[k]);
              rl2 = fast_rcp46(l2);
            } else {
              l2 = slow_log2(post, tbl, &rl2);
            }
            H[k] = __fma_rn(post, l2, H[k]);
            L = __fma_rn(id2, rl2, L);
          } else {
            const double l2 = log2hot(post, tbl);                       // :106
            H[k] = __fma_rn(post, l2, H[k]);                            // :113-114
            L = __dadd_rn(L, __ddiv_rn(id2, l2));                       // :116-117
          }
          const double d = __dsub_rn(post, p);                          // :119
          V[k] = __fma_rn(d, d, V[k]);                                  // :126-127
        }
      }
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double h = warp_sum(H[k]), vv = warp_sum(V[k]);
        if (lane == 0) { sPart[q][warp][k] = h; sPart[q][warp][K + k] = vv; }
      }
      L = warp_sum(L);
      if (lane == 0) sPart[q][warp][2 * K] = L;
    }
    __syncthreads();
    // ---- finish: warp w <-> quiz b0 + w
    if (sSlot[warp] >= 0 && lane == 0) {
      double W[K], H[K], V[K], L = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { W[k] = sWk[warp][0][k]; H[k] = 0.0; V[k] = 0.0; }
      for (int w = 0; w < kSmallWarps; w++) {
#pragma unroll
        for (int k = 0; k < K; k++) { H[k] = __dadd_rn(H[k], sPart[warp][w][k]); V[k] = __dadd_rn(V[k], sPart[warp][w][K + k]); }
        L = __dadd_rn(L, sPart[warp][w][2 * K]);
      }
      write_priority<K>(P, i, b0 + warp, W, H, V, L);
    }
    __syncthreads();   // sWk / sPart / sSlot are rewritten by the next round
  }
}

// ---------------------------------------------------------------------------------------------------------
// Target-sharded evaluation (SURVEY.md 8e "Targets"; pqa_kernels.cuh): this device holds the columns
// [tFirst, tFirst + P.kb.T) of every row. Same CTA shape and chunk loop as the chunked branch of k_eval_staged, split at
// the point where the reference needs the complete W_k (:88-91):
//   PHASE 1  pass 1 over the local targets -> partial W_k (the local 4 Kahan lanes + PreciseSum) into every peer slot
//   PHASE 2  W_k = sum of the shards' partials in shard order (identical bits on every shard) -> pass 2 over the local
//            targets -> partial H_k (sum post*log2 post), V_k, lack sum into every peer slot
// A quiz' priors are full-length; the thread reads them at offset tFirst. Shard boundaries are multiples of 4 targets
// (the Kahan lane of a target stays j % 4) and only the last shard may have padding lanes, where the priors are +0.
struct TShardParams {
  StagedParams S;
  int64_t tFirst;
  PeerBufs outW, inW, outHVL;
  // Exact-order pipeline (phase 1 only): the 4-lane Kahan state (s, c) of every (quiz, question, answer) is handed from
  // shard to shard in target order, so the LAST shard finishes the reference's own sum: W_k bit-exact across shards.
  // Layout [((b*Q + i)*K + k)*4 + lane]*2 + {s, c}. inState = the previous shard's hand-over (nullptr on the first
  // shard), outState = the next shard's inbox (nullptr on the last shard, which writes W_k to outW instead).
  const double *inState;
  double *outState;
  PipeCtl pipe;
};

__device__ __forceinline__ uint64_t tshard_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// thread 0 of the CTA spins until flags[tile] >= epoch (acquire at system scope), everybody else waits at the barrier
__device__ __forceinline__ void pipe_wait(const PipeCtl &pc, int64_t tile) {
  if (pc.waitFlags != nullptr) {
    if (threadIdx.x == 0) {
      const uint64_t t0 = tshard_timer_ns();
      for (;;) {
        uint64_t seen;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(pc.waitFlags + tile) : "memory");
        if (seen >= pc.epoch) break;
        if (tshard_timer_ns() - t0 > pc.timeoutNs) { *pc.errFlag = pc.epoch; break; }
        __nanosleep(100);
      }
    }
    __syncthreads();
  }
}

template <int K, int KL, int WARPS, int PHASE>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 2) k_eval_tshard(const TShardParams TP) {
  constexpr int THREADS = WARPS * 32;
  constexpr int LPQ = 4 / KL;
  constexpr int QPW = 32 / LPQ;
  constexpr int NV = 2 * K + 1;
  const StagedParams &P = TP.S;
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bar;
  double *sR = (double *)smRaw;
  double *sLR = sR + K * P.Jc;
  double *sID2 = sLR + K * P.Jc;
  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, TpL = P.kb.Tp;
  if (bit32(P.kb.qgaps, i)) return;      // the epilogue writes the NaN
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qw = lane / LPQ, l0 = (lane % LPQ) * KL;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t parity = 0;
  const int64_t b = tileFirst + (int64_t)warp * QPW + qw;
  bool live = b < tileLimit;
  const int64_t slot = P.slots[live ? b : tileLimit - 1];
  if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) live = false;
  const double *pr = P.qp.priors + slot * P.qp.Tp + TP.tFirst, *lpr = P.qp.logPriors + slot * P.qp.Tp + TP.tFirst;
  const int64_t o = b * Q + i;
  const int64_t pipeTile = iLocal / TP.pipe.tileQ;
  pipe_wait(TP.pipe, pipeTile);          // phase 1: the previous shard's hand-over; phase 2: the complete W_k of this tile
  if (PHASE == 1) {
    Kahan kw[KL][K];
#pragma unroll
    for (int e = 0; e < KL; e++)
#pragma unroll
      for (int k = 0; k < K; k++) kw[e][k].init();
    if (live && TP.inState != nullptr) {                 // continue the previous shard's Kahan lanes
#pragma unroll
      for (int e = 0; e < KL; e++)
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double2 sc = *reinterpret_cast<const double2 *>(TP.inState + (((o * K + k) * 4 + l0 + e) * 2));
          kw[e][k].s = sc.x; kw[e][k].c = sc.y;
        }
    }
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K, THREADS>(P, iLocal, c, false, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int nVects = (int)(((TpL - j0 < P.Jc) ? (TpL - j0) : P.Jc) >> 2);
      if (live) pass1_chunk<K, KL>(sR, P.Jc, nVects, j0, pr, l0, kw);
      __syncthreads();
    }
    if (live && TP.outState != nullptr) {                // hand the lanes over to the next shard
#pragma unroll
      for (int e = 0; e < KL; e++)
#pragma unroll
        for (int k = 0; k < K; k++)
          *reinterpret_cast<double2 *>(TP.outState + (((o * K + k) * 4 + l0 + e) * 2)) = make_double2(kw[e][k].s, kw[e][k].c);
    } else if (live) {
      double W[K], iW[K], lW[K];
      finish_pass1<K, KL>(kw, W, iW, lW);
      if (l0 == 0) {
        for (int r = 0; r < TP.outW.n; r++)
#pragma unroll
          for (int k = 0; k < K; k++) TP.outW.p[r][o * K + k] = W[k];
      }
    }
    if (TP.pipe.nSignal > 0) {             // last CTA of the tile publishes it
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int64_t tileLen = (P.kb.qCount - pipeTile * TP.pipe.tileQ < TP.pipe.tileQ) ? P.kb.qCount - pipeTile * TP.pipe.tileQ : TP.pipe.tileQ;
        const unsigned want = (unsigned)(tileLen * gridDim.y);
        if (atomicAdd(TP.pipe.tileCounters + pipeTile, 1u) + 1u == want) {
          TP.pipe.tileCounters[pipeTile] = 0u;
          __threadfence_system();
          for (int r = 0; r < TP.pipe.nSignal; r++)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(TP.pipe.signalFlags[r] + pipeTile), "l"(TP.pipe.epoch) : "memory");
        }
      }
    }
  } else {
    double W[K], iW[K], lW[K], H[K], V[K], L[KL];
#pragma unroll
    for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; W[k] = 0.0; iW[k] = 0.0; lW[k] = 0.0; }
#pragma unroll
    for (int e = 0; e < KL; e++) L[e] = 0.0;
    if (live) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        double w = TP.inW.p[0][o * K + k];
        for (int r = 1; r < TP.inW.n; r++) w = __dadd_rn(w, TP.inW.p[r][o * K + k]);
        W[k] = w;
        iW[k] = __ddiv_rn(1.0, w);                                         // :91
        lW[k] = log2(w);
      }
    }
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K, THREADS>(P, iLocal, c, true, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int nVects = (int)(((TpL - j0 < P.Jc) ? (TpL - j0) : P.Jc) >> 2);
      if (live) pass2_chunk<K, KL>(sR, sLR, sID2, P.Jc, nVects, j0, pr, lpr, P.kb.log2tbl, l0, iW, lW, H, V, L);
      __syncthreads();
    }
    if (live) {
      double Ls = L[0];
#pragma unroll
      for (int e = 1; e < KL; e++) Ls = __dadd_rn(Ls, L[e]);
#pragma unroll
      for (int k = 0; k < K; k++) { H[k] = quiz_sum<KL>(H[k]); V[k] = quiz_sum<KL>(V[k]); }
      Ls = quiz_sum<KL>(Ls);
      if (l0 == 0) {
        for (int r = 0; r < TP.outHVL.n; r++) {
          double *dst = TP.outHVL.p[r] + o * NV;
#pragma unroll
          for (int k = 0; k < K; k++) { dst[k] = H[k]; dst[K + k] = V[k]; }
          dst[2 * K] = Ls;
        }
      }
    }
  }
}

// Epilogue of the target-sharded evaluation: one thread per (quiz, question) sums the shards' slots in shard order and
// computes the priority exactly like the single-device kernel does from its own sums.
struct TShardEpilogueParams {
  StagedParams S;
  PeerBufs inW, inHVL;
};
template <int K>
__global__ void __launch_bounds__(128) k_tshard_priority(const TShardEpilogueParams EP) {
  constexpr int NV = 2 * K + 1;
  const StagedParams &P = EP.S;
  const int64_t Q = P.kb.Q, total = P.n * Q;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = o / Q, i = o - b * Q;
    const int64_t slot = P.slots[b];
    if (bit32(P.kb.qgaps, i) || bit64(P.qp.asked + slot * P.qp.askedWords, i)) { store_priority(P, o, qnan); continue; }
    double W[K], H[K], V[K], L;
#pragma unroll
    for (int k = 0; k < K; k++) { W[k] = EP.inW.p[0][o * K + k]; H[k] = EP.inHVL.p[0][o * NV + k]; V[k] = EP.inHVL.p[0][o * NV + K + k]; }
    L = EP.inHVL.p[0][o * NV + 2 * K];
    for (int r = 1; r < EP.inW.n; r++) {       // n = 1 when the exact-order pipeline delivered the complete W_k
#pragma unroll
      for (int k = 0; k < K; k++) W[k] = __dadd_rn(W[k], EP.inW.p[r][o * K + k]);
    }
    for (int r = 1; r < EP.inHVL.n; r++) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        H[k] = __dadd_rn(H[k], EP.inHVL.p[r][o * NV + k]);
        V[k] = __dadd_rn(V[k], EP.inHVL.p[r][o * NV + K + k]);
      }
      L = __dadd_rn(L, EP.inHVL.p[r][o * NV + 2 * K]);
    }
    write_priority<K>(P, i, b, W, H, V, L);
  }
}

constexpr int64_t kSlabBudgetWide = 100 * 1024;    // two CTAs of 8 warps per SM
constexpr int64_t kSlabBudgetNarrow = 50 * 1024;   // four CTAs of 4 warps per SM (batches of <= 64 quizzes)

template <int K, int PHASE>
static void launch_tshard_k(TShardParams TP, size_t smem, cudaStream_t st) {
  // two threads per quiz; 128 quizzes per CTA pass (8 warps), or 64 (4 warps, twice the CTAs per SM) for small batches
  const bool wide = TP.S.n > 64;
  // function attributes are per device: several engines of one process may sit on different GPUs (ShardGroup)
  static std::atomic<unsigned long long> attrDevices{0};
  int attrDev = 0;
  cudaGetDevice(&attrDev);
  if (!((attrDevices.load(std::memory_order_relaxed) >> (attrDev & 63)) & 1ull)) {
    cudaFuncSetAttribute(k_eval_tshard<K, 2, 8, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_eval_tshard<K, 2, 4, PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attrDevices.fetch_or(1ull << (attrDev & 63), std::memory_order_relaxed);
  }
  const int64_t perPass = wide ? 128 : 64;
  TP.S.quizzesPerCta = perPass;
  dim3 grid((unsigned)TP.S.kb.qCount, (unsigned)((TP.S.n + perPass - 1) / perPass));
  if (wide) k_eval_tshard<K, 2, 8, PHASE><<<grid, 256, smem, st>>>(TP);
  else k_eval_tshard<K, 2, 4, PHASE><<<grid, 128, smem, st>>>(TP);
  count_launch();
}

// Chunk geometry: the whole local row when it fits the budget, else chunks of a multiple of 32 targets.
static size_t slab_geometry(StagedParams &P, const EvalConfig &cfg, int64_t budget) {
  const int64_t bytesPerTarget = (2 * P.kb.K + 1) * (int64_t)sizeof(double);
  int64_t Jc = cfg.chunkTargets > 0 ? ((cfg.chunkTargets + 3) & ~3ll) : P.kb.Tp;
  if (Jc > P.kb.Tp) Jc = P.kb.Tp;
  if (Jc * bytesPerTarget > budget) Jc = (budget / bytesPerTarget) & ~31ll;
  P.Jc = Jc;
  P.nChunks = (P.kb.Tp + Jc - 1) / Jc;
  P.quizzesPerCta = 0;
  return (size_t)(Jc * bytesPerTarget);
}
int64_t tshard_quiz_tiles(int64_t n) {
  const int64_t perPass = n > 64 ? 128 : 64;
  return (n + perPass - 1) / perPass;
}
static size_t tshard_geometry(StagedParams &P, const EvalConfig &cfg) {
  return slab_geometry(P, cfg, P.n > 64 ? kSlabBudgetWide : kSlabBudgetNarrow);
}

#define PQA_K_SWITCH(K_, CALL)                                  \
  switch (K_) {                                                 \
    case 2: { constexpr int KK = 2; CALL; } break;              \
    case 3: { constexpr int KK = 3; CALL; } break;              \
    case 4: { constexpr int KK = 4; CALL; } break;              \
    case 5: { constexpr int KK = 5; CALL; } break;              \
    case 6: { constexpr int KK = 6; CALL; } break;              \
    case 7: { constexpr int KK = 7; CALL; } break;              \
    case 8: { constexpr int KK = 8; CALL; } break;              \
    default: throw std::runtime_error("probqa_b200: target-sharded evaluation supports 2..8 answer options"); \
  }

void launch_eval_tshard_w(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n, const int64_t *dSlots,
                          const PeerBufs &outW, const EvalConfig &cfg, cudaStream_t st, const double *inState,
                          double *outState, const PipeCtl *pipe) {
  TShardParams TP;
  TP.inState = inState; TP.outState = outState;
  if (pipe) TP.pipe = *pipe;
  TP.S.kb = kbLocal; TP.S.qp = qp; TP.S.n = n; TP.S.slots = dSlots; TP.S.priority = nullptr;
  TP.S.det = EvalDetail{nullptr, nullptr, nullptr, nullptr};
  TP.tFirst = tFirst; TP.outW = outW; TP.inW.n = 0; TP.outHVL.n = 0;
  const size_t smem = tshard_geometry(TP.S, cfg);
  PQA_K_SWITCH(kbLocal.K, (launch_tshard_k<KK, 1>(TP, smem, st)))
}

void launch_eval_tshard_hvl(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n, const int64_t *dSlots,
                            const PeerBufs &inW, const PeerBufs &outHVL, const EvalConfig &cfg, cudaStream_t st,
                            const PipeCtl *pipe) {
  TShardParams TP;
  if (pipe) { TP.pipe.waitFlags = pipe->waitFlags; TP.pipe.epoch = pipe->epoch; TP.pipe.tileQ = pipe->tileQ;
              TP.pipe.timeoutNs = pipe->timeoutNs; TP.pipe.errFlag = pipe->errFlag; }
  TP.S.kb = kbLocal; TP.S.qp = qp; TP.S.n = n; TP.S.slots = dSlots; TP.S.priority = nullptr;
  TP.S.det = EvalDetail{nullptr, nullptr, nullptr, nullptr};
  TP.tFirst = tFirst; TP.outW.n = 0; TP.inW = inW; TP.outHVL = outHVL; TP.inState = nullptr; TP.outState = nullptr;
  const size_t smem = tshard_geometry(TP.S, cfg);
  PQA_K_SWITCH(kbLocal.K, (launch_tshard_k<KK, 2>(TP, smem, st)))
}

void launch_tshard_priority(const DeviceKB &kbLocal, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                            const PeerBufs &inW, const PeerBufs &inHVL, double *dPriority, const EvalDetail &det,
                            cudaStream_t st) {
  TShardEpilogueParams EP;
  EP.S.kb = kbLocal; EP.S.qp = qp; EP.S.n = n; EP.S.slots = dSlots; EP.S.priority = dPriority; EP.S.det = det;
  EP.S.Jc = 0; EP.S.nChunks = 0; EP.S.quizzesPerCta = 0;
  EP.inW = inW; EP.inHVL = inHVL;
  const int64_t total = n * kbLocal.Q;
  int64_t g = (total + 127) / 128;
  if (g > 148 * 16) g = 148 * 16;
  PQA_K_SWITCH(kbLocal.K, (k_tshard_priority<KK><<<(unsigned)g, 128, 0, st>>>(EP)))
  count_launch();
}

template <int K>
static void launch_small(StagedParams P, size_t smem, cudaStream_t st) {
  // function attributes are per device: several engines of one process may sit on different GPUs (ShardGroup)
  static std::atomic<unsigned long long> attrDevices{0};
  int attrDev = 0;
  cudaGetDevice(&attrDev);
  if (!((attrDevices.load(std::memory_order_relaxed) >> (attrDev & 63)) & 1ull)) {
    cudaFuncSetAttribute(k_eval_small<K, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_eval_small<K, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attrDevices.fetch_or(1ull << (attrDev & 63), std::memory_order_relaxed);
  }
  if (P.n <= 2) {
    P.quizzesPerCta = 4;
    dim3 grid((unsigned)P.kb.qCount, (unsigned)((P.n + 3) / 4));
    smem = (size_t)((K + 1) * P.Jc) * sizeof(double);
    k_eval_small<K, 4, false><<<grid, 4 * 32, smem, st>>>(P);
  } else {
    P.quizzesPerCta = 32;                                    // one CTA per question, rounds of eight quizzes
    dim3 grid((unsigned)P.kb.qCount, (unsigned)((P.n + 31) / 32));
    smem = (size_t)((2 * K + 1) * P.Jc) * sizeof(double);
    k_eval_small<K, 8, true><<<grid, 8 * 32, smem, st>>>(P);
  }
  count_launch();
}

template <int K, int KL, int WARPS>
static void launch_cfg(StagedParams P, const EvalConfig &cfg, size_t smem, cudaStream_t st) {
  // function attributes are per device: several engines of one process may sit on different GPUs (ShardGroup)
  static std::atomic<unsigned long long> attrDevices{0};
  int attrDev = 0;
  cudaGetDevice(&attrDev);
  if (!((attrDevices.load(std::memory_order_relaxed) >> (attrDev & 63)) & 1ull)) {
    cudaFuncSetAttribute(k_eval_staged<K, KL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attrDevices.fetch_or(1ull << (attrDev & 63), std::memory_order_relaxed);
  }
  const int64_t perPass = (int64_t)WARPS * (32 / (4 / KL));   // quizzes one CTA evaluates concurrently
  if (P.nChunks > 1) {
    P.quizzesPerCta = perPass;
  } else if (cfg.quizzesPerCta > 0) {
    P.quizzesPerCta = ((cfg.quizzesPerCta + perPass - 1) / perPass) * perPass;
  } else {
    // enough CTAs for ~12 waves of 2 CTAs/SM when the batch allows it, else one pass per CTA
    const int64_t passesTotal = (P.n + perPass - 1) / perPass;
    const int64_t targetCtas = (int64_t)cfg.smCount * 2 * 12;
    int64_t passesPerCta = (P.kb.qCount * passesTotal) / targetCtas;
    if (passesPerCta < 1) passesPerCta = 1;
    if (passesPerCta > passesTotal) passesPerCta = passesTotal;
    P.quizzesPerCta = passesPerCta * perPass;
  }
  const int64_t tiles = (P.n + P.quizzesPerCta - 1) / P.quizzesPerCta;
  dim3 grid((unsigned)P.kb.qCount, (unsigned)tiles);
  k_eval_staged<K, KL, WARPS><<<grid, WARPS * 32, smem, st>>>(P);
  count_launch();
}

template <int K>
static void launch_k(const StagedParams &P, const EvalConfig &cfg, size_t smem, cudaStream_t st) {
  if (cfg.kahanLanesPerThread == 0 && P.n < 32 && P.nChunks == 1 && cfg.chunkTargets == 0) { launch_small<K>(P, smem, st); return; }
  // batches >= 32: two threads per quiz (16 quizzes per warp, 16 warps per SM; measured fastest); small batches: four
  // threads per quiz; one thread per quiz (4 lanes, 4 warps per CTA) is kept selectable for experiments
  const int lanesPerThread = cfg.kahanLanesPerThread > 0 ? cfg.kahanLanesPerThread : (P.n >= 32 ? 2 : 1);
  if (lanesPerThread == 4) launch_cfg<K, 4, 4>(P, cfg, smem, st);
  else if (lanesPerThread == 2 && P.n <= 64 && P.nChunks > 1) launch_cfg<K, 2, 4>(P, cfg, smem, st);   // 64 quizzes per CTA pass, 4 CTAs/SM
  else if (lanesPerThread == 2 && P.n <= 64 && cfg.kahanLanesPerThread == 0) launch_cfg<K, 1, 8>(P, cfg, smem, st);   // whole slab (2 CTAs/SM):
                                                               // four threads per quiz fill all 8 warps (measured 0.58 vs 0.67 ms at B = 64)
  else if (lanesPerThread == 2) launch_cfg<K, 2, 8>(P, cfg, smem, st);
  else launch_cfg<K, 1, 8>(P, cfg, smem, st);
}

void launch_eval_staged(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, double *dPriority,
                        const EvalDetail &det, const EvalConfig &cfg, cudaStream_t st) {
  if (kb.K > 8) {  // more answer options than the register tile covers: use the exact kernel
    EvalConfig c2 = cfg;
    c2.which = 1;
    launch_eval_questions(kb, qp, n, dSlots, dPriority, det, c2, st);
    return;
  }
  StagedParams P;
  P.kb = kb; P.qp = qp; P.n = n; P.slots = dSlots; P.priority = dPriority; P.det = det; P.mirror = cfg.mirror;
  // a slab that fits 100 KB is staged whole; a chunked one uses 50 KB chunks for batches of <= 64 quizzes, which run
  // four 4-warp CTAs per SM (launch_k)
  size_t smem = slab_geometry(P, cfg, kSlabBudgetWide);
  if (P.nChunks > 1 && n <= 64 && cfg.chunkTargets == 0 && (cfg.kahanLanesPerThread == 0 || cfg.kahanLanesPerThread == 2))
    smem = slab_geometry(P, cfg, kSlabBudgetNarrow);
  switch (kb.K) {
    case 2: launch_k<2>(P, cfg, smem, st); break;
    case 3: launch_k<3>(P, cfg, smem, st); break;
    case 4: launch_k<4>(P, cfg, smem, st); break;
    case 5: launch_k<5>(P, cfg, smem, st); break;
    case 6: launch_k<6>(P, cfg, smem, st); break;
    case 7: launch_k<7>(P, cfg, smem, st); break;
    default: launch_k<8>(P, cfg, smem, st); break;
  }
}

template <int K> static void preload_staged_k() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_eval_staged<K, 1, 8>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 2, 8>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 4, 4>);
  cudaFuncGetAttributes(&a, k_eval_small<K, 4, false>);
  cudaFuncGetAttributes(&a, k_eval_small<K, 8, true>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 2, 4>);
  cudaFuncGetAttributes(&a, k_eval_tshard<K, 2, 4, 1>);
  cudaFuncGetAttributes(&a, k_eval_tshard<K, 2, 8, 1>);
  cudaFuncGetAttributes(&a, k_eval_tshard<K, 2, 4, 2>);
  cudaFuncGetAttributes(&a, k_eval_tshard<K, 2, 8, 2>);
  cudaFuncGetAttributes(&a, k_tshard_priority<K>);
}
void preload_staged_kernels(int K) {
  if (K < 2 || K > 8) return;
  PQA_K_SWITCH(K, (preload_staged_k<KK>()))
}

} // namespace pqa
