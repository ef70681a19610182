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
//  * A warp evaluates 8 quizzes at a time: lane = (quiz q8 = lane/4, Kahan lane l = lane%4). Thread (q8, l) owns the
//    targets j with j%4 == l of its quiz and walks them in vector order -- exactly the element-to-lane assignment
//    and order of the reference's AVX2 code, so pass 1 reproduces the reference's normaliser W_k BIT FOR BIT
//    (4-lane Kahan sum + PreciseSum, :81-88), hence its posteriors post = lik * (1/W_k) (:91,:97) and the differences
//    post - prior of the velocity term (:119) bit for bit. This matters: for a question that is uninformative under
//    the current posterior, sum (post - prior)^2 is pure rounding noise and the reference's priority depends on it.
//    The four lanes of a quiz read the same r/lr/id2 vector (32 contiguous bytes), and the 8 quizzes of a warp read
//    the same vector: shared-memory loads are warp broadcasts; the end-of-pass reductions are width-4 shuffles.
//  * Pass 2 needs log2(posterior) per element (:106): instead of the reference's Log2Hot (one IEEE divide + series)
//    it uses log2(post) = lr[k][j] + log2(prior[j]) - log2(W_k) (two adds; log2 prior is kept per quiz), and falls
//    back to the bit-faithful Log2Hot for the elements where that split would lose accuracy or where Log2Hot's edge
//    semantics matter: post >= 0.5 (cancellation; also Log2Hot(1) = -6.56e-20 != 0) and post < 2^-1022 (Log2Hot(0) =
//    -1023, subnormals). The lack term's divide is a MUFU seed + 3 DFMA reciprocal. The entropy / lack / velocity
//    sums are plain per-lane sums (the reference uses Kahan sums): their terms are same-signed, so this costs ~1e-14
//    relative. Net: W_k bit-exact, H_k / V_k / lack / priority within the tolerance stated in DESIGN.md and enforced
//    by tests/test_gpu_parity.py; the fully bit-level path is k_eval_exact in pqa_kernels.cu.
//  * When a slab does not fit in shared memory (large T) the targets are processed in chunks; pass 1 runs over all
//    chunks (the Kahan state lives in registers across chunks), then pass 2 re-stages them (sA/mD are then read twice
//    per CTA, from L2 when resident).
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

#include <math.h>
#include <stdio.h>

namespace pqa {
void count_launch();

constexpr int kEvalWarps = 8;
constexpr int kEvalThreads = kEvalWarps * 32;
constexpr int kQuizzesPerWarp = 8;   // lane = (quiz, Kahan lane): 8 quizzes x 4 lanes

struct StagedParams {
  DeviceKB kb;
  QuizPool qp;
  int64_t n;
  const int64_t *slots;
  double *priority;
  EvalDetail det;
  int64_t Jc;             // targets per shared-memory chunk (multiple of 4)
  int64_t nChunks;
  int64_t quizzesPerCta;  // == kEvalWarps*kQuizzesPerWarp when nChunks > 1
};

template <int K>
__device__ __forceinline__ void stage_chunk(const StagedParams &P, int64_t i, int64_t c, bool withLog, double *sR,
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
    for (int k = 0; k < K; k++) bulk_g2s(sR + k * Jc, P.kb.sA + (i * K + k) * Tp + j0, bytes, bar);
    bulk_g2s(sID2, P.kb.mD + i * Tp + j0, bytes, bar);
  }
  mbar_wait(bar, parity);
  parity ^= 1u;
  for (int64_t j = threadIdx.x; j < cnt; j += kEvalThreads) {
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

// lane = (q8, l): q8 = quiz within the warp's group of 8, l = Kahan lane. One pass-1 step per 4-target vector.
template <int K>
__device__ __forceinline__ void pass1_chunk(const double *__restrict__ sR, int64_t Jc, int nVects, int64_t j0,
                                            const double *__restrict__ pr, int l, Kahan (&kw)[K]) {
#pragma unroll 2
  for (int v = 0; v < nVects; v++) {
    const int j = 4 * v + l;
    const double p = __ldg(pr + j0 + j);               // padding lanes hold +0 (and r = 0 there)
#pragma unroll
    for (int k = 0; k < K; k++) kw[k].add(__dmul_rn(sR[k * Jc + j], p));   // :81-86
  }
}

template <int K>
__device__ __forceinline__ void pass2_chunk(const double *__restrict__ sR, const double *__restrict__ sLR,
                                            const double *__restrict__ sID2, int64_t Jc, int nVects, int valid,
                                            int64_t j0, const double *__restrict__ pr, const double *__restrict__ lpr,
                                            const double *__restrict__ tbl, int l, const double (&iW)[K],
                                            const double (&lW)[K], double (&H)[K], double (&V)[K], double &L) {
#pragma unroll 2
  for (int v = 0; v < nVects; v++) {
    const int j = 4 * v + l;
    if (j >= valid) break;                               // padding lanes of the last vector (gap mask, :103-117)
    const double p = __ldg(pr + j0 + j), lp = __ldg(lpr + j0 + j);
    const double id2 = sID2[j];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double lik = __dmul_rn(sR[k * Jc + j], p);                  // :81-82
      const double post = __dmul_rn(lik, iW[k]);                        // :97
      double l2 = __dsub_rn(__dadd_rn(sLR[k * Jc + j], lp), lW[k]);
      double rl2;
      // high word in [0x00100000, 0x3FE00000) <=> 2^-1022 <= post < 0.5 (positive, normal): the split log is accurate
      const unsigned hi = (unsigned)__double2hiint(post);
      if (hi - 0x00100000u >= 0x3FE00000u - 0x00100000u) {
        l2 = log2hot(post, tbl);                                        // :106, reference semantics
        rl2 = __ddiv_rn(1.0, l2);
      } else {
        rl2 = fast_rcp(l2);
      }
      H[k] = __fma_rn(post, l2, H[k]);                                  // :113-114
      L = __fma_rn(id2, rl2, L);                                        // :116-117
      const double d = __dsub_rn(post, p);                              // :119
      V[k] = __fma_rn(d, d, V[k]);                                      // :126-127
    }
  }
}

// sum over the four lanes of a quiz; every lane of the group returns the total
__device__ __forceinline__ double group_sum4(double v) {
  const unsigned mask = 0xFu << (threadIdx.x & 28u);
  v = __dadd_rn(v, __shfl_xor_sync(mask, v, 1, 4));
  v = __dadd_rn(v, __shfl_xor_sync(mask, v, 2, 4));
  return v;
}

template <int K>
__device__ __forceinline__ void finish_pass1(const Kahan (&kw)[K], double (&W)[K], double (&iW)[K], double (&lW)[K]) {
#pragma unroll
  for (int k = 0; k < K; k++) {
    W[k] = group_precise_sum(kw[k]);                                    // :88 (PreciseSum of the 4 Kahan lanes)
    iW[k] = __ddiv_rn(1.0, W[k]);                                       // :91
    lW[k] = log2(W[k]);
  }
}

template <int K>
__device__ __forceinline__ void finish_pass2(const StagedParams &P, int64_t i, int64_t b, int l, const double (&W)[K],
                                             double (&H)[K], double (&V)[K], double L) {
#pragma unroll
  for (int k = 0; k < K; k++) { H[k] = group_sum4(H[k]); V[k] = group_sum4(V[k]); }
  L = group_sum4(L);
  if (l != 0) return;
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
  P.priority[o] = lack * pow(vComp, 9.0) * pow(nExp, -2.0);             // :207
  if (P.det.lack) P.det.lack[o] = lack;
}

template <int K>
__global__ void __launch_bounds__(kEvalThreads, 2) k_eval_staged(const StagedParams P) {
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bar;
  double *sR = (double *)smRaw;        // [K][Jc]  sA, then r = sA/mD
  double *sLR = sR + K * P.Jc;         // [K][Jc]  log2 r
  double *sID2 = sLR + K * P.Jc;       // [Jc]     mD, then 1/mD^2

  const int64_t i = blockIdx.x, Q = P.kb.Q, Tp = P.kb.Tp, T = P.kb.T;
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q8 = lane >> 2, l = lane & 3;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);

  if (bit32(P.kb.qgaps, i)) {          // CEEvalQsSubtaskConsider.cpp:54-58
    for (int64_t b = tileFirst + threadIdx.x; b < tileLimit; b += kEvalThreads) P.priority[b * Q + i] = qnan;
    return;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t parity = 0;
  const double *__restrict__ tbl = P.kb.log2tbl;

  if (P.nChunks == 1) {
    stage_chunk<K>(P, i, 0, true, sR, sLR, sID2, &bar, parity);
    const int nVects = (int)(Tp >> 2), valid = (int)T;
    for (int64_t g0 = tileFirst + (int64_t)warp * kQuizzesPerWarp; g0 < tileLimit; g0 += (int64_t)kEvalWarps * kQuizzesPerWarp) {
      const int64_t b = g0 + q8;
      bool live = b < tileLimit;
      const int64_t slot = P.slots[live ? b : tileLimit - 1];
      if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
        if (l == 0) P.priority[b * Q + i] = qnan;
        live = false;
      }
      if (!live) continue;             // the whole 4-lane group of this quiz leaves together
      const double *pr = P.qp.priors + slot * Tp, *lpr = P.qp.logPriors + slot * Tp;
      double W[K], iW[K], lW[K], H[K], V[K], L = 0.0;
      {
        Kahan kw[K];
#pragma unroll
        for (int k = 0; k < K; k++) kw[k].init();
        pass1_chunk<K>(sR, P.Jc, nVects, 0, pr, l, kw);
        finish_pass1<K>(kw, W, iW, lW);
      }
#pragma unroll
      for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
      pass2_chunk<K>(sR, sLR, sID2, P.Jc, nVects, valid, 0, pr, lpr, tbl, l, iW, lW, H, V, L);
      finish_pass2<K>(P, i, b, l, W, H, V, L);
    }
  } else {
    // chunked targets: this lane keeps its quiz for the whole question
    const int64_t b = tileFirst + (int64_t)warp * kQuizzesPerWarp + q8;
    bool live = b < tileLimit;
    const int64_t slot = P.slots[live ? b : tileLimit - 1];
    if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
      if (l == 0) P.priority[b * Q + i] = qnan;
      live = false;
    }
    const double *pr = P.qp.priors + slot * Tp, *lpr = P.qp.logPriors + slot * Tp;
    double W[K], iW[K], lW[K], H[K], V[K], L = 0.0;
    Kahan kw[K];
#pragma unroll
    for (int k = 0; k < K; k++) { kw[k].init(); H[k] = 0.0; V[k] = 0.0; W[k] = 0.0; iW[k] = 0.0; lW[k] = 0.0; }
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K>(P, i, c, false, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int nVects = (int)(((Tp - j0 < P.Jc) ? (Tp - j0) : P.Jc) >> 2);
      if (live) pass1_chunk<K>(sR, P.Jc, nVects, j0, pr, l, kw);
      __syncthreads();  // everyone is done with the buffers before the next stage overwrites them
    }
    if (live) finish_pass1<K>(kw, W, iW, lW);
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K>(P, i, c, true, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int64_t cnt = (Tp - j0 < P.Jc) ? (Tp - j0) : P.Jc;
      const int nVects = (int)(cnt >> 2);
      const int valid = (int)((T - j0 < cnt) ? (T - j0 > 0 ? T - j0 : 0) : cnt);
      if (live) pass2_chunk<K>(sR, sLR, sID2, P.Jc, nVects, valid, j0, pr, lpr, tbl, l, iW, lW, H, V, L);
      __syncthreads();
    }
    if (live) finish_pass2<K>(P, i, b, l, W, H, V, L);
  }
}

template <int K>
static void launch_k(const StagedParams &P, size_t smem, cudaStream_t st) {
  static bool attrSet = false;
  if (!attrSet) {
    cudaFuncSetAttribute(k_eval_staged<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attrSet = true;
  }
  const int64_t tiles = (P.n + P.quizzesPerCta - 1) / P.quizzesPerCta;
  dim3 grid((unsigned)P.kb.Q, (unsigned)tiles);
  k_eval_staged<K><<<grid, kEvalThreads, smem, st>>>(P);
  count_launch();
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
  P.kb = kb; P.qp = qp; P.n = n; P.slots = dSlots; P.priority = dPriority; P.det = det;
  const int64_t bytesPerTarget = (2 * kb.K + 1) * (int64_t)sizeof(double);
  const int64_t budget = 100 * 1024;  // two CTAs per SM
  int64_t Jc = cfg.chunkTargets > 0 ? ((cfg.chunkTargets + 3) & ~3ll) : kb.Tp;
  if (Jc > kb.Tp) Jc = kb.Tp;
  if (Jc * bytesPerTarget > budget) Jc = (budget / bytesPerTarget) & ~31ll;
  P.Jc = Jc;
  P.nChunks = (kb.Tp + Jc - 1) / Jc;
  const int64_t perPass = (int64_t)kEvalWarps * kQuizzesPerWarp;
  if (P.nChunks > 1) {
    P.quizzesPerCta = perPass;
  } else if (cfg.quizzesPerCta > 0) {
    P.quizzesPerCta = ((cfg.quizzesPerCta + perPass - 1) / perPass) * perPass;
  } else {
    // enough CTAs for ~12 waves of 2 CTAs/SM when the batch allows it, else one pass per CTA
    const int64_t passesTotal = (n + perPass - 1) / perPass;
    const int64_t targetCtas = (int64_t)cfg.smCount * 2 * 12;
    int64_t passesPerCta = (kb.Q * passesTotal) / targetCtas;
    if (passesPerCta < 1) passesPerCta = 1;
    if (passesPerCta > passesTotal) passesPerCta = passesTotal;
    P.quizzesPerCta = passesPerCta * perPass;
  }
  const size_t smem = (size_t)(Jc * bytesPerTarget);
  switch (kb.K) {
    case 2: launch_k<2>(P, smem, st); break;
    case 3: launch_k<3>(P, smem, st); break;
    case 4: launch_k<4>(P, smem, st); break;
    case 5: launch_k<5>(P, smem, st); break;
    case 6: launch_k<6>(P, smem, st); break;
    case 7: launch_k<7>(P, smem, st); break;
    default: launch_k<8>(P, smem, st); break;
  }
}

} // namespace pqa
