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
//  * One warp evaluates BT quizzes at a time; lanes stride over targets. Pass 1 is W_k = sum_j r*prior (:81-88).
//    Pass 2 needs log2(posterior) per element (:106): instead of the reference's Log2Hot (one IEEE divide + series)
//    it uses log2(post) = lr[k][j] + log2(prior[j]) - log2(W_k) (two adds; log2 prior is kept per quiz), and falls
//    back to the bit-faithful Log2Hot for the elements where that split would lose accuracy or where Log2Hot's edge
//    semantics matter: post >= 0.5 (cancellation; also Log2Hot(1) = -6.56e-20 != 0) and post < 2^-1022 (Log2Hot(0) =
//    -1023, subnormals). The lack term's divide is a MUFU seed + 3 DFMA reciprocal. Sums are plain per-lane sums
//    followed by a warp butterfly (the reference uses 4-lane Kahan sums): tolerance-level, not bit-level, parity --
//    see DESIGN.md for the stated tolerance and tests/test_gpu_parity.py for its enforcement; the bit-level path is
//    k_eval_exact in pqa_kernels.cu.
//  * When a slab does not fit in shared memory (large T) the targets are processed in chunks; pass 1 runs over all
//    chunks, then pass 2 re-stages them (sA/mD are then read twice per CTA, from L2 when resident).
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

#include <math.h>
#include <stdio.h>

namespace pqa {
void count_launch();

constexpr int kEvalWarps = 8;
constexpr int kEvalThreads = kEvalWarps * 32;

struct StagedParams {
  DeviceKB kb;
  QuizPool qp;
  int64_t n;
  const int64_t *slots;
  double *priority;
  EvalDetail det;
  int64_t Jc;             // targets per shared-memory chunk (multiple of 4)
  int64_t nChunks;
  int64_t quizzesPerCta;  // == kEvalWarps*BT when nChunks > 1
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
    const double invD = 1.0 / sID2[j];
    sID2[j] = gap ? 0.0 : invD * invD;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double r = gap ? 0.0 : sR[k * Jc + j] * invD;
      sR[k * Jc + j] = r;
      if (withLog) sLR[k * Jc + j] = log2(r);
    }
  }
  __syncthreads();
}

template <int K, int BT>
__device__ __forceinline__ void pass1_chunk(const double *__restrict__ sR, int64_t Jc, int valid, int64_t j0,
                                            const double *const (&pr)[BT], int lane, double (&W)[BT][K]) {
  for (int j = lane; j < valid; j += 32) {
    double p[BT];
#pragma unroll
    for (int bt = 0; bt < BT; bt++) p[bt] = __ldg(pr[bt] + j0 + j);
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double rr = sR[k * Jc + j];
#pragma unroll
      for (int bt = 0; bt < BT; bt++) W[bt][k] = __fma_rn(rr, p[bt], W[bt][k]);
    }
  }
}

template <int K, int BT>
__device__ __forceinline__ void pass2_chunk(const double *__restrict__ sR, const double *__restrict__ sLR,
                                            const double *__restrict__ sID2, int64_t Jc, int valid, int64_t j0,
                                            const double *const (&pr)[BT], const double *const (&lpr)[BT],
                                            const double *__restrict__ tbl, int lane, const double (&iW)[BT][K],
                                            const double (&lW)[BT][K], double (&H)[BT][K], double (&V)[BT][K],
                                            double (&L)[BT]) {
  for (int j = lane; j < valid; j += 32) {
    double p[BT], lp[BT];
#pragma unroll
    for (int bt = 0; bt < BT; bt++) {
      p[bt] = __ldg(pr[bt] + j0 + j);
      lp[bt] = __ldg(lpr[bt] + j0 + j);
    }
    const double id2 = sID2[j];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double rr = sR[k * Jc + j], lrr = sLR[k * Jc + j];
#pragma unroll
      for (int bt = 0; bt < BT; bt++) {
        const double lik = __dmul_rn(rr, p[bt]);                        // :81-82
        const double post = __dmul_rn(lik, iW[bt][k]);                  // :97
        double l2 = __dsub_rn(__dadd_rn(lrr, lp[bt]), lW[bt][k]);
        double rl2;
        // high word in [0x00100000, 0x3FE00000) <=> 2^-1022 <= post < 0.5 (positive, normal): the split log is accurate
        const unsigned hi = (unsigned)__double2hiint(post);
        if (hi - 0x00100000u >= 0x3FE00000u - 0x00100000u) {
          l2 = log2hot(post, tbl);                                      // :106, reference semantics
          rl2 = __ddiv_rn(1.0, l2);
        } else {
          rl2 = fast_rcp(l2);
        }
        H[bt][k] = __fma_rn(post, l2, H[bt][k]);                        // :113-114
        L[bt] = __fma_rn(id2, rl2, L[bt]);                              // :116-117
        const double d = __dsub_rn(post, p[bt]);                        // :119
        V[bt][k] = __fma_rn(d, d, V[bt][k]);                            // :126-127
      }
    }
  }
}

template <int K, int BT>
__device__ __forceinline__ void finish_pass1(double (&W)[BT][K], double (&iW)[BT][K], double (&lW)[BT][K]) {
#pragma unroll
  for (int bt = 0; bt < BT; bt++)
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double w = warp_sum(W[bt][k]);
      W[bt][k] = w;
      iW[bt][k] = 1.0 / w;                                              // :91
      lW[bt][k] = log2(w);
    }
}

template <int K, int BT>
__device__ __forceinline__ void finish_pass2(const StagedParams &P, int64_t i, const int64_t (&bq)[BT],
                                             const bool (&live)[BT], int lane, const double (&W)[BT][K],
                                             double (&H)[BT][K], double (&V)[BT][K], double (&L)[BT]) {
#pragma unroll
  for (int bt = 0; bt < BT; bt++) {
#pragma unroll
    for (int k = 0; k < K; k++) { H[bt][k] = warp_sum(H[bt][k]); V[bt][k] = warp_sum(V[bt][k]); }
    L[bt] = warp_sum(L[bt]);
  }
#pragma unroll
  for (int bt = 0; bt < BT; bt++) {
    if (lane != bt || !live[bt]) continue;
    const int64_t o = bq[bt] * P.kb.Q + i;
    double totW = 0.0, sumH = 0.0, sumV = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) {
      totW += W[bt][k];                                                 // :89,:134
      sumH = __fma_rn(W[bt][k], -H[bt][k], sumH);                       // :148-172
      sumV = __fma_rn(W[bt][k], sqrt(V[bt][k]), sumV);
      if (P.det.W) P.det.W[o * K + k] = W[bt][k];
      if (P.det.H) P.det.H[o * K + k] = -H[bt][k];
      if (P.det.V) P.det.V[o * K + k] = V[bt][k];
    }
    const double avgH = sumH / totW, avgV = sumV / totW;                // :176-177
    const double nExp = exp2(avgH);                                     // :181
    const double cLnMaxV = 0.34657359027997265470861606072909;          // SRMath::_cLnSqrt2
    const double lnV = (avgV == 0) ? -746.0 : log(avgV);                // :27-29
    const double n1 = (double)(P.kb.nValidTargets + 1);
    const double vComp = 1.0 / (cLnMaxV - lnV + cLnMaxV / (n1 * n1));   // :30-33
    const double lack = -L[bt];                                         // :201
    P.priority[o] = lack * pow(vComp, 9.0) * pow(nExp, -2.0);           // :207
    if (P.det.lack) P.det.lack[o] = lack;
  }
}

template <int K, int BT>
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
    const int valid = (int)T;
    for (int64_t b0 = tileFirst + (int64_t)warp * BT; b0 < tileLimit; b0 += (int64_t)kEvalWarps * BT) {
      int64_t bq[BT]; bool live[BT]; const double *pr[BT]; const double *lpr[BT];
      bool any = false;
#pragma unroll
      for (int bt = 0; bt < BT; bt++) {
        bq[bt] = (b0 + bt < tileLimit) ? b0 + bt : tileLimit - 1;
        const int64_t slot = P.slots[bq[bt]];
        live[bt] = (b0 + bt < tileLimit) && !bit64(P.qp.asked + slot * P.qp.askedWords, i);
        any |= live[bt];
        pr[bt] = P.qp.priors + slot * Tp;
        lpr[bt] = P.qp.logPriors + slot * Tp;
        if (lane == bt && (b0 + bt < tileLimit) && !live[bt]) P.priority[bq[bt] * Q + i] = qnan;
      }
      if (!any) continue;
      double W[BT][K], iW[BT][K], lW[BT][K], H[BT][K], V[BT][K], L[BT];
#pragma unroll
      for (int bt = 0; bt < BT; bt++) {
        L[bt] = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) { W[bt][k] = 0.0; H[bt][k] = 0.0; V[bt][k] = 0.0; }
      }
      pass1_chunk<K, BT>(sR, P.Jc, valid, 0, pr, lane, W);
      finish_pass1<K, BT>(W, iW, lW);
      pass2_chunk<K, BT>(sR, sLR, sID2, P.Jc, valid, 0, pr, lpr, tbl, lane, iW, lW, H, V, L);
      finish_pass2<K, BT>(P, i, bq, live, lane, W, H, V, L);
    }
  } else {
    // chunked targets: this warp keeps its BT quizzes for the whole question
    const int64_t b0 = tileFirst + (int64_t)warp * BT;
    int64_t bq[BT]; bool live[BT]; const double *pr[BT]; const double *lpr[BT];
    bool any = false;
#pragma unroll
    for (int bt = 0; bt < BT; bt++) {
      const bool inTile = b0 + bt < tileLimit;
      bq[bt] = inTile ? b0 + bt : tileLimit - 1;
      const int64_t slot = P.slots[bq[bt]];
      live[bt] = inTile && !bit64(P.qp.asked + slot * P.qp.askedWords, i);
      any |= live[bt];
      pr[bt] = P.qp.priors + slot * Tp;
      lpr[bt] = P.qp.logPriors + slot * Tp;
      if (lane == bt && inTile && !live[bt]) P.priority[bq[bt] * Q + i] = qnan;
    }
    double W[BT][K], iW[BT][K], lW[BT][K], H[BT][K], V[BT][K], L[BT];
#pragma unroll
    for (int bt = 0; bt < BT; bt++) {
      L[bt] = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { W[bt][k] = 0.0; H[bt][k] = 0.0; V[bt][k] = 0.0; }
    }
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K>(P, i, c, false, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int valid = (int)((T - j0 < P.Jc) ? (T - j0 > 0 ? T - j0 : 0) : P.Jc);
      if (any) pass1_chunk<K, BT>(sR, P.Jc, valid, j0, pr, lane, W);
      __syncthreads();  // everyone is done with the buffers before the next stage overwrites them
    }
    finish_pass1<K, BT>(W, iW, lW);
    for (int64_t c = 0; c < P.nChunks; c++) {
      stage_chunk<K>(P, i, c, true, sR, sLR, sID2, &bar, parity);
      const int64_t j0 = c * P.Jc;
      const int valid = (int)((T - j0 < P.Jc) ? (T - j0 > 0 ? T - j0 : 0) : P.Jc);
      if (any) pass2_chunk<K, BT>(sR, sLR, sID2, P.Jc, valid, j0, pr, lpr, tbl, lane, iW, lW, H, V, L);
      __syncthreads();
    }
    if (any) finish_pass2<K, BT>(P, i, bq, live, lane, W, H, V, L);
  }
}

template <int K, int BT>
static void launch_kbt(const StagedParams &P, size_t smem, cudaStream_t st) {
  static bool attrSet = false;
  if (!attrSet) {
    cudaFuncSetAttribute(k_eval_staged<K, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attrSet = true;
  }
  const int64_t tiles = (P.n + P.quizzesPerCta - 1) / P.quizzesPerCta;
  dim3 grid((unsigned)P.kb.Q, (unsigned)tiles);
  k_eval_staged<K, BT><<<grid, kEvalThreads, smem, st>>>(P);
  count_launch();
}

template <int BT>
static bool dispatch_k(const StagedParams &P, size_t smem, cudaStream_t st) {
  switch (P.kb.K) {
    case 2: launch_kbt<2, BT>(P, smem, st); return true;
    case 3: launch_kbt<3, BT>(P, smem, st); return true;
    case 4: launch_kbt<4, BT>(P, smem, st); return true;
    case 5: launch_kbt<5, BT>(P, smem, st); return true;
    case 6: launch_kbt<6, BT>(P, smem, st); return true;
    case 7: launch_kbt<7, BT>(P, smem, st); return true;
    case 8: launch_kbt<8, BT>(P, smem, st); return true;
    default: return false;
  }
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
  const int BT = n >= 2 ? 2 : 1;
  const int64_t perPass = (int64_t)kEvalWarps * BT;
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
  if (BT == 2) dispatch_k<2>(P, smem, st); else dispatch_k<1>(P, smem, st);
}

} // namespace pqa
