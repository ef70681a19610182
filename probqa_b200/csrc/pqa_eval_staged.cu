// probqa_b200: the throughput question-evaluation kernels (sm_100a) and the derived KB they stream.
//
// They compute what CEEvalQsSubtaskConsider<SRDoubleNumber>::Run (CEEvalQsSubtaskConsider.cpp:41-217) computes for
// every (quiz, question) pair of a batch of concurrent quizzes, restructured for the GPU:
//
//  * Derived KB. Everything the reference derives from sA / mD per question before it touches a quiz -- invD = 1/mD,
//    the likelihood factor r[k][j] = sA*invD (:72-81, same roundings), lr = log2 r, id2 = invD^2 (numerator of the
//    "lack" term, :116-117) -- is kept in HBM next to the KB, in the order the kernels consume it: per question and
//    per 4-target vector v, dR[v][k][lane] and dL[v][{lr_0..lr_{K-1}, id2}][lane] (lane = j % 4 = the reference's
//    AVX lane). k_build_derived writes it; the engine rebuilds the questions a Train / RecordQuizTarget touched (and
//    everything after an upload / resize / gap change) before the next evaluation.
//  * One CTA owns one question i and a tile of quizzes. The question's derived slab goes to shared memory with two 1-D
//    bulk-async copies (TMA engine, cp.async.bulk + mbarrier): R first, L behind it, so pass 1 starts while L is still in
//    flight. No transform, no index arithmetic: a thread walks the slab with one running shared-memory pointer and
//    immediate offsets. Every quiz of the tile reuses the staged slab.
//  * The reference sums with a 4-lane AVX2 Kahan accumulator: target j goes to lane j%4, lanes advance in vector
//    order (SRAccumVectDbl256.h:40-46). A GPU thread owns KL of those 4 Kahan lanes of ONE quiz (KL = 2: two threads per
//    quiz, 16 quizzes per warp -- the throughput shape) and walks the vectors in order, so pass 1 reproduces the
//    reference's normaliser W_k BIT FOR BIT (4-lane Kahan sum + PreciseSum, :81-88), hence its posteriors
//    post = lik * (1/W_k) (:91,:97) and the differences post - prior of the velocity term (:119) bit for bit. This
//    matters: for a question that is uninformative under the current posterior, sum (post - prior)^2 is pure rounding
//    noise and the reference's priority depends on it. All threads of a warp read the same slab vector: shared-memory
//    loads are 128-bit warp broadcasts.
//  * Pass 2 needs log2(posterior) per element (:106): instead of the reference's Log2Hot (one IEEE divide + series) it
//    uses log2(post) = lr[k][j] + log2(prior[j]) - log2(W_k) (two adds; log2 prior is kept per quiz), and falls back to
//    the bit-faithful Log2Hot for the elements where that split would lose accuracy or where Log2Hot's edge semantics
//    matter: post >= 0.5 (cancellation; also Log2Hot(1) = -6.56e-20 != 0) and post < 2^-1022 (Log2Hot(0) = -1023,
//    subnormals). The range test is two 3-input integer min / max trees over the high words of the step's posteriors.
//    The lack term sum_k id2 / log2(post_k) of one target is ONE division: the K reciprocals are folded into a single
//    fraction n/d (2(K-1) multiply-adds: n/d + 1/l = (n*l + d)/(d*l); all terms same-signed, |d| <= 1022^K) and d is
//    inverted with a MUFU seed + one Newton step. The entropy / lack / velocity sums are plain sums (the reference uses
//    Kahan sums): their terms are same-signed. Net: W_k bit-exact, H_k / V_k / lack / priority within the tolerance
//    stated in DESIGN.md and enforced by tests/; the fully bit-level path is k_eval_exact in pqa_kernels.cu.
//  * When a slab does not fit in shared memory (large T) the targets stream through a two-stage ring of chunks filled by
//    bulk copies one chunk ahead of the arithmetic: pass 1 over all chunks (R only; the Kahan state lives in registers
//    across chunks), then pass 2 (R and L).
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"
#include "pqa_select.cuh"

#include <atomic>
#include <cstdlib>
#include <math.h>
#include <stdio.h>
#include <stdexcept>

namespace pqa {
void count_launch();

// ---------------------------------------------------------------------------------------------------------
// Derived KB builder: one CTA row per question (all questions, or the listed local question indices), threads over
// targets. Gap / padding lanes: r = 0, id2 = 0 (the +0 of the reference's gap masks), lr = -inf (never used: a zero
// posterior takes the Log2Hot path).
template <int K>
__global__ void __launch_bounds__(256) k_build_derived(const DeviceKB kb, const int64_t *__restrict__ list) {
  const int64_t iLocal = list ? list[blockIdx.x] : (int64_t)blockIdx.x;
  const int64_t Tp = kb.Tp, T = kb.T, nV = Tp >> 2;
  const double *__restrict__ mD = kb.mD + iLocal * Tp;
  const double *__restrict__ sA = kb.sA + iLocal * K * Tp;
  double *oR = kb.dR + iLocal * nV * (K * 4);
  double *oL = kb.dL + iLocal * nV * ((K + 1) * 4);
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < Tp; j += (int64_t)gridDim.y * blockDim.x) {
    const bool gap = j >= T || bit32(kb.tgaps, j);
    const double invD = __ddiv_rn(1.0, mD[j]);                            // CEEvalQsSubtaskConsider.cpp:72-76
    const int64_t v = j >> 2, l = j & 3;
    oL[(v * (K + 1) + K) * 4 + l] = gap ? 0.0 : __dmul_rn(invD, invD);    // :116
#pragma unroll
    for (int k = 0; k < K; k++) {
      const double r = gap ? 0.0 : __dmul_rn(sA[k * Tp + j], invD);       // :81
      oR[(v * K + k) * 4 + l] = r;
      oL[(v * (K + 1) + k) * 4 + l] = log2(r);
    }
  }
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
    default: throw std::runtime_error("probqa_b200: the staged evaluation supports 2..8 answer options"); \
  }

void derived_kb_doubles(const DeviceKB &kb, size_t *nR, size_t *nL) {
  const size_t nV = (size_t)(kb.Tp >> 2);
  *nR = (size_t)kb.qCount * nV * (size_t)(kb.K * 4);
  *nL = (size_t)kb.qCount * nV * (size_t)((kb.K + 1) * 4);
}
void launch_build_derived(const DeviceKB &kb, const int64_t *dList, int64_t nList, cudaStream_t st) {
  if (kb.K > 8) return;    // the exact kernel serves K > 8 and reads sA / mD directly
  const int64_t rows = dList ? nList : kb.qCount;
  if (rows <= 0) return;
  int64_t gy = (kb.Tp + 255) / 256;
  if (gy > 64) gy = 64;
  for (int64_t r0 = 0; r0 < rows; r0 += 32768) {
    const int64_t nr = rows - r0 < 32768 ? rows - r0 : 32768;
    DeviceKB part = kb;
    const int64_t *lst = dList ? dList + r0 : nullptr;
    if (!dList) {   // all questions: shift the views instead of passing a list
      part.sA += r0 * kb.K * kb.Tp; part.mD += r0 * kb.Tp;
      part.dR += r0 * (kb.Tp >> 2) * (kb.K * 4); part.dL += r0 * (kb.Tp >> 2) * ((kb.K + 1) * 4);
    }
    dim3 grid((unsigned)nr, (unsigned)gy);
    PQA_K_SWITCH(kb.K, (k_build_derived<KK><<<grid, 256, 0, st>>>(part, lst)))
    count_launch();
  }
}

// ---------------------------------------------------------------------------------------------------------
struct StagedParams {
  DeviceKB kb;
  QuizPool qp;
  int64_t n;
  const int64_t *slots;
  double *priority;
  EvalDetail det;
  int64_t Vc;             // 4-target vectors per shared-memory chunk
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

// Thread 0: bulk copies of chunk c of question iLocal's derived slab into one stage (R, and L when pass 2 needs it),
// completing on `bar`. Sizes and addresses are multiples of 32 bytes.
template <int K>
__device__ __forceinline__ void issue_chunk(const StagedParams &P, int64_t iLocal, int64_t c, bool withL, double *sR,
                                            double *sL, uint64_t *bar) {
  const int64_t nV = P.kb.Tp >> 2;
  const int64_t v0 = c * P.Vc;
  const int64_t nv = (nV - v0 < P.Vc) ? (nV - v0) : P.Vc;
  const uint32_t bytesR = (uint32_t)(nv * (K * 32)), bytesL = (uint32_t)(nv * ((K + 1) * 32));
  mbar_arrive_expect_tx(bar, withL ? bytesR + bytesL : bytesR);
  bulk_g2s(sR, P.kb.dR + (iLocal * nV + v0) * (K * 4), bytesR, bar);
  if (withL) bulk_g2s(sL, P.kb.dL + (iLocal * nV + v0) * ((K + 1) * 4), bytesL, bar);
}

// Pass 1 over one chunk: the thread owns Kahan lanes l0 .. l0+KL-1 of its quiz. Padding / gap lanes hold prior = +0 and
// r = 0, which add +0 exactly as the reference's masked lanes do. sR: the chunk's [v][k][lane]; prc: the quiz' priors
// at (first target of the chunk) + l0. The priors come from L2 (every quiz has its own row) one vector ahead; the read
// past the chunk's last vector stays inside the quiz pool (rows are contiguous, the pool has a vector of slack).
template <int K, int KL>
__device__ __forceinline__ void pass1_chunk(const double *__restrict__ sR, int nVects, const double *__restrict__ prc,
                                            int l0, Kahan (&kw)[KL][K]) {
  const double *rp = sR + l0;
  VecD<KL> pn = ldg_vec<KL>(prc);
#pragma unroll 2
  for (int v = 0; v < nVects; v++) {
    const VecD<KL> p = pn;
    pn = ldg_vec<KL>(prc + 4 * (v + 1));
#pragma unroll
    for (int k = 0; k < K; k++) {
      const VecD<KL> r = lds_vec<KL>(rp + k * 4);
#pragma unroll
      for (int e = 0; e < KL; e++) kw[e][k].add(__dmul_rn(r.v[e], p.v[e]));   // :81-86
    }
    rp += 4 * K;
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
constexpr int kFastLo = 0x00100000, kFastHi = 0x3FE00000;
__device__ __forceinline__ bool in_fast_range(double post) {
  return (unsigned)__double2hiint(post) - (unsigned)kFastLo < (unsigned)(kFastHi - kFastLo);
}
template <int N> __device__ __forceinline__ int max_tree(const int (&h)[N]) {
  if constexpr (N == 1) return h[0];
  else if constexpr (N == 2) return max(h[0], h[1]);
  else {
    constexpr int M = (N + 2) / 3;
    int g[M];
#pragma unroll
    for (int i = 0; i < M; i++)
      g[i] = (3 * i + 2 < N) ? __vimax3_s32(h[3 * i], h[(3 * i + 1) % N], h[(3 * i + 2) % N])
                             : (3 * i + 1 < N) ? max(h[3 * i], h[(3 * i + 1) % N]) : h[3 * i];
    return max_tree<M>(g);
  }
}
template <int N> __device__ __forceinline__ int min_tree(const int (&h)[N]) {
  if constexpr (N == 1) return h[0];
  else if constexpr (N == 2) return min(h[0], h[1]);
  else {
    constexpr int M = (N + 2) / 3;
    int g[M];
#pragma unroll
    for (int i = 0; i < M; i++)
      g[i] = (3 * i + 2 < N) ? __vimin3_s32(h[3 * i], h[(3 * i + 1) % N], h[(3 * i + 2) % N])
                             : (3 * i + 1 < N) ? min(h[3 * i], h[(3 * i + 1) % N]) : h[3 * i];
    return min_tree<M>(g);
  }
}

// One vector step of pass 2 with every element classified on its own: the split logarithm where the posterior is in the
// fast range, the reference's Log2Hot (and an IEEE reciprocal) elsewhere. Only for the rare steps that hold an element
// outside the fast range.
template <int K, int KL, bool WITH_V>
__device__ __forceinline__ void pass2_careful_step(const double *__restrict__ rp, const double *__restrict__ lp_s,
                                                   const double *__restrict__ prv, const double *__restrict__ lprv,
                                                   const double *__restrict__ tbl, const double (&iW)[K], const double (&lW)[K],
                                                   double (&H)[K], double (&V)[K], double (&L)[KL]) {
  const VecD<KL> p = ldg_vec<KL>(prv), lp = ldg_vec<KL>(lprv);
  const VecD<KL> id2 = lds_vec<KL>(lp_s + K * 4);
#pragma unroll
  for (int k = 0; k < K; k++) {
    const VecD<KL> r = lds_vec<KL>(rp + k * 4), lr = lds_vec<KL>(lp_s + k * 4);
#pragma unroll
    for (int e = 0; e < KL; e++) {
      const double post = __dmul_rn(__dmul_rn(r.v[e], p.v[e]), iW[k]);  // :81-82, :97
      double l2, rl2;
      if (in_fast_range(post)) {
        l2 = __dsub_rn(__dadd_rn(lr.v[e], lp.v[e]), lW[k]);
        rl2 = fast_rcp(l2);
      } else {
        l2 = slow_log2(post, tbl, &rl2);
      }
      H[k] = __fma_rn(post, l2, H[k]);
      L[e] = __fma_rn(id2.v[e], rl2, L[e]);
      if (WITH_V) {
        const double d = __dsub_rn(post, p.v[e]);
        V[k] = __fma_rn(d, d, V[k]);
      }
    }
  }
}

// Pass 2 over one chunk. sR / sL: the chunk's [v][k][lane] and [v][{lr_k, id2}][lane]; prc / lprc as in pass 1.
// A step whose K*KL posteriors are all in the fast range runs the branch-free common path. The others (on a trained KB
// typically one step per quiz and question: the target the question is "about") are only noted -- up to six step numbers
// packed into one 64-bit register -- and evaluated when the list is full or the chunk ends: a thread with an exception
// idles for one step instead of holding its warp for the length of the careful path, and the careful steps of a warp's
// quizzes run side by side.
constexpr int kDeferBits = 10, kMaxDeferred = 6;     // chunks hold fewer than 2^10 vectors (slab_geometry)
template <int K, int KL>
__device__ __forceinline__ void pass2_chunk(const double *__restrict__ sR, const double *__restrict__ sL, int nVects,
                                            const double *__restrict__ prc, const double *__restrict__ lprc,
                                            const double *__restrict__ tbl, int l0, const double (&iW)[K],
                                            const double (&lW)[K], double (&H)[K], double (&V)[K], double (&L)[KL]) {
  const double *rp = sR + l0, *lp_s = sL + l0;
  int v = 0;
  for (;;) {
    unsigned long long deferred = 0ull;
    int nDeferred = 0;
    VecD<KL> pn = ldg_vec<KL>(prc + 4 * v), lpn = ldg_vec<KL>(lprc + 4 * v);
#pragma unroll 1
    for (; v < nVects; v++) {
      const VecD<KL> p = pn, lp = lpn;
      pn = ldg_vec<KL>(prc + 4 * (v + 1)); lpn = ldg_vec<KL>(lprc + 4 * (v + 1));
      // Everything that does not depend on the range test comes before it, so that the test's latency is covered: the
      // posteriors, a = log2 r + log2 prior, and the velocity term (the same on both paths).
      double post[K][KL], a[K][KL];
      int hi[K * KL];
#pragma unroll
      for (int k = 0; k < K; k++) {
        const VecD<KL> r = lds_vec<KL>(rp + k * 4);
        const VecD<KL> lr = lds_vec<KL>(lp_s + k * 4);
#pragma unroll
        for (int e = 0; e < KL; e++) {
          post[k][e] = __dmul_rn(__dmul_rn(r.v[e], p.v[e]), iW[k]);     // :81-82, :97
          hi[k * KL + e] = __double2hiint(post[k][e]);
          a[k][e] = __dadd_rn(lr.v[e], lp.v[e]);
          const double d = __dsub_rn(post[k][e], p.v[e]);               // :119
          V[k] = __fma_rn(d, d, V[k]);                                  // :126-127
        }
      }
      const VecD<KL> id2 = lds_vec<KL>(lp_s + K * 4);
      rp += 4 * K; lp_s += 4 * (K + 1);
#if defined(PQA_EXP_NOCHECK)
      const bool allFast = true;
#else
      const bool allFast = (min_tree<K * KL>(hi) >= kFastLo) & (max_tree<K * KL>(hi) < kFastHi);   // one branch, not two
#endif
      if (allFast) {
        // common case: no masks, no selects
        double l2[K][KL];
#pragma unroll
        for (int k = 0; k < K; k++) {
#pragma unroll
          for (int e = 0; e < KL; e++) {
            l2[k][e] = __dsub_rn(a[k][e], lW[k]);
            H[k] = __fma_rn(post[k][e], l2[k][e], H[k]);                // :113-114
          }
        }
#pragma unroll
        for (int e = 0; e < KL; e++) {
          // sum_k 1/l2[k] as one fraction nn/dd (:116-117 divides K times)
          double nn = __dadd_rn(l2[0][e], l2[1][e]), dd = __dmul_rn(l2[0][e], l2[1][e]);
#pragma unroll
          for (int k = 2; k < K; k++) { nn = __fma_rn(nn, l2[k][e], dd); dd = __dmul_rn(dd, l2[k][e]); }
          L[e] = __fma_rn(id2.v[e], __dmul_rn(nn, fast_rcp(dd)), L[e]);
        }
      } else {
        deferred = (deferred << kDeferBits) | (unsigned)v;
        if (++nDeferred == kMaxDeferred) { v++; break; }
      }
    }
#pragma unroll 1
    for (; nDeferred > 0; nDeferred--, deferred >>= kDeferBits) {
      const int dv = (int)(deferred & ((1u << kDeferBits) - 1u));
      pass2_careful_step<K, KL, false>(sR + l0 + dv * (4 * K), sL + l0 + dv * (4 * (K + 1)), prc + 4 * dv, lprc + 4 * dv, tbl,
                                       iW, lW, H, V, L);
    }
    if (v >= nVects) break;
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
__device__ __forceinline__ double priority_value(int64_t nValidTargets, const double (&W)[K], const double (&H)[K],
                                                 const double (&V)[K], double L) {
  double totW = 0.0, sumH = 0.0, sumV = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    totW += W[k];                                                       // :89,:134
    sumH = __fma_rn(W[k], -H[k], sumH);                                 // :148-172
    sumV = __fma_rn(W[k], sqrt(V[k]), sumV);
  }
  const double avgH = sumH / totW, avgV = sumV / totW;                  // :176-177
  const double nExp = exp2(avgH);                                       // :181
  const double cLnMaxV = 0.34657359027997265470861606072909;            // SRMath::_cLnSqrt2
  const double lnV = (avgV == 0) ? -746.0 : log(avgV);                  // :27-29
  const double n1 = (double)(nValidTargets + 1);
  const double vComp = 1.0 / (cLnMaxV - lnV + cLnMaxV / (n1 * n1));     // :30-33
  return -L * pow(vComp, 9.0) * pow(nExp, -2.0);                        // :201, :207
}
template <int K>
__device__ __forceinline__ void write_priority(const StagedParams &P, int64_t i, int64_t b, const double (&W)[K],
                                               const double (&H)[K], const double (&V)[K], double L) {
  const int64_t o = b * P.kb.Q + i;
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (P.det.W) P.det.W[o * K + k] = W[k];
    if (P.det.H) P.det.H[o * K + k] = -H[k];
    if (P.det.V) P.det.V[o * K + k] = V[k];
  }
  store_priority(P, o, priority_value<K>(P.kb.nValidTargets, W, H, V, L));
  if (P.det.lack) P.det.lack[o] = -L;
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

constexpr int kStages = 2;   // ring depth of the streamed (large T) shape

// Two-level summation for the streamed shape. The entropy / velocity / lack sums of pass 2 are plain sums; over tens of
// thousands of (often identical: a trained KB has long runs of equal cells) same-signed terms a single accumulator drifts
// by a fraction of an ulp of the running total per term. Each thread therefore sums one chunk in registers and adds that
// to totals kept in shared memory ([value][thread], conflict-free), so that the per-term rounding is relative to a chunk's
// sum, not to the whole row's.
template <int K, int KL, int THREADS>
__device__ __forceinline__ void flush_chunk_sums(double *sTotals, double (&H)[K], double (&V)[K], double (&L)[KL]) {
  double *o = sTotals + threadIdx.x;
#pragma unroll
  for (int k = 0; k < K; k++) {
    o[k * THREADS] = __dadd_rn(o[k * THREADS], H[k]); H[k] = 0.0;
    o[(K + k) * THREADS] = __dadd_rn(o[(K + k) * THREADS], V[k]); V[k] = 0.0;
  }
#pragma unroll
  for (int e = 0; e < KL; e++) { o[(2 * K + e) * THREADS] = __dadd_rn(o[(2 * K + e) * THREADS], L[e]); L[e] = 0.0; }
}
template <int K, int KL, int THREADS>
__device__ __forceinline__ void load_chunk_sums(const double *sTotals, double (&H)[K], double (&V)[K], double (&L)[KL]) {
  const double *o = sTotals + threadIdx.x;
#pragma unroll
  for (int k = 0; k < K; k++) { H[k] = o[k * THREADS]; V[k] = o[(K + k) * THREADS]; }
#pragma unroll
  for (int e = 0; e < KL; e++) L[e] = o[(2 * K + e) * THREADS];
}
template <int K, int KL, int THREADS> __device__ __forceinline__ void zero_chunk_sums(double *sTotals) {
#pragma unroll
  for (int x = 0; x < 2 * K + KL; x++) sTotals[x * THREADS + threadIdx.x] = 0.0;
}

// WARPS = 8: 256 threads, two CTAs per SM (slab budget 100 KB); WARPS = 4: 128 threads, four CTAs per SM (50 KB), used
// when the batch has at most 64 quizzes so that no warp of a CTA pass is idle.
template <int K, int KL, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 2) k_eval_staged(const StagedParams P) {
  constexpr int THREADS = WARPS * 32;
  constexpr int LPQ = 4 / KL;              // threads per quiz
  constexpr int QPW = 32 / LPQ;            // quizzes per warp
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bars[kStages];

  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, Tp = P.kb.Tp, nV = Tp >> 2;
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qw = lane / LPQ, l0 = (lane % LPQ) * KL;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);

  if (bit32(P.kb.qgaps, i)) {          // CEEvalQsSubtaskConsider.cpp:54-58
    for (int64_t b = tileFirst + threadIdx.x; b < tileLimit; b += THREADS) store_priority(P, b * Q + i, qnan);
    return;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const double *__restrict__ tbl = P.kb.log2tbl;

  if (P.nChunks == 1) {
    // resident shape: the whole slab stays in shared memory for every quiz group of the tile; R completes on bars[0],
    // L (needed by pass 2 only) on bars[1]
    double *sR = (double *)smRaw, *sL = sR + nV * (K * 4);
    if (threadIdx.x == 0) {
      const uint32_t bytesR = (uint32_t)(nV * (K * 32)), bytesL = (uint32_t)(nV * ((K + 1) * 32));
      mbar_arrive_expect_tx(&bars[0], bytesR);
      bulk_g2s(sR, P.kb.dR + iLocal * nV * (K * 4), bytesR, &bars[0]);
      mbar_arrive_expect_tx(&bars[1], bytesL);
      bulk_g2s(sL, P.kb.dL + iLocal * nV * ((K + 1) * 4), bytesL, &bars[1]);
    }
    const int nVects = (int)nV;
    for (int64_t g0 = tileFirst + (int64_t)warp * QPW; g0 < tileLimit; g0 += (int64_t)WARPS * QPW) {
      const int64_t b = g0 + qw;
      bool live = b < tileLimit;
      const int64_t slot = P.slots[live ? b : tileLimit - 1];
      if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
        if (l0 == 0) store_priority(P, b * Q + i, qnan);
        live = false;
      }
      if (!live) continue;             // all threads of this quiz leave together
      const double *pr = P.qp.priors + slot * Tp + l0, *lpr = P.qp.logPriors + slot * Tp + l0;
      double W[K], iW[K], lW[K], H[K], V[K], L[KL];
      {
        Kahan kw[KL][K];
#pragma unroll
        for (int e = 0; e < KL; e++)
#pragma unroll
          for (int k = 0; k < K; k++) kw[e][k].init();
        mbar_wait(&bars[0], 0);
#if !defined(PQA_EXP_NOPASS1)
        pass1_chunk<K, KL>(sR, nVects, pr, l0, kw);
#else
        for (int e = 0; e < KL; e++) for (int k = 0; k < K; k++) kw[e][k].init(0.05);
#endif
        finish_pass1<K, KL>(kw, W, iW, lW);
      }
#pragma unroll
      for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
#pragma unroll
      for (int e = 0; e < KL; e++) L[e] = 0.0;
      mbar_wait(&bars[1], 0);
#if !defined(PQA_EXP_NOPASS2)
      pass2_chunk<K, KL>(sR, sL, nVects, pr, lpr, tbl, l0, iW, lW, H, V, L);
#endif
      finish_pass2<K, KL>(P, i, b, l0, W, H, V, L);
    }
  } else {
    // streamed shape: this thread keeps its quiz for the whole question; 2 * nChunks ring items (pass 1: R chunks,
    // pass 2: R + L chunks), item g lives in stage g % kStages and is issued when item g - kStages has been consumed
    const int64_t stageDoubles = P.Vc * ((2 * K + 1) * 4);
    double *const ring = (double *)smRaw;
    double *const sTotals = ring + kStages * stageDoubles;     // [2K + KL][THREADS] (flush_chunk_sums)
    zero_chunk_sums<K, KL, THREADS>(sTotals);
    const int64_t b = tileFirst + (int64_t)warp * QPW + qw;
    bool live = b < tileLimit;
    const int64_t slot = P.slots[live ? b : tileLimit - 1];
    if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) {
      if (l0 == 0) store_priority(P, b * Q + i, qnan);
      live = false;
    }
    const double *pr = P.qp.priors + slot * Tp + l0, *lpr = P.qp.logPriors + slot * Tp + l0;
    const int64_t total = 2 * P.nChunks;
    if (threadIdx.x == 0) {
      for (int64_t g = 0; g < kStages && g < total; g++) {
        double *sR = ring + (g % kStages) * stageDoubles;
        issue_chunk<K>(P, iLocal, g % P.nChunks, g >= P.nChunks, sR, sR + P.Vc * (K * 4), &bars[g % kStages]);
      }
    }
    // consumes ring item g (its stage is free afterwards) and issues item g + kStages into the same stage
    auto next_item = [&](int64_t g) {
      __syncthreads();  // everyone is done with the stage before the next copy overwrites it
      if (threadIdx.x == 0 && g + kStages < total) {
        fence_proxy_async_smem();
        const int64_t gn = g + kStages;
        double *dR = ring + (g % kStages) * stageDoubles;
        issue_chunk<K>(P, iLocal, gn % P.nChunks, gn >= P.nChunks, dR, dR + P.Vc * (K * 4), &bars[g % kStages]);
      }
    };
    double W[K], iW[K], lW[K];
#pragma unroll
    for (int k = 0; k < K; k++) { W[k] = 0.0; iW[k] = 0.0; lW[k] = 0.0; }
    {
      Kahan kw[KL][K];
#pragma unroll
      for (int e = 0; e < KL; e++)
#pragma unroll
        for (int k = 0; k < K; k++) kw[e][k].init();
      for (int64_t g = 0; g < P.nChunks; g++) {
        const double *sR = ring + (g % kStages) * stageDoubles;
        const int64_t v0 = g * P.Vc;
        const int nVects = (int)((nV - v0 < P.Vc) ? (nV - v0) : P.Vc);
        mbar_wait(&bars[g % kStages], (uint32_t)((g / kStages) & 1));
        if (live) pass1_chunk<K, KL>(sR, nVects, pr + 4 * v0, l0, kw);
        next_item(g);
      }
      if (live) finish_pass1<K, KL>(kw, W, iW, lW);
    }
    double H[K], V[K], L[KL];
#pragma unroll
    for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
#pragma unroll
    for (int e = 0; e < KL; e++) L[e] = 0.0;
    for (int64_t g = P.nChunks; g < total; g++) {
      const double *sR = ring + (g % kStages) * stageDoubles, *sL = sR + P.Vc * (K * 4);
      const int64_t v0 = (g - P.nChunks) * P.Vc;
      const int nVects = (int)((nV - v0 < P.Vc) ? (nV - v0) : P.Vc);
      mbar_wait(&bars[g % kStages], (uint32_t)((g / kStages) & 1));
      if (live) {
        pass2_chunk<K, KL>(sR, sL, nVects, pr + 4 * v0, lpr + 4 * v0, tbl, l0, iW, lW, H, V, L);
        flush_chunk_sums<K, KL, THREADS>(sTotals, H, V, L);
      }
      next_item(g);
    }
    if (live) load_chunk_sums<K, KL, THREADS>(sTotals, H, V, L);
    if (live) finish_pass2<K, KL>(P, i, b, l0, W, H, V, L);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Medium batches (5 ... 64 quizzes, a slab that would fit shared memory whole). One question per CTA would leave most of
// the CTA's warps without a quiz, and the shared memory of an SM holds only two whole slabs. Here a CTA takes
// PLACES / G consecutive questions (G = quiz places per question, P.quizzesPerCta; PLACES = 256 threads / threads per
// quiz) and streams their slabs through the two-stage ring in chunks of P.Vc vectors; thread group p serves quiz p % G of
// question p / G for the whole launch, with the arithmetic of the throughput kernel (KL Kahan lanes per thread, W_k
// bit-exact, pass 1 over all chunks, then pass 2). All 8 warps work whatever the batch size, and the grid shrinks with
// it. KL = 2 (two threads per quiz) for 17 ... 64 quizzes; KL = 1 (four threads per quiz: half the serial work per
// thread) up to 16, where there are too few quizzes to fill the SMs and the length of a thread's chain is what counts.
template <int K, int KL>
__global__ void __launch_bounds__(256, 2) k_eval_multi(const StagedParams P) {
  constexpr int THREADS = 256, LPQ = 4 / KL, PLACES = THREADS / LPQ;
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bars[kStages];
  const int G = (int)P.quizzesPerCta, qpc = PLACES / G;
  const int64_t Q = P.kb.Q, Tp = P.kb.Tp, nV = Tp >> 2;
  const int pair = threadIdx.x / LPQ, l0 = (threadIdx.x % LPQ) * KL;
  const int qs = pair / G;
  const int64_t b = pair % G;
  const int64_t iFirst = (int64_t)blockIdx.x * qpc;                     // first local question of this CTA
  const int nQuestions = (int)((P.kb.qCount - iFirst < qpc) ? (P.kb.qCount - iFirst) : qpc);
  const int64_t iLocal = iFirst + qs, i = P.kb.qFirst + iLocal;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  bool live = qs < nQuestions && b < P.n;
  const int64_t slot = P.slots[b < P.n ? b : P.n - 1];
  if (live && (bit32(P.kb.qgaps, i) || bit64(P.qp.asked + slot * P.qp.askedWords, i))) {   // CEEvalQsSubtaskConsider.cpp:54-58
    if (l0 == 0) store_priority(P, b * Q + i, qnan);
    live = false;
  }
  const double *pr = P.qp.priors + slot * Tp + l0, *lpr = P.qp.logPriors + slot * Tp + l0;
  const int64_t perQuestion = P.Vc * ((2 * K + 1) * 4);                  // doubles of one question's chunk: R, then L
  const int64_t stageDoubles = perQuestion * qpc;
  double *const ring = (double *)smRaw;
  const int64_t total = 2 * P.nChunks;
  // thread 0: ring item g (pass 1: chunk g of R; pass 2: chunk g - nChunks of R and L) of every question of the CTA
  auto issue_item = [&](int64_t g) {
    const int64_t c = g % P.nChunks, v0 = c * P.Vc;
    const bool withL = g >= P.nChunks;
    const int64_t nv = (nV - v0 < P.Vc) ? (nV - v0) : P.Vc;
    const uint32_t bytesR = (uint32_t)(nv * (K * 32)), bytesL = (uint32_t)(nv * ((K + 1) * 32));
    uint64_t *bar = &bars[g % kStages];
    double *stage = ring + (g % kStages) * stageDoubles;
    mbar_arrive_expect_tx(bar, (uint32_t)nQuestions * (withL ? bytesR + bytesL : bytesR));
    for (int q = 0; q < nQuestions; q++) {
      double *sR = stage + q * perQuestion;
      bulk_g2s(sR, P.kb.dR + ((iFirst + q) * nV + v0) * (K * 4), bytesR, bar);
      if (withL) bulk_g2s(sR + P.Vc * (K * 4), P.kb.dL + ((iFirst + q) * nV + v0) * ((K + 1) * 4), bytesL, bar);
    }
  };
  if (threadIdx.x == 0)
    for (int64_t g = 0; g < kStages && g < total; g++) issue_item(g);
  auto next_item = [&](int64_t g) {
    __syncthreads();  // everyone is done with the stage before the next copy overwrites it
    if (threadIdx.x == 0 && g + kStages < total) {
      fence_proxy_async_smem();
      issue_item(g + kStages);
    }
  };
  double W[K], iW[K], lW[K];
#pragma unroll
  for (int k = 0; k < K; k++) { W[k] = 0.0; iW[k] = 0.0; lW[k] = 0.0; }
  {
    Kahan kw[KL][K];
#pragma unroll
    for (int e = 0; e < KL; e++)
#pragma unroll
      for (int k = 0; k < K; k++) kw[e][k].init();
    for (int64_t g = 0; g < P.nChunks; g++) {
      const double *sR = ring + (g % kStages) * stageDoubles + qs * perQuestion;
      const int64_t v0 = g * P.Vc;
      const int nVects = (int)((nV - v0 < P.Vc) ? (nV - v0) : P.Vc);
      mbar_wait(&bars[g % kStages], (uint32_t)((g / kStages) & 1));
      if (live) pass1_chunk<K, KL>(sR, nVects, pr + 4 * v0, l0, kw);
      next_item(g);
    }
    if (live) finish_pass1<K, KL>(kw, W, iW, lW);      // the threads of a quiz are live together
  }
  double H[K], V[K], L[KL];
#pragma unroll
  for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
#pragma unroll
  for (int e = 0; e < KL; e++) L[e] = 0.0;
  for (int64_t g = P.nChunks; g < total; g++) {
    const double *sR = ring + (g % kStages) * stageDoubles + qs * perQuestion, *sL = sR + P.Vc * (K * 4);
    const int64_t v0 = (g - P.nChunks) * P.Vc;
    const int nVects = (int)((nV - v0 < P.Vc) ? (nV - v0) : P.Vc);
    mbar_wait(&bars[g % kStages], (uint32_t)((g / kStages) & 1));
    if (live) pass2_chunk<K, KL>(sR, sL, nVects, pr + 4 * v0, lpr + 4 * v0, P.kb.log2tbl, l0, iW, lW, H, V, L);
    next_item(g);
  }
  if (live) finish_pass2<K, KL>(P, i, b, l0, W, H, V, L);
}

// ---------------------------------------------------------------------------------------------------------
// Small batches (fewer quizzes than one warp of the kernel above would hold): latency matters more than throughput.
// One CTA per question handles SW quizzes per round. Pass 1 cannot be spread over targets (the reference's Kahan order
// is a serial dependency per (answer, Kahan lane)): warp w runs the 4K chains of quiz w on lanes 4k + l -- W_k stays
// bit-exact. Pass 2 has no such dependency: for every quiz of the round ALL threads of the CTA stride over the targets
// with the throughput kernel's split logarithm, warp butterflies leave per-warp partial sums in shared memory, and
// warp w finishes quiz w. Whole slab in shared memory only.
template <int K, int SW>
__global__ void __launch_bounds__(SW * 32, 2) k_eval_small(const StagedParams P) {
  constexpr int THREADS = SW * 32;
  constexpr int NV = 2 * K + 1;                       // H_k, V_k, L
  extern __shared__ __align__(128) unsigned char smRaw[];
  __shared__ uint64_t bars[2];
  __shared__ double sWk[SW][3][K];                    // per quiz of the round: W_k, 1/W_k, log2 W_k
  __shared__ double sPart[SW][SW][NV];                // [quiz][warp][value] partial sums of pass 2
  __shared__ int64_t sSlot[SW];                       // slot of the quiz, or -1 when absent / already asked
  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, Tp = P.kb.Tp, T = P.kb.T, nV = Tp >> 2;
  double *sR = (double *)smRaw, *sL = sR + nV * (K * 4);
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  if (bit32(P.kb.qgaps, i)) {
    for (int64_t b = tileFirst + threadIdx.x; b < tileLimit; b += THREADS) store_priority(P, b * Q + i, qnan);
    return;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init();
    const uint32_t bytesR = (uint32_t)(nV * (K * 32)), bytesL = (uint32_t)(nV * ((K + 1) * 32));
    mbar_arrive_expect_tx(&bars[0], bytesR);
    bulk_g2s(sR, P.kb.dR + iLocal * nV * (K * 4), bytesR, &bars[0]);
    mbar_arrive_expect_tx(&bars[1], bytesL);
    bulk_g2s(sL, P.kb.dL + iLocal * nV * ((K + 1) * 4), bytesL, &bars[1]);
  }
  __syncthreads();
  const double *__restrict__ tbl = P.kb.log2tbl;
  const int nVects = (int)nV;
  for (int64_t b0 = tileFirst; b0 < tileLimit; b0 += SW) {      // CTA-uniform
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
      mbar_wait(&bars[0], 0);
      if (slot >= 0 && lane < 4 * K) {
        const int k = lane >> 2, l = lane & 3;
        const double *rk = sR + k * 4 + l;
        const double *__restrict__ prl = P.qp.priors + slot * Tp + l;
        Kahan kw; kw.init();
#pragma unroll 4
        for (int v = 0; v < nVects; v++) kw.add(__dmul_rn(rk[v * (4 * K)], __ldg(prl + 4 * v)));   // :81-86
        const double w = group_precise_sum(kw);                          // :88
        if (l == 0) { sWk[warp][0][k] = w; sWk[warp][1][k] = __ddiv_rn(1.0, w); sWk[warp][2][k] = log2(w); }
      }
    }
    __syncthreads();
    mbar_wait(&bars[1], 0);
    // ---- pass 2: every quiz of the round, all threads over the targets
    for (int q = 0; q < SW; q++) {
      const int64_t slot = sSlot[q];
      if (slot < 0) continue;                                            // CTA-uniform
      const double *__restrict__ pr = P.qp.priors + slot * Tp;
      const double *__restrict__ lpr = P.qp.logPriors + slot * Tp;
      double iW[K], lW[K], H[K], V[K], L = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { iW[k] = sWk[q][1][k]; lW[k] = sWk[q][2][k]; H[k] = 0.0; V[k] = 0.0; }
      for (int j = threadIdx.x; j < (int)T; j += THREADS) {
        const int vb = (j >> 2), l = j & 3;
        const double *rj = sR + vb * (4 * K) + l, *lj = sL + vb * (4 * (K + 1)) + l;
        const double p = __ldg(pr + j), id2 = lj[4 * K], lp = __ldg(lpr + j);
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double post = __dmul_rn(__dmul_rn(rj[4 * k], p), iW[k]);   // :81-82, :97
          double l2, rl2;
          if (in_fast_range(post)) {                                     // the throughput kernel's split logarithm
            l2 = __dsub_rn(__dadd_rn(lj[4 * k], lp), lW[k]);
            rl2 = fast_rcp46(l2);
          } else {
            l2 = slow_log2(post, tbl, &rl2);
          }
          H[k] = __fma_rn(post, l2, H[k]);                              // :113-114
          L = __fma_rn(id2, rl2, L);                                    // :116-117
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
      for (int w = 0; w < SW; w++) {
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
// One to four quizzes per call -- the shape of the reference ABI's PqaEngine_NextQuestion -- is a latency problem: the
// reference's W_k is a serial chain of T/4 Kahan steps per (answer, lane) (~4.6 us of dependent fp64 adds at T = 1000)
// and nothing can shorten it, so everything else is arranged around it. One CTA per question, every question of the KB
// in flight at once (no shared-memory slab: the derived KB streams from L2 with coalesced loads, eight CTAs per SM); warp
// w runs the 4K chains of quiz w; then all threads take the elementwise pass of each quiz; the last CTA to finish a quiz
// (a ticket counter) runs the selection of CpuEngine::NextQuestionSpec right there, and the chosen questions go to mapped
// host memory followed by a sequence word the caller polls: ONE launch, no copy, no stream synchronisation.
constexpr int kFewMax = 4;                 // quizzes per CTA = warps per CTA
constexpr int kFewBatchMax = 64;           // quizzes per launch (grid.y tiles of kFewMax quizzes)
constexpr int kFewGroups = 32;             // first-level ticket counters per quiz (1000 same-address atomics would serialise)
struct FewParams {
  DeviceKB kb;
  QuizPool qp;
  int n;
  int64_t slots[kFewMax];    // n <= kFewMax: ids and draws travel in the launch parameters ...
  uint64_t randoms[kFewMax];
  const int64_t *dSlots;     // ... larger batches: device arrays [n] (nullptr otherwise)
  const uint64_t *dRandoms;
  double *priority;          // [n][Q]
  double *runLength;         // [n][Q]
  unsigned *tickets;         // [kFewBatchMax][kFewGroups + 1] + 1 counters zeroed at creation; each is reset by its last incrementer
  int64_t *hostQuestions;    // mapped pinned host memory [kFewBatchMax]
  uint64_t *hostSeq;         // mapped pinned host memory: receives `seq` after the questions
  uint64_t seq;
  int W;
  int setActive;             // 1: NextQuestion (the quiz' active question is set); 0: evaluation only (inspection)
  double *grandOut;          // optional [n][nChunks] grand totals (inspection)
};

template <int K>
__global__ void __launch_bounds__(kFewMax * 32, 7) k_eval_few(const FewParams P) {
  constexpr int THREADS = kFewMax * 32;
  constexpr int NV = 2 * K + 1;
  extern __shared__ double sGrand[];                      // selection scratch (nChunks doubles), last CTA of a quiz only
  __shared__ double sW[kFewMax][3][K];                    // W_k, 1/W_k, log2 W_k per quiz
  __shared__ double sPart[kFewMax][NV];                   // per-warp partial sums of the quiz in flight
  __shared__ int sLast;
  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, Tp = P.kb.Tp, T = P.kb.T, nV = Tp >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double qnan = __longlong_as_double(0x7FF8000000000000ll);
  const double *__restrict__ gR = P.kb.dR + iLocal * nV * (K * 4);
  const double *__restrict__ gL = P.kb.dL + iLocal * nV * ((K + 1) * 4);
  const bool qgap = bit32(P.kb.qgaps, i);
  const int b0 = blockIdx.y * kFewMax;                    // this CTA's quizzes: b0 .. b0 + nHere - 1 of the batch
  const int nHere = P.n - b0 < kFewMax ? P.n - b0 : kFewMax;
  auto slot_of = [&](int b) -> int64_t { return P.dSlots ? P.dSlots[b0 + b] : P.slots[b]; };
  // ---- pass 1: warp w <-> quiz w, lane 4k + l <-> Kahan lane l of answer k
  if (warp < nHere && !qgap) {
    const int64_t slot = slot_of(warp);
    if (!bit64(P.qp.asked + slot * P.qp.askedWords, i) && lane < 4 * K) {
      const int k = lane >> 2, l = lane & 3;
      const double *rk = gR + k * 4 + l;
      const double *__restrict__ prl = P.qp.priors + slot * Tp + l;
      Kahan kw; kw.init();
      // The chain advances one vector per ~36 cycles (4 dependent adds); an L2 hit takes ~250: sixteen independent loads
      // are in flight ahead of the dependent adds.
      constexpr int kB = 16;
      int v = 0;
      for (; v + kB <= (int)nV; v += kB) {
        double r[kB], p[kB];
#pragma unroll
        for (int u = 0; u < kB; u++) { r[u] = __ldg(rk + (v + u) * (4 * K)); p[u] = __ldg(prl + 4 * (v + u)); }
#pragma unroll
        for (int u = 0; u < kB; u++) kw.add(__dmul_rn(r[u], p[u]));      // :81-86
      }
      for (; v < (int)nV; v++) kw.add(__dmul_rn(__ldg(rk + v * (4 * K)), __ldg(prl + 4 * v)));
      const double w = group_precise_sum(kw);                            // :88
      if (l == 0) { sW[warp][0][k] = w; sW[warp][1][k] = __ddiv_rn(1.0, w); sW[warp][2][k] = log2(w); }
    }
  }
  __syncthreads();
  // ---- pass 2 and epilogue, quiz by quiz, all threads over the targets
  for (int b = 0; b < nHere; b++) {
    const int64_t slot = slot_of(b);
    const int64_t o = (int64_t)(b0 + b) * Q + i;
    const bool live = !qgap && !bit64(P.qp.asked + slot * P.qp.askedWords, i);     // CTA-uniform
    if (live) {
      const double *__restrict__ pr = P.qp.priors + slot * Tp;
      const double *__restrict__ lpr = P.qp.logPriors + slot * Tp;
      double H[K], V[K], L = 0.0;
#pragma unroll
      for (int k = 0; k < K; k++) { H[k] = 0.0; V[k] = 0.0; }
      for (int j = threadIdx.x; j < (int)T; j += THREADS) {
        const int vb = j >> 2, l = j & 3;
        const double *rj = gR + vb * (4 * K) + l, *lj = gL + vb * (4 * (K + 1)) + l;
        const double p = __ldg(pr + j), lp = __ldg(lpr + j), id2 = __ldg(lj + 4 * K);
        double post[K], l2[K];
        bool allFast = true;
#pragma unroll
        for (int k = 0; k < K; k++) {
          post[k] = __dmul_rn(__dmul_rn(__ldg(rj + 4 * k), p), sW[b][1][k]);   // :81-82, :97
          l2[k] = __dsub_rn(__dadd_rn(__ldg(lj + 4 * k), lp), sW[b][2][k]);
          allFast = allFast & in_fast_range(post[k]);
          const double d = __dsub_rn(post[k], p);                        // :119
          V[k] = __fma_rn(d, d, V[k]);                                   // :126-127
        }
        if (allFast) {
          double nn = __dadd_rn(l2[0], l2[1]), dd = __dmul_rn(l2[0], l2[1]);
#pragma unroll
          for (int k = 0; k < K; k++) H[k] = __fma_rn(post[k], l2[k], H[k]);   // :113-114
#pragma unroll
          for (int k = 2; k < K; k++) { nn = __fma_rn(nn, l2[k], dd); dd = __dmul_rn(dd, l2[k]); }
          L = __fma_rn(id2, __dmul_rn(nn, fast_rcp(dd)), L);             // :116-117
        } else {
#pragma unroll
          for (int k = 0; k < K; k++) {
            double x = l2[k], rx;
            if (in_fast_range(post[k])) rx = fast_rcp(x);
            else x = slow_log2(post[k], P.kb.log2tbl, &rx);
            H[k] = __fma_rn(post[k], x, H[k]);
            L = __fma_rn(id2, rx, L);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; k++) {
        const double h = warp_sum(H[k]), vv = warp_sum(V[k]);
        if (lane == 0) { sPart[warp][k] = h; sPart[warp][K + k] = vv; }
      }
      L = warp_sum(L);
      if (lane == 0) sPart[warp][2 * K] = L;
      __syncthreads();
      if (threadIdx.x == 0) {
        double W[K], Hs[K], Vs[K], Ls = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) { W[k] = sW[b][0][k]; Hs[k] = 0.0; Vs[k] = 0.0; }
        for (int w = 0; w < kFewMax; w++) {
#pragma unroll
          for (int k = 0; k < K; k++) { Hs[k] = __dadd_rn(Hs[k], sPart[w][k]); Vs[k] = __dadd_rn(Vs[k], sPart[w][K + k]); }
          Ls = __dadd_rn(Ls, sPart[w][2 * K]);
        }
        P.priority[o] = priority_value<K>(P.kb.nValidTargets, W, Hs, Vs, Ls);
      }
    } else if (threadIdx.x == 0) {
      P.priority[o] = qnan;                                              // CEEvalQsSubtaskConsider.cpp:54-58
    }
    // ---- the last CTA of this quiz selects its question
    if (threadIdx.x == 0) {
      // two-level ticket: CTA x counts in group x % nG; the last of a group counts at the top; the last there is the last
      __threadfence();
      const unsigned nG = gridDim.x < (unsigned)kFewGroups ? gridDim.x : (unsigned)kFewGroups;
      const unsigned g = blockIdx.x % nG, groupSize = (gridDim.x - g + nG - 1) / nG;
      unsigned *cnt = P.tickets + (b0 + b) * (kFewGroups + 1);
      int last = 0;
      if (atomicAdd(cnt + g, 1u) + 1u == groupSize) {
        cnt[g] = 0u;
        __threadfence();
        if (atomicAdd(cnt + kFewGroups, 1u) + 1u == nG) { cnt[kFewGroups] = 0u; last = 1; }
      }
      sLast = last;
    }
    __syncthreads();
    if (sLast) {                                                         // CTA-uniform
      __threadfence();
      const int64_t nSel = split_count(Q, (int64_t)P.W * 8);
      const uint64_t draw = P.dRandoms ? P.dRandoms[b0 + b] : P.randoms[b];
      bool anomaly = false;
      const int64_t chosen = select_question_cta(P.kb, P.qp, slot, P.priority + (int64_t)(b0 + b) * Q, draw, P.W,
                                                 P.runLength + (int64_t)(b0 + b) * Q,
                                                 P.grandOut ? P.grandOut + (b0 + b) * nSel : nullptr,
                                                 P.hostQuestions != nullptr, P.setActive, sGrand, &anomaly);
      if (threadIdx.x == 0 && P.hostQuestions) {
        P.hostQuestions[b0 + b] = chosen;
        if (anomaly && P.kb.anomalies)       // the CTA that saw one publishes the counters beside the sequence word
          for (int a = 0; a < kAnomalyKinds; a++)
            reinterpret_cast<volatile uint64_t *>(P.hostSeq)[1 + a] = atomicAdd(P.kb.anomalies + a, 0ull);
        __threadfence_system();
        unsigned *done = P.tickets + kFewBatchMax * (kFewGroups + 1);
        if (atomicAdd(done, 1u) + 1u == (unsigned)P.n) {                 // every quiz of the call has its question
          *done = 0u;
          *reinterpret_cast<volatile uint64_t *>(P.hostSeq) = P.seq;
        }
      }
    }
    __syncthreads();       // sPart / sLast / sGrand are reused by the next quiz
  }
}

// ---------------------------------------------------------------------------------------------------------
// Target-sharded evaluation (SURVEY.md 8e "Targets"; pqa_kernels.cuh): this device holds the columns
// [tFirst, tFirst + P.kb.T) of every row (and the derived slabs of those columns). Same CTA shape and chunk ring as the
// streamed shape of k_eval_staged, split at the point where the reference needs the complete W_k (:88-91):
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
  // Layout [((i*K + k)*4 + lane)*n + b]*2 + {s, c}: for one (question, answer, lane) the quizzes of the batch are contiguous,
  // so a warp's store of one lane state is two runs of 256 bytes -- NVLink-sized writes into the next shard's inbox instead
  // of sixteen 32-byte ones. inState = the previous shard's hand-over (nullptr on the first
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
  __shared__ uint64_t bars[kStages];
  const int64_t iLocal = blockIdx.x, i = P.kb.qFirst + iLocal, Q = P.kb.Q, nV = P.kb.Tp >> 2;
  if (bit32(P.kb.qgaps, i)) return;      // the epilogue writes the NaN
  const int64_t tileFirst = (int64_t)blockIdx.y * P.quizzesPerCta;
  const int64_t tileLimit = (tileFirst + P.quizzesPerCta < P.n) ? tileFirst + P.quizzesPerCta : P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qw = lane / LPQ, l0 = (lane % LPQ) * KL;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t stageDoubles = P.Vc * ((2 * K + 1) * 4);
  double *const ring = (double *)smRaw;
  double *const sTotals = ring + (P.nChunks == 1 ? 1 : kStages) * stageDoubles;   // phase 2: [2K + KL][THREADS]
  if (PHASE == 2) zero_chunk_sums<K, KL, THREADS>(sTotals);
  const int64_t total = P.nChunks;
  if (threadIdx.x == 0) {      // the slab does not depend on the other shards: start the copies before any waiting
    for (int64_t g = 0; g < kStages && g < total; g++) {
      double *sR = ring + g * stageDoubles;
      issue_chunk<K>(P, iLocal, g, PHASE == 2, sR, sR + P.Vc * (K * 4), &bars[g]);
    }
  }
  const int64_t b = tileFirst + (int64_t)warp * QPW + qw;
  bool live = b < tileLimit;
  const int64_t slot = P.slots[live ? b : tileLimit - 1];
  if (live && bit64(P.qp.asked + slot * P.qp.askedWords, i)) live = false;
  const double *pr = P.qp.priors + slot * P.qp.Tp + TP.tFirst + l0, *lpr = P.qp.logPriors + slot * P.qp.Tp + TP.tFirst + l0;
  const int64_t o = b * Q + i;
  const int64_t pipeTile = iLocal / TP.pipe.tileQ;
  pipe_wait(TP.pipe, pipeTile);          // phase 1: the previous shard's hand-over; phase 2: the complete W_k of this tile
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
  if (PHASE == 1) {
    if (live && TP.inState != nullptr) {                 // continue the previous shard's Kahan lanes
#pragma unroll
      for (int e = 0; e < KL; e++)
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double2 sc = *reinterpret_cast<const double2 *>(TP.inState + ((((i * K + k) * 4 + l0 + e) * P.n + b) * 2));
          kw[e][k].s = sc.x; kw[e][k].c = sc.y;
        }
    }
  } else if (live) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      double w = TP.inW.p[0][o * K + k];
      for (int r = 1; r < TP.inW.n; r++) w = __dadd_rn(w, TP.inW.p[r][o * K + k]);
      W[k] = w;
      iW[k] = __ddiv_rn(1.0, w);                                         // :91
      lW[k] = log2(w);
    }
  }
  for (int64_t g = 0; g < total; g++) {
    const int s = (int)(g % kStages);
    const double *sR = ring + s * stageDoubles, *sL = sR + P.Vc * (K * 4);
    const int64_t v0 = g * P.Vc;
    const int nVects = (int)((nV - v0 < P.Vc) ? (nV - v0) : P.Vc);
    mbar_wait(&bars[s], (uint32_t)((g / kStages) & 1));
    if (live) {
      if (PHASE == 1) {
        pass1_chunk<K, KL>(sR, nVects, pr + 4 * v0, l0, kw);
      } else {
        pass2_chunk<K, KL>(sR, sL, nVects, pr + 4 * v0, lpr + 4 * v0, P.kb.log2tbl, l0, iW, lW, H, V, L);
        flush_chunk_sums<K, KL, THREADS>(sTotals, H, V, L);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && g + kStages < total) {
      fence_proxy_async_smem();
      double *dR = ring + s * stageDoubles;
      issue_chunk<K>(P, iLocal, g + kStages, PHASE == 2, dR, dR + P.Vc * (K * 4), &bars[s]);
    }
  }
  if (PHASE == 1) {
    if (live && TP.outState != nullptr) {                // hand the lanes over to the next shard
#pragma unroll
      for (int e = 0; e < KL; e++)
#pragma unroll
        for (int k = 0; k < K; k++)
          *reinterpret_cast<double2 *>(TP.outState + ((((i * K + k) * 4 + l0 + e) * P.n + b) * 2)) = make_double2(kw[e][k].s, kw[e][k].c);
    } else if (live) {
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
  } else if (live) {
    load_chunk_sums<K, KL, THREADS>(sTotals, H, V, L);
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

// function attributes are per device: several engines of one process may sit on different GPUs (ShardGroup)
template <typename KernelT> static void allow_big_smem(KernelT kernel, std::atomic<unsigned long long> &devices) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((devices.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    devices.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
  }
}

template <int K, int PHASE>
static void launch_tshard_k(TShardParams TP, size_t smem, cudaStream_t st) {
  // two threads per quiz; 128 quizzes per CTA pass (8 warps), or 64 (4 warps, twice the CTAs per SM) for small batches
  const bool wide = TP.S.n > 64;
  static std::atomic<unsigned long long> devWide{0}, devNarrow{0};
  allow_big_smem(k_eval_tshard<K, 2, 8, PHASE>, devWide);
  allow_big_smem(k_eval_tshard<K, 2, 4, PHASE>, devNarrow);
  const int64_t perPass = wide ? 128 : 64;
  TP.S.quizzesPerCta = perPass;
  dim3 grid((unsigned)TP.S.kb.qCount, (unsigned)((TP.S.n + perPass - 1) / perPass));
  if (wide) k_eval_tshard<K, 2, 8, PHASE><<<grid, 256, smem, st>>>(TP);
  else k_eval_tshard<K, 2, 4, PHASE><<<grid, 128, smem, st>>>(TP);
  count_launch();
}

// Chunk geometry: the whole local slab when it fits the budget (one chunk, resident), else a ring of kStages chunks of a
// multiple of 8 vectors (32 targets). Returns the dynamic shared memory of one CTA.
// `totals` = shared memory the streamed shape needs besides the ring (flush_chunk_sums).
static size_t slab_geometry(StagedParams &P, const EvalConfig &cfg, int64_t budget, int64_t totals) {
  const int64_t bytesPerVector = (2 * P.kb.K + 1) * 32;
  const int64_t nV = P.kb.Tp >> 2;
  int64_t Vc = cfg.chunkTargets > 0 ? (cfg.chunkTargets + 3) / 4 : nV;
  if (Vc > nV) Vc = nV;
  int64_t ringVc = ((budget - totals) / (kStages * bytesPerVector)) & ~7ll;
  if (ringVc > 1016) ringVc = 1016;                      // pass 2 packs step numbers into 10 bits
  if (Vc == nV && nV * bytesPerVector > budget) Vc = ringVc;
  if (Vc < nV && Vc > ringVc) Vc = ringVc;
  P.Vc = Vc;
  P.nChunks = (nV + Vc - 1) / Vc;
  P.quizzesPerCta = 0;
  return (size_t)((P.nChunks == 1 ? 1 : kStages) * Vc * bytesPerVector);
}
static int64_t totals_bytes(int64_t K, int KL, int threads) { return (2 * K + KL) * (int64_t)threads * 8; }
int64_t tshard_quiz_tiles(int64_t n) {
  const int64_t perPass = n > 64 ? 128 : 64;
  return (n + perPass - 1) / perPass;
}
// the target-sharded kernels always carry the totals (phase 2), resident or streamed
static size_t tshard_geometry(StagedParams &P, const EvalConfig &cfg) {
  const int64_t totals = totals_bytes(P.kb.K, 2, P.n > 64 ? 256 : 128);
  const int64_t budget = P.n > 64 ? kSlabBudgetWide : kSlabBudgetNarrow;
  size_t smem = slab_geometry(P, cfg, budget, totals);
  if (P.nChunks == 1 && (int64_t)smem + totals > budget) {      // whole slab + totals do not fit: stream it
    EvalConfig c2 = cfg;
    c2.chunkTargets = 4 * (((budget - totals) / (kStages * (2 * P.kb.K + 1) * 32)) & ~7ll);
    smem = slab_geometry(P, c2, budget, totals);
  }
  return smem + (size_t)totals;
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
  EP.S.Vc = 0; EP.S.nChunks = 0; EP.S.quizzesPerCta = 0;
  EP.inW = inW; EP.inHVL = inHVL;
  const int64_t total = n * kbLocal.Q;
  int64_t g = (total + 127) / 128;
  if (g > 148 * 16) g = 148 * 16;
  PQA_K_SWITCH(kbLocal.K, (k_tshard_priority<KK><<<(unsigned)g, 128, 0, st>>>(EP)))
  count_launch();
}

template <int K>
static void launch_small(StagedParams P, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> dev{0};
  allow_big_smem(k_eval_small<K, 8>, dev);
  P.quizzesPerCta = 32;                                    // one CTA per question, rounds of eight quizzes
  dim3 grid((unsigned)P.kb.qCount, (unsigned)((P.n + 31) / 32));
  k_eval_small<K, 8><<<grid, 8 * 32, smem, st>>>(P);
  count_launch();
}

// Medium batches: geometry and launch of k_eval_multi (see there). Only for slabs that would be resident.
constexpr int64_t kMultiRingBudget = 100 * 1024;
static bool multi_applies(const StagedParams &P, const EvalConfig &cfg) {
  return P.n > 0 && P.n <= 64 && P.nChunks == 1 && cfg.kahanLanesPerThread == 0 && cfg.chunkTargets == 0 &&
         cfg.quizzesPerCta == 0 && cfg.which != 1;
}
static int multi_min() {     // batches of this many ... 64 quizzes go to k_eval_multi (PQA_B200_MULTI_MIN: experiments)
  static const int v = [] { const char *e = std::getenv("PQA_B200_MULTI_MIN"); return e && *e ? std::atoi(e) : 5; }();
  return v;
}
template <int K, int KL>
static void launch_multi_kl(StagedParams P, int64_t G, cudaStream_t st) {
  static std::atomic<unsigned long long> dev{0};
  allow_big_smem(k_eval_multi<K, KL>, dev);
  const int64_t places = 256 / (4 / KL), qpc = places / G, nV = P.kb.Tp >> 2;
  int64_t Vc = (kMultiRingBudget / (kStages * qpc * (2 * K + 1) * 32)) & ~7ll;
  if (Vc > 1016) Vc = 1016;                                   // pass 2 packs step numbers into 10 bits
  if (Vc > nV) Vc = nV;
  P.Vc = Vc; P.nChunks = (nV + Vc - 1) / Vc; P.quizzesPerCta = G;
  const size_t smem = (size_t)(kStages * qpc * Vc * (2 * K + 1) * 32);
  const unsigned grid = (unsigned)((P.kb.qCount + qpc - 1) / qpc);
  k_eval_multi<K, KL><<<grid, 256, smem, st>>>(P);
  count_launch();
}
template <int K>
static void launch_multi(const StagedParams &P, cudaStream_t st) {
  static const int kl1Max = [] { const char *e = std::getenv("PQA_B200_MULTI_KL1_MAX"); return e && *e ? std::atoi(e) : 16; }();
  if (P.n <= kl1Max && P.n <= 64) launch_multi_kl<K, 1>(P, P.n <= 8 ? 8 : P.n <= 16 ? 16 : P.n <= 32 ? 32 : 64, st);
  else launch_multi_kl<K, 2>(P, P.n <= 16 ? 16 : P.n <= 32 ? 32 : 64, st);
}

template <int K, int KL, int WARPS>
static void launch_cfg(StagedParams P, const EvalConfig &cfg, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> dev{0};
  allow_big_smem(k_eval_staged<K, KL, WARPS>, dev);
  const int64_t perPass = (int64_t)WARPS * (32 / (4 / KL));   // quizzes one CTA evaluates concurrently
  if (P.nChunks > 1) {
    smem += (size_t)totals_bytes(K, KL, WARPS * 32);
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
  if (multi_applies(P, cfg) && P.n >= multi_min()) { launch_multi<K>(P, st); return; }
  if (cfg.kahanLanesPerThread == 0 && P.n < 32 && P.nChunks == 1 && cfg.chunkTargets == 0) { launch_small<K>(P, smem, st); return; }
  // batches >= 32: two threads per quiz (16 quizzes per warp, 16 warps per SM; measured fastest); small batches: four
  // threads per quiz; one thread per quiz (4 lanes, 4 warps per CTA) is kept selectable for experiments
  const int lanesPerThread = cfg.kahanLanesPerThread > 0 ? cfg.kahanLanesPerThread : (P.n >= 32 ? 2 : 1);
  if (lanesPerThread == 4) launch_cfg<K, 4, 4>(P, cfg, smem, st);
  else if (lanesPerThread == 2 && P.n <= 64 && P.nChunks > 1) launch_cfg<K, 2, 4>(P, cfg, smem, st);   // 64 quizzes per CTA pass, 4 CTAs/SM
  else if (lanesPerThread == 2 && P.n <= 64 && cfg.kahanLanesPerThread == 0) launch_cfg<K, 1, 8>(P, cfg, smem, st);   // whole slab (2 CTAs/SM):
                                                               // four threads per quiz fill all 8 warps
  else if (lanesPerThread == 2) launch_cfg<K, 2, 8>(P, cfg, smem, st);
  else launch_cfg<K, 1, 8>(P, cfg, smem, st);
}

int eval_few_inline() { return kFewMax; }
int eval_few_max() { return kFewBatchMax; }
int eval_few_ticket_count() { return kFewBatchMax * (kFewGroups + 1) + 1; }
void launch_eval_few_select(const DeviceKB &kb, const QuizPool &qp, int n, const int64_t *slots, const uint64_t *randoms, int W,
                            double *dPriority, double *dRunLength, unsigned *dTickets, int64_t *hostQuestions,
                            uint64_t *hostSeq, uint64_t seq, double *dGrand, const int64_t *dSlots, const uint64_t *dRandoms,
                            cudaStream_t st) {
  FewParams P;
  P.dSlots = n > kFewMax ? dSlots : nullptr; P.dRandoms = n > kFewMax ? dRandoms : nullptr;
  P.kb = kb; P.qp = qp; P.n = n; P.priority = dPriority; P.runLength = dRunLength; P.tickets = dTickets;
  P.hostQuestions = hostQuestions; P.hostSeq = hostSeq; P.seq = seq; P.W = W;
  P.setActive = hostQuestions != nullptr; P.grandOut = dGrand;
  for (int x = 0; x < kFewMax; x++) { P.slots[x] = (x < n && n <= kFewMax) ? slots[x] : 0; P.randoms[x] = (x < n && n <= kFewMax) ? randoms[x] : 0; }
  const size_t smem = sizeof(double) * (size_t)select_chunk_count(kb.Q, W);
  const dim3 grid((unsigned)kb.qCount, (unsigned)((n + kFewMax - 1) / kFewMax));
  PQA_K_SWITCH(kb.K, (k_eval_few<KK><<<grid, kFewMax * 32, smem, st>>>(P)))
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
  P.kb = kb; P.qp = qp; P.n = n; P.slots = dSlots; P.priority = dPriority; P.det = det; P.mirror = cfg.mirror;
  // a slab that fits 100 KB is staged whole; a streamed one uses a 50 KB ring for batches of <= 64 quizzes, which run
  // four 4-warp CTAs per SM (launch_k)
  size_t smem = slab_geometry(P, cfg, kSlabBudgetWide, totals_bytes(kb.K, 4, 256));
  if (P.nChunks > 1 && n <= 64 && cfg.chunkTargets == 0 && (cfg.kahanLanesPerThread == 0 || cfg.kahanLanesPerThread == 2))
    smem = slab_geometry(P, cfg, kSlabBudgetNarrow, totals_bytes(kb.K, 2, 128));
  PQA_K_SWITCH(kb.K, (launch_k<KK>(P, cfg, smem, st)))
}

template <int K> static void preload_staged_k() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_build_derived<K>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 1, 8>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 2, 8>);
  cudaFuncGetAttributes(&a, k_eval_staged<K, 4, 4>);
  cudaFuncGetAttributes(&a, k_eval_small<K, 8>);
  cudaFuncGetAttributes(&a, k_eval_multi<K, 1>);
  cudaFuncGetAttributes(&a, k_eval_multi<K, 2>);
  cudaFuncGetAttributes(&a, k_eval_few<K>);
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
