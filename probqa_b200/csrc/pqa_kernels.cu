// probqa_b200: sm_100a kernels of the quiz path other than the throughput question-evaluation kernel
// (pqa_eval_staged.cu): KB fill / row packing, StartQuiz, RecordAnswer (bit-exact posterior update), the exact
// question evaluation (reference rounding order), question selection, ListTopTargets, RecordQuizTarget/Train.
// Citations are relative to /root/reference/ProbQA/. Built with -fmad=false: every fused multiply-add is explicit.
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"
#include "pqa_select.cuh"

#include <atomic>
#include <limits.h>
#include <math.h>

namespace pqa {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launch_count() { return g_launches.load(std::memory_order_relaxed); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static inline int grid_for(int64_t n, int block, int cap = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------------------
// KB initialisation (CpuEngine.cpp:34-93: sA = init^2, mD = K*init^2, vB = init) and row packing.
__global__ void k_fill_kb(DeviceKB kb, double initSqr, double initMD, double init1) {
  const int64_t nA = kb.qCount * kb.K * kb.Tp, nD = kb.qCount * kb.Tp, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < nA; x += stride)
    kb.sA[x] = (x % kb.Tp) < kb.T ? initSqr : 0.0;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < nD; x += stride)
    kb.mD[x] = (x % kb.Tp) < kb.T ? initMD : 1.0;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < kb.Tp; x += stride)
    kb.vB[x] = x < kb.T ? init1 : 0.0;
}
void launch_fill_kb(const DeviceKB &kb, double initSqr, double initMD, double init1, cudaStream_t st) {
  k_fill_kb<<<148 * 8, 256, 0, st>>>(kb, initSqr, initMD, init1);
  count_launch();
}

__global__ void k_pad_rows(double *__restrict__ dst, const double *__restrict__ src, int64_t nRows, int64_t T,
                           int64_t Tp, double padValue) {
  const int64_t n = nRows * Tp, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
    const int64_t r = x / Tp, j = x - r * Tp;
    dst[x] = j < T ? src[r * T + j] : padValue;
  }
}
void launch_pad_rows(double *dst, const double *src, int64_t nRows, int64_t T, int64_t Tp, double padValue,
                     cudaStream_t st) {
  k_pad_rows<<<grid_for(nRows * Tp, 256), 256, 0, st>>>(dst, src, nRows, T, Tp, padValue);
  count_launch();
}
__global__ void k_unpad_rows(double *__restrict__ dst, const double *__restrict__ src, int64_t nRows, int64_t T,
                             int64_t Tp) {
  const int64_t n = nRows * T, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
    const int64_t r = x / T, j = x - r * T;
    dst[x] = src[r * Tp + j];
  }
}
void launch_unpad_rows(double *dst, const double *src, int64_t nRows, int64_t T, int64_t Tp, cudaStream_t st) {
  k_unpad_rows<<<grid_for(nRows * T, 256), 256, 0, st>>>(dst, src, nRows, T, Tp);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// StartQuiz / RecordAnswer: one CTA per quiz. Bit-exact with CpuEngine for the emulated worker count W:
//   m[j]   = gap ? +0 : vB[j]                                   CESetPriorsSubtaskSum.cpp:30-35
//          | gap ? +0 : prior[j] * (sA[q][a][j] / mD[q][j])     CERecordAnswerSubtaskMul.cpp:27-37 (divide first)
//   pieces = CalcSplit(ceil(T/4) vectors, W); per piece four Kahan lanes in vector order, PreciseSum
//            (SRAccumVectDbl256.h:83-91); scalar Kahan over the piece sums (Summator.h:11-21)
//   prior[j] = m[j] / S                                         CEDivTargPriorsSubtask.h:16-21
// Shared memory: 4W lane sums + 4W lane corrections + W piece sums + 1.
// prior[0..Tp) holds the un-normalised m[j] (padding lanes +0). Sums it in the reference's order -- pieces =
// CalcSplit(ceil(T/4) vectors, W), four sequential Kahan lanes per piece, PreciseSum, scalar Kahan over the pieces
// (SRPoolRunner.h:96-110, SRAccumVectDbl256.h:40-46,83-91, Summator.h:11-21) -- divides (CEDivTargPriorsSubtask.h:16-21)
// and refreshes log2 prior. Called by every thread of the CTA; sm: 9W+1 doubles of shared memory.
__device__ void kahan_normalise(double *prior, double *lprior, int64_t T, int64_t Tp, int W, double *sm) {
  double *laneS = sm, *laneC = sm + 4 * W, *pieceSum = sm + 8 * W, *total = sm + 9 * W;
  const int64_t nVects = Tp >> 2;
  const int64_t nPieces = split_count(nVects, W);
  for (int64_t t = threadIdx.x; t < nPieces * 4; t += blockDim.x) {
    const int64_t p = t >> 2, lane = t & 3;
    const int64_t first = split_start(nVects, W, p), limit = split_start(nVects, W, p + 1);
    Kahan k; k.init();
    for (int64_t v = first; v < limit; v++) k.add(prior[4 * v + lane]);
    laneS[t] = k.s; laneC[t] = k.c;
  }
  __syncthreads();
  for (int64_t p = threadIdx.x; p < nPieces; p += blockDim.x)
    pieceSum[p] = precise_sum4(laneS[4 * p], laneS[4 * p + 1], laneS[4 * p + 2], laneS[4 * p + 3],
                               laneC[4 * p], laneC[4 * p + 1], laneC[4 * p + 2], laneC[4 * p + 3]);
  __syncthreads();
  if (threadIdx.x == 0) {
    Kahan k; k.init(0.0);
    for (int64_t p = 0; p < nPieces; p++) k.add(pieceSum[p]);
    total[0] = k.get();
  }
  __syncthreads();
  const double S = total[0];
  for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) {
    const double v = j < T ? __ddiv_rn(prior[j], S) : 0.0;
    prior[j] = v;
    lprior[j] = log2(v);
  }
}

template <int MODE>  // 0 = StartQuiz, 1 = RecordAnswer
__global__ void __launch_bounds__(256) k_update_priors(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                                       const int64_t *__restrict__ answers, int W) {
  extern __shared__ double sm[];
  const int64_t slot = slots[blockIdx.x];
  double *prior = qp.priors + slot * qp.Tp;
  double *lprior = qp.logPriors + slot * qp.Tp;
  const int64_t T = kb.T, Tp = kb.Tp;
  const double *rowA = nullptr, *rowD = nullptr;
  int64_t q = -1;
  if (MODE == 1) {
    q = qp.active[slot];
    if (q < 0) return;      // no active question: the host validation never lets this through; never touch the slot
    if (q < kb.qFirst || q >= kb.qFirst + kb.qCount) {
      // question-sharded engine, question owned by another device: contribute zeros to the all-reduce of the
      // updated priors (x + 0 is exact), keep the bookkeeping of CEQuiz.h:90-92 identical on every device
      for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) prior[j] = 0.0;
      if (threadIdx.x == 0) {
        qp.asked[slot * qp.askedWords + (q >> 6)] |= 1ull << (q & 63);
        qp.active[slot] = -1;
      }
      return;
    }
    const int64_t a = answers[blockIdx.x];
    rowA = kb.sA + ((q - kb.qFirst) * kb.K + a) * Tp;
    rowD = kb.mD + (q - kb.qFirst) * Tp;
  }
  for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) {
    double m = 0.0;
    if (j < T && !bit32(kb.tgaps, j)) {
      if (MODE == 0) m = kb.vB[j];
      else m = __dmul_rn(prior[j], __ddiv_rn(rowA[j], rowD[j]));
    }
    prior[j] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (MODE == 0) {
      for (int64_t w = 0; w < qp.askedWords; w++) qp.asked[slot * qp.askedWords + w] = 0;
      qp.active[slot] = -1;
    } else {
      qp.asked[slot * qp.askedWords + (q >> 6)] |= 1ull << (q & 63);  // CEQuiz.h:91
      qp.active[slot] = -1;                                           // CEQuiz.h:92
    }
  }
  kahan_normalise(prior, lprior, T, Tp, W, sm);
}

static size_t priors_smem(int W) { return sizeof(double) * (size_t)(9 * W + 1); }

// ---- long rows (T >= kLongRow): the same arithmetic spread over the chip. One CTA per quiz cannot feed a row of 10^5
// targets: its 4W Kahan chains wait an L2 round trip per step and the divide / log2 sweep has 256 threads. Here
//   k_update_lanes   one thread per (piece, AVX lane) chain of the reference's sum: it forms m[j] for its own elements
//                    (so the row is read once), eight elements ahead of the dependent adds, stores the un-normalised row,
//                    and the CTA finishes PreciseSum / the Kahan sum over the pieces -> S per quiz (qp.normS[slot])
//   k_normalise_rows grid over (quiz, row chunks): prior = m / S, log2 prior
// Operation for operation what k_update_priors / k_tshard_ra_finish do, hence the same bits.
constexpr int64_t kLongRow = 8192;
template <int MODE>  // 0 = StartQuiz, 1 = RecordAnswer, 2 = rows[] holds the complete un-normalised row (target shards)
__global__ void __launch_bounds__(256) k_update_lanes(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                                       const int64_t *__restrict__ answers, const double *__restrict__ rows,
                                                       int W) {
  extern __shared__ double sm[];
  double *laneS = sm, *laneC = sm + 4 * W, *pieceSum = sm + 8 * W;
  const int64_t slot = slots[blockIdx.x];
  double *prior = qp.priors + slot * qp.Tp;
  const int64_t T = kb.T, Tp = kb.Tp;
  const double *rowA = nullptr, *rowD = nullptr, *rowM = nullptr;
  int64_t q = -1;
  if (MODE != 0) q = qp.active[slot];
  if (MODE == 1) {
    if (q < 0) return;
    if (q < kb.qFirst || q >= kb.qFirst + kb.qCount) {     // question owned by another shard: see k_update_priors
      for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) prior[j] = 0.0;
      if (threadIdx.x == 0) {
        qp.asked[slot * qp.askedWords + (q >> 6)] |= 1ull << (q & 63);
        qp.active[slot] = -1;
        qp.normS[slot] = __longlong_as_double(0x7FF8000000000000ll);      // tells k_normalise_rows to leave the zeros alone
      }
      return;
    }
    const int64_t a = answers[blockIdx.x];
    rowA = kb.sA + ((q - kb.qFirst) * kb.K + a) * Tp;
    rowD = kb.mD + (q - kb.qFirst) * Tp;
  }
  if (MODE == 2) rowM = rows + (int64_t)blockIdx.x * qp.Tp;
  const int64_t nVects = Tp >> 2;
  const int64_t nPieces = split_count(nVects, W);
  // rows are padded to Tp: every load below is unconditional (so that the eight steps' loads are all in flight before the
  // first divide), the gap / padding mask is applied to the value
  const double *src = MODE == 0 ? kb.vB : MODE == 1 ? prior : rowM;
  auto value = [&](int64_t j, double x, double a, double d) -> double {
    if (j >= T || bit32(kb.tgaps, j)) return 0.0;
    return MODE == 1 ? __dmul_rn(x, __ddiv_rn(a, d)) : x;
  };
  const int64_t t = threadIdx.x;
  if (t < nPieces * 4) {
    const int64_t p = t >> 2, lane = t & 3;
    const int64_t first = split_start(nVects, W, p), limit = split_start(nVects, W, p + 1);
    Kahan k; k.init();
    int64_t v = first;
    // two batches of eight steps in flight: the loads of the next batch are issued before the dependent adds of this one
    struct Batch { double x[8], a[8], d[8]; };
    auto load = [&](Batch &bt, int64_t v0) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int64_t j = 4 * (v0 + u) + lane;
        bt.x[u] = src[j];
        bt.a[u] = MODE == 1 ? rowA[j] : 0.0;
        bt.d[u] = MODE == 1 ? rowD[j] : 1.0;
      }
    };
    auto process = [&](const Batch &bt, int64_t v0) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int64_t j = 4 * (v0 + u) + lane;
        const double m = value(j, bt.x[u], bt.a[u], bt.d[u]);
        prior[j] = m;
        k.add(m);
      }
    };
    Batch A, B;
    if (v + 8 <= limit) {
      load(A, v);
      for (;;) {
        if (v + 16 > limit) { process(A, v); v += 8; break; }
        load(B, v + 8);
        process(A, v); v += 8;
        if (v + 16 > limit) { process(B, v); v += 8; break; }
        load(A, v + 8);
        process(B, v); v += 8;
      }
    }
    for (; v < limit; v++) {
      const int64_t j = 4 * v + lane;
      const double m = value(j, src[j], MODE == 1 ? rowA[j] : 0.0, MODE == 1 ? rowD[j] : 1.0);
      prior[j] = m;
      k.add(m);
    }
    laneS[t] = k.s; laneC[t] = k.c;
  }
  __syncthreads();
  for (int64_t p = threadIdx.x; p < nPieces; p += blockDim.x)
    pieceSum[p] = precise_sum4(laneS[4 * p], laneS[4 * p + 1], laneS[4 * p + 2], laneS[4 * p + 3],
                               laneC[4 * p], laneC[4 * p + 1], laneC[4 * p + 2], laneC[4 * p + 3]);
  __syncthreads();
  if (threadIdx.x == 0) {
    Kahan k; k.init(0.0);
    for (int64_t p = 0; p < nPieces; p++) k.add(pieceSum[p]);
    qp.normS[slot] = k.get();
    if (MODE == 0) {
      for (int64_t w = 0; w < qp.askedWords; w++) qp.asked[slot * qp.askedWords + w] = 0;
      qp.active[slot] = -1;
    } else {
      qp.asked[slot * qp.askedWords + (q >> 6)] |= 1ull << (q & 63);  // CEQuiz.h:91
      qp.active[slot] = -1;                                           // CEQuiz.h:92
    }
  }
}
__global__ void __launch_bounds__(256) k_normalise_rows(QuizPool qp, const int64_t *__restrict__ slots, int64_t T) {
  const int64_t slot = slots[blockIdx.y];
  const double S = qp.normS[slot];
  if (S != S) return;                                       // foreign-shard question: the row stays zero
  double *prior = qp.priors + slot * qp.Tp, *lprior = qp.logPriors + slot * qp.Tp;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < qp.Tp; j += (int64_t)gridDim.x * blockDim.x) {
    const double v = j < T ? __ddiv_rn(prior[j], S) : 0.0;  // CEDivTargPriorsSubtask.h:16-21
    prior[j] = v;
    lprior[j] = log2(v);
  }
}
static bool long_row_path(const QuizPool &qp, int64_t Tp, int W) {
  return Tp >= kLongRow && qp.normS != nullptr && 4 * split_count(Tp >> 2, W) <= 256;
}
template <int MODE>
static void launch_update_long(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, const int64_t *dAnswers,
                               const double *dRows, int W, cudaStream_t st) {
  const int threads = (int)((4 * split_count(kb.Tp >> 2, W) + 31) / 32 * 32);
  k_update_lanes<MODE><<<(unsigned)n, threads, priors_smem(W), st>>>(kb, qp, dSlots, dAnswers, dRows, W);
  count_launch();
  int64_t gx = (kb.Tp + 255) / 256;
  if (gx > 64) gx = 64;
  k_normalise_rows<<<dim3((unsigned)gx, (unsigned)n), 256, 0, st>>>(qp, dSlots, kb.T);
  count_launch();
}


void launch_start_quiz(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, int W,
                       cudaStream_t st) {
  if (n <= 0) return;
  if (long_row_path(qp, kb.Tp, W)) { launch_update_long<0>(kb, qp, n, dSlots, nullptr, nullptr, W, st); return; }
  k_update_priors<0><<<(unsigned)n, 256, priors_smem(W), st>>>(kb, qp, dSlots, nullptr, W);
  count_launch();
}
void launch_record_answer(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                          const int64_t *dAnswers, int W, cudaStream_t st) {
  if (n <= 0) return;
  if (long_row_path(qp, kb.Tp, W)) { launch_update_long<1>(kb, qp, n, dSlots, dAnswers, nullptr, W, st); return; }
  k_update_priors<1><<<(unsigned)n, 256, priors_smem(W), st>>>(kb, qp, dSlots, dAnswers, W);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// RecordAnswer on a target shard (pqa_kernels.cuh): the elementwise half on the local columns, then -- once the row is
// complete on every shard -- the bookkeeping and the reference's Kahan normalisation on the full row. The division and
// the multiplication are the reference's (CERecordAnswerSubtaskMul.cpp:31-34), so the finished priors are bit-exact.
__global__ void __launch_bounds__(256) k_tshard_ra_partial(DeviceKB kbL, QuizPool qp, int64_t tFirst,
                                                           const int64_t *__restrict__ slots,
                                                           const int64_t *__restrict__ answers, PeerBufs out) {
  const int64_t b = blockIdx.y, slot = slots[b];
  const int64_t q = qp.active[slot], a = answers[b];
  const double *rowA = kbL.sA + ((q - kbL.qFirst) * kbL.K + a) * kbL.Tp, *rowD = kbL.mD + (q - kbL.qFirst) * kbL.Tp;
  const double *prior = qp.priors + slot * qp.Tp + tFirst;
  for (int64_t jl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; jl < kbL.T; jl += (int64_t)gridDim.x * blockDim.x) {
    const double m = bit32(kbL.tgaps, jl) ? 0.0 : __dmul_rn(prior[jl], __ddiv_rn(rowA[jl], rowD[jl]));
    for (int r = 0; r < out.n; r++) out.p[r][b * qp.Tp + tFirst + jl] = m;
  }
}
void launch_tshard_record_answer_partial(const DeviceKB &kbLocal, const QuizPool &qp, int64_t tFirst, int64_t n,
                                         const int64_t *dSlots, const int64_t *dAnswers, const PeerBufs &out,
                                         cudaStream_t st) {
  if (n <= 0) return;
  int64_t gx = (kbLocal.T + 255) / 256;
  if (gx > 64) gx = 64;
  k_tshard_ra_partial<<<dim3((unsigned)gx, (unsigned)n), 256, 0, st>>>(kbLocal, qp, tFirst, dSlots, dAnswers, out);
  count_launch();
}

__global__ void __launch_bounds__(256) k_tshard_ra_finish(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                                          const double *__restrict__ rows, int W) {
  extern __shared__ double sm[];
  const int64_t slot = slots[blockIdx.x];
  double *prior = qp.priors + slot * qp.Tp, *lprior = qp.logPriors + slot * qp.Tp;
  const double *row = rows + (int64_t)blockIdx.x * qp.Tp;
  for (int64_t j = threadIdx.x; j < kb.Tp; j += blockDim.x) prior[j] = (j < kb.T && !bit32(kb.tgaps, j)) ? row[j] : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t q = qp.active[slot];
    qp.asked[slot * qp.askedWords + (q >> 6)] |= 1ull << (q & 63);  // CEQuiz.h:91
    qp.active[slot] = -1;                                           // CEQuiz.h:92
  }
  kahan_normalise(prior, lprior, kb.T, kb.Tp, W, sm);
}
void launch_tshard_record_answer_finish(const DeviceKB &kbFull, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                        const double *dRows, int W, cudaStream_t st) {
  if (n <= 0) return;
  if (long_row_path(qp, kbFull.Tp, W)) { launch_update_long<2>(kbFull, qp, n, dSlots, nullptr, dRows, W, st); return; }
  k_tshard_ra_finish<<<(unsigned)n, 256, priors_smem(W), st>>>(kbFull, qp, dSlots, dRows, W);
  count_launch();
}

// Closed-form binary-search KB (probqa_b200/synth.py answer_rule / binary_search_kb): the same double operations in the
// same order, so the device-filled shard is bit-identical to the numpy arrays.
__global__ void k_fill_binary_search_kb(DeviceKB kb, int64_t tFirst, int64_t Tg, double init, double rounds) {
  const int64_t w = (32 * Tg) / 1000 > 1 ? (32 * Tg) / 1000 : 1;
  const int64_t nD = kb.qCount * kb.Tp, stride = (int64_t)gridDim.x * blockDim.x;
  const double hit = __dadd_rn(init, rounds), hit2 = __dmul_rn(hit, hit), miss2 = __dmul_rn(init, init);
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < nD; x += stride) {
    const int64_t il = x / kb.Tp, jl = x - il * kb.Tp;
    const int64_t i = kb.qFirst + il, j = tFirst + jl;
    const int64_t piv = (i * Tg) / kb.Q;
    int64_t ans = j < piv - w ? 0 : j < piv ? 1 : j == piv ? 2 : j <= piv + w ? 3 : 4;
    if (ans > kb.K - 1) ans = kb.K - 1;
    double d = 0.0;
    for (int64_t k = 0; k < kb.K; k++) {
      const double a = jl < kb.T ? (k == ans ? hit2 : miss2) : 0.0;
      kb.sA[(il * kb.K + k) * kb.Tp + jl] = a;
      d = __dadd_rn(d, a);
    }
    kb.mD[x] = jl < kb.T ? d : 1.0;
  }
}
__global__ void k_fill_vector(double *v, int64_t n, int64_t nValid, double value) {
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (int64_t)gridDim.x * blockDim.x)
    v[x] = x < nValid ? value : 0.0;
}
void launch_fill_binary_search_kb(const DeviceKB &kbLocal, int64_t tFirst, int64_t Tglobal, double init, double rounds,
                                  cudaStream_t st) {
  k_fill_binary_search_kb<<<148 * 8, 256, 0, st>>>(kbLocal, tFirst, Tglobal, init, rounds);
  count_launch();
  const int64_t TpG = (Tglobal + 3) & ~3ll;
  k_fill_vector<<<grid_for(TpG, 256), 256, 0, st>>>(kbLocal.vB, TpG, Tglobal, init + rounds);
  count_launch();
}

// ResumeQuiz: CECreateQuizResume::UpdateLikelihoods (CECreateQuizOperation.cpp:55-83) = CEUpdatePriorsSubtaskMul
// (CEUpdatePriorsSubtaskMul.cpp:40-82) + CpuEngine::NormalizePriors (CpuEngine.cpp:284-335; CENormPriorsSubtaskMax.cpp,
// CENormPriorsSubtaskCorrSum.cpp). One CTA per quiz; bit-exact. The likelihood product over the answered questions is
// kept as a mantissa in [1,2) plus an integer sum of biased exponents so that it cannot underflow; the reference
// multiplies by vB[j % 4] (it loads the FIRST vector of vB for every target vector, :48) and that is reproduced.
// status[b] = 1 reports the reference's I64Underflow (:316-319). The log-prior row doubles as the exponent scratch.
// Generalised over where the cells live (ResumeSource): one engine, or the shard engines of a group in one process read
// through peer pointers; the finished row is stored into every pool of the list (replicated quiz state of a group).
__device__ __forceinline__ double resume_ratio(const ResumeSource &S, int64_t q, int64_t a, int64_t j) {
  for (int r = 0; r < S.nShards; r++) {
    if (q < S.qFirst[r] || q >= S.qFirst[r] + S.qCount[r]) continue;
    if (j < S.tFirst[r] || j >= S.tFirst[r] + S.TpL[r]) continue;       // TpL: the shard's padded column count
    const int64_t ql = q - S.qFirst[r], jl = j - S.tFirst[r];
    return __ddiv_rn(S.sA[r][(ql * S.K + a) * S.TpL[r] + jl], S.mD[r][ql * S.TpL[r] + jl]);   // padding lanes: 0 / 1
  }
  return 0.0;
}
__global__ void __launch_bounds__(256) k_resume_quiz(ResumeSource S, PoolList pools, const double *__restrict__ vB,
                                                     const uint32_t *__restrict__ tgaps, int64_t T,
                                                     const int64_t *__restrict__ slots,
                                                     const int64_t *__restrict__ aqStart, const int64_t *__restrict__ aqQ,
                                                     const int64_t *__restrict__ aqA, int W, int *__restrict__ status) {
  extern __shared__ double sm[];
  __shared__ long long sMax[256];
  const QuizPool &qp = pools.p[0];
  const int64_t slot = slots[blockIdx.x];
  double *prior = qp.priors + slot * qp.Tp;
  double *lprior = qp.logPriors + slot * qp.Tp;
  long long *totExp = reinterpret_cast<long long *>(lprior);
  const int64_t Tp = qp.Tp, first = aqStart[blockIdx.x], limit = aqStart[blockIdx.x + 1];
  const unsigned long long EXPMASK = 0x7FF0000000000000ull, EXP0 = 0x3FF0000000000000ull;
  long long myMax = LLONG_MIN;
  for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) {
    double m = 0.0;
    long long e = 0;
    for (int64_t x = first; x < limit; x++) {
      const double P = resume_ratio(S, aqQ[x], aqA[x], j);
      const double old = (x == first) ? vB[j & 3] : m;
      const unsigned long long pb = (unsigned long long)__double_as_longlong(__dmul_rn(old, P));
      m = __longlong_as_double((long long)((pb & ~EXPMASK) | EXP0));                   // MakeExponent0
      const long long pe = (long long)((pb & EXPMASK) >> 52);                          // ExtractExponents64<false>
      e = (x == first) ? pe : e + pe;
    }
    const long long tot = e + 1023;   // + exponent field of the mantissa, which MakeExponent0 just set to 1023
    prior[j] = m;
    totExp[j] = tot;
    if (j < T && !bit32(tgaps, j)) myMax = max(myMax, tot);
  }
  sMax[threadIdx.x] = myMax;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sMax[threadIdx.x] = max(sMax[threadIdx.x], sMax[threadIdx.x + o]);
    __syncthreads();
  }
  const long long fullMax = sMax[0];
  int ceilLog2T = 0;
  while ((1ll << ceilLog2T) < T) ceilLog2T++;                                          // SRMath::CeilLog2
  const long long highBound = 1023 + 1023 - ceilLog2T - 2;                             // CpuEngine.cpp:314
  if (fullMax <= LLONG_MIN + highBound + 1) {                                          // :316-319
    if (threadIdx.x == 0) status[blockIdx.x] = 1;
    return;
  }
  const long long corr = highBound - fullMax;                                          // :320
  for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) {                             // CENormPriorsSubtaskCorrSum.cpp:24-41
    const unsigned long long mb = (unsigned long long)__double_as_longlong(prior[j]);
    const long long normExp = totExp[j] + corr;
    const bool zero = normExp < 1 || j >= T || bit32(tgaps, j);
    prior[j] = zero ? 0.0 : __longlong_as_double((long long)((mb & ~EXPMASK) | ((unsigned long long)normExp << 52)));
  }
  if (threadIdx.x == 0) {
    status[blockIdx.x] = 0;
    for (int r = 0; r < pools.n; r++) {
      const QuizPool &pr = pools.p[r];
      for (int64_t w = 0; w < pr.askedWords; w++) pr.asked[slot * pr.askedWords + w] = 0;
      for (int64_t x = first; x < limit; x++) pr.asked[slot * pr.askedWords + (aqQ[x] >> 6)] |= 1ull << (aqQ[x] & 63);
      pr.active[slot] = -1;
    }
  }
  __syncthreads();
  kahan_normalise(prior, lprior, T, Tp, W, sm);
  if (pools.n > 1) {         // the other shards' replicas of the quiz (peer memory)
    __syncthreads();
    for (int64_t j = threadIdx.x; j < Tp; j += blockDim.x) {
      const double v = prior[j], lv = lprior[j];
      for (int r = 1; r < pools.n; r++) { pools.p[r].priors[slot * Tp + j] = v; pools.p[r].logPriors[slot * Tp + j] = lv; }
    }
    __threadfence_system();
  }
}
void launch_resume_quiz_multi(const ResumeSource &src, const PoolList &pools, const double *dVB, const uint32_t *dTGaps, int64_t T,
                              int64_t n, const int64_t *dSlots, const int64_t *dAqStart, const int64_t *dAqQ, const int64_t *dAqA,
                              int W, int *dStatus, cudaStream_t st) {
  if (n <= 0) return;
  k_resume_quiz<<<(unsigned)n, 256, priors_smem(W), st>>>(src, pools, dVB, dTGaps, T, dSlots, dAqStart, dAqQ, dAqA, W, dStatus);
  count_launch();
}
void launch_resume_quiz(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, const int64_t *dAqStart,
                        const int64_t *dAqQ, const int64_t *dAqA, int W, int *dStatus, cudaStream_t st) {
  ResumeSource S;
  S.nShards = 1; S.K = kb.K;
  S.sA[0] = kb.sA; S.mD[0] = kb.mD; S.qFirst[0] = kb.qFirst; S.qCount[0] = kb.qCount; S.tFirst[0] = 0; S.TpL[0] = kb.Tp;
  PoolList pools;
  pools.n = 1; pools.p[0] = qp;
  launch_resume_quiz_multi(S, pools, kb.vB, kb.tgaps, kb.T, n, dSlots, dAqStart, dAqQ, dAqA, W, dStatus, st);
}

__global__ void k_refresh_log_priors(QuizPool qp, const int64_t *__restrict__ slots) {
  const int64_t slot = slots[blockIdx.x];
  for (int64_t j = threadIdx.x; j < qp.Tp; j += blockDim.x)
    qp.logPriors[slot * qp.Tp + j] = log2(qp.priors[slot * qp.Tp + j]);
}
void launch_refresh_log_priors(const QuizPool &qp, int64_t n, const int64_t *dSlots, cudaStream_t st) {
  if (n <= 0) return;
  k_refresh_log_priors<<<(unsigned)n, 256, 0, st>>>(qp, dSlots);
  count_launch();
}

// prior rows of a batch <-> a contiguous [n][Tp] staging buffer (all-reduce of updated priors in question-sharded mode)
__global__ void k_gather_prior_rows(QuizPool qp, const int64_t *__restrict__ slots, double *__restrict__ buf) {
  const int64_t slot = slots[blockIdx.x];
  for (int64_t j = threadIdx.x; j < qp.Tp; j += blockDim.x) buf[blockIdx.x * qp.Tp + j] = qp.priors[slot * qp.Tp + j];
}
__global__ void k_scatter_prior_rows(QuizPool qp, const int64_t *__restrict__ slots, const double *__restrict__ buf) {
  const int64_t slot = slots[blockIdx.x];
  for (int64_t j = threadIdx.x; j < qp.Tp; j += blockDim.x) {
    const double v = buf[blockIdx.x * qp.Tp + j];
    qp.priors[slot * qp.Tp + j] = v;
    qp.logPriors[slot * qp.Tp + j] = log2(v);
  }
}
void launch_gather_prior_rows(const QuizPool &qp, int64_t n, const int64_t *dSlots, double *dBuf, cudaStream_t st) {
  if (n <= 0) return;
  k_gather_prior_rows<<<(unsigned)n, 256, 0, st>>>(qp, dSlots, dBuf);
  count_launch();
}
void launch_scatter_prior_rows(const QuizPool &qp, int64_t n, const int64_t *dSlots, const double *dBuf, cudaStream_t st) {
  if (n <= 0) return;
  k_scatter_prior_rows<<<(unsigned)n, 256, 0, st>>>(qp, dSlots, dBuf);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// Exchange over peer memory (pqa_kernels.cuh). The stores of the kernels that ran before this one on the stream are
// complete at the kernel boundary; the system fence + release store order them before the flag for the peer GPUs.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void k_p2p_barrier(P2PFlags f, uint64_t epoch, uint64_t timeoutNs) {
  const int r = threadIdx.x;
  if (r >= f.nRanks) return;
  __threadfence_system();
  uint64_t *dst = f.flags[r] + f.rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(epoch) : "memory");
  const uint64_t *mine = f.flags[f.rank] + r;
  const uint64_t t0 = globaltimer_ns();
  for (;;) {
    uint64_t seen;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    if (seen >= epoch) break;
    if (globaltimer_ns() - t0 > timeoutNs) { *f.errFlag = epoch; break; }
    __nanosleep(200);
  }
  __threadfence_system();
}
void launch_p2p_barrier(const P2PFlags &f, uint64_t epoch, uint64_t timeoutNs, cudaStream_t st) {
  k_p2p_barrier<<<1, 32, 0, st>>>(f, epoch, timeoutNs);
  count_launch();
}

__global__ void k_p2p_wait(const uint64_t *flags, int nFlags, uint64_t value, uint64_t *errFlag, uint64_t timeoutNs) {
  const uint64_t t0 = globaltimer_ns();
  for (int f = threadIdx.x; f < nFlags; f += blockDim.x) {
    for (;;) {
      uint64_t seen;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + f) : "memory");
      if (seen >= value) break;
      if (globaltimer_ns() - t0 > timeoutNs) { *errFlag = value; break; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}
void launch_p2p_wait(const uint64_t *flags, int nFlags, uint64_t value, uint64_t *errFlag, uint64_t timeoutNs, cudaStream_t st) {
  k_p2p_wait<<<1, 32, 0, st>>>(flags, nFlags, value, errFlag, timeoutNs);
  count_launch();
}
__global__ void k_p2p_signal(PeerBufs targets, uint64_t value) {
  const int r = threadIdx.x;
  if (r >= targets.n) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(targets.p[r]), "l"(value) : "memory");
}
void launch_p2p_signal(const PeerBufs &targets, uint64_t value, cudaStream_t st) {
  k_p2p_signal<<<1, 32, 0, st>>>(targets, value);
  count_launch();
}

__global__ void k_p2p_push_prior_rows(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                      const int64_t *__restrict__ questions, PeerBufs out) {
  const int64_t b = blockIdx.y, q = questions[b];
  if (q < kb.qFirst || q >= kb.qFirst + kb.qCount) return;        // another shard answers for this quiz
  const double *prior = qp.priors + slots[b] * qp.Tp;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < qp.Tp; j += (int64_t)gridDim.x * blockDim.x) {
    const double v = prior[j];
    for (int r = 0; r < out.n; r++) out.p[r][b * qp.Tp + j] = v;
  }
}
void launch_p2p_push_prior_rows(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                const int64_t *dQuestions, const PeerBufs &out, cudaStream_t st) {
  if (n <= 0 || out.n <= 0) return;
  int64_t gx = (qp.Tp + 255) / 256;
  if (gx > 64) gx = 64;
  k_p2p_push_prior_rows<<<dim3((unsigned)gx, (unsigned)n), 256, 0, st>>>(kb, qp, dSlots, dQuestions, out);
  count_launch();
}
__global__ void k_p2p_pull_prior_rows(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                      const int64_t *__restrict__ questions, const double *__restrict__ rows) {
  const int64_t b = blockIdx.y, q = questions[b];
  if (q >= kb.qFirst && q < kb.qFirst + kb.qCount) return;        // own row: already in place
  double *prior = qp.priors + slots[b] * qp.Tp, *lprior = qp.logPriors + slots[b] * qp.Tp;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < qp.Tp; j += (int64_t)gridDim.x * blockDim.x) {
    const double v = rows[b * qp.Tp + j];
    prior[j] = v;
    lprior[j] = log2(v);
  }
}
void launch_p2p_pull_prior_rows(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                                const int64_t *dQuestions, const double *dRows, cudaStream_t st) {
  if (n <= 0) return;
  int64_t gx = (qp.Tp + 255) / 256;
  if (gx > 64) gx = 64;
  k_p2p_pull_prior_rows<<<dim3((unsigned)gx, (unsigned)n), 256, 0, st>>>(kb, qp, dSlots, dQuestions, dRows);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// Maintenance helpers (pqa_kernels.cuh)
__global__ void k_copy_rows(double *__restrict__ dst, int64_t dstStride, const double *__restrict__ src, int64_t srcStride,
                            int64_t nRows, int64_t nCols) {
  const int64_t n = nRows * nCols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
    const int64_t r = x / nCols, c = x - r * nCols;
    dst[r * dstStride + c] = src[r * srcStride + c];
  }
}
void launch_copy_rows(double *dst, int64_t dstStride, const double *src, int64_t srcStride, int64_t nRows, int64_t nCols,
                      cudaStream_t st) {
  if (nRows <= 0 || nCols <= 0) return;
  k_copy_rows<<<grid_for(nRows * nCols, 256), 256, 0, st>>>(dst, dstStride, src, srcStride, nRows, nCols);
  count_launch();
}
__global__ void k_fill_rects(const FillRect *__restrict__ rects) {
  const FillRect R = rects[blockIdx.y];
  const int64_t n = R.nRows * R.nCols, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
    const int64_t r = x / R.nCols, c = x - r * R.nCols;
    R.base[(R.row0 + r) * R.stride + R.col0 + c] = R.value;
  }
}
void launch_fill_rects(const FillRect *dRects, int64_t nRects, cudaStream_t st) {
  for (int64_t r0 = 0; r0 < nRects; r0 += 32768) {     // grid.y limit
    const int64_t nr = nRects - r0 < 32768 ? nRects - r0 : 32768;
    k_fill_rects<<<dim3(64, (unsigned)nr), 256, 0, st>>>(dRects + r0);
    count_launch();
  }
}
__global__ void k_gather_kb(double *__restrict__ dst, int64_t dstStride, const double *__restrict__ src, int64_t srcStride,
                            const int64_t *__restrict__ oldRow, int64_t nNewRows, int64_t rowsPer,
                            const int64_t *__restrict__ oldCol, int64_t nNewCols, double padValue) {
  const int64_t n = nNewRows * rowsPer * dstStride, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
    const int64_t row = x / dstStride, j = x - row * dstStride;
    const int64_t i = row / rowsPer, k = row - i * rowsPer;
    dst[x] = j < nNewCols ? src[(oldRow[i] * rowsPer + k) * srcStride + oldCol[j]] : padValue;
  }
}
void launch_gather_kb(double *dst, int64_t dstStride, const double *src, int64_t srcStride, const int64_t *dOldRow,
                      int64_t nNewRows, int64_t rowsPer, const int64_t *dOldCol, int64_t nNewCols, double padValue,
                      cudaStream_t st) {
  if (nNewRows <= 0) return;
  k_gather_kb<<<grid_for(nNewRows * rowsPer * dstStride, 256), 256, 0, st>>>(dst, dstStride, src, srcStride, dOldRow, nNewRows,
                                                                            rowsPer, dOldCol, nNewCols, padValue);
  count_launch();
}

__global__ void k_set_active(QuizPool qp, int64_t n, const int64_t *__restrict__ slots,
                             const int64_t *__restrict__ questions) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n) qp.active[slots[x]] = questions[x];
}
void launch_set_active(const QuizPool &qp, int64_t n, const int64_t *dSlots, const int64_t *dQuestions,
                       cudaStream_t st) {
  if (n <= 0) return;
  k_set_active<<<grid_for(n, 128), 128, 0, st>>>(qp, n, dSlots, dQuestions);
  count_launch();
}

__global__ void k_flush_l2(double *buf, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) buf[x] = (double)x;
}
void launch_flush_l2(void *buf, size_t bytes, cudaStream_t st) {
  k_flush_l2<<<148 * 8, 256, 0, st>>>((double *)buf, bytes / 8);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// Exact question evaluation: CEEvalQsSubtaskConsider<SRDoubleNumber>::Run (CEEvalQsSubtaskConsider.cpp:41-217)
// with every rounding in the reference's order. Four threads play the four AVX lanes of one (quiz, question)
// evaluation: thread l owns targets j with j % 4 == l and runs the lane's Kahan sums sequentially in vector
// order; PreciseSum/PairSum gather the lanes. A warp holds 8 quizzes of the same question, so its sA/mD loads are
// warp-broadcasts. lik and invD are recomputed in pass 2 instead of being spilled (deterministic => same bits).
__global__ void __launch_bounds__(128) k_eval_exact(DeviceKB kb, QuizPool qp, int64_t n,
                                                    const int64_t *__restrict__ slots,
                                                    double *__restrict__ priority, EvalDetail det) {
  const int64_t iLocal = blockIdx.x, i = kb.qFirst + iLocal;
  const int lane = threadIdx.x & 3;
  const int64_t b = (int64_t)blockIdx.y * 32 + (threadIdx.x >> 2);
  if (b >= n) return;  // whole 4-thread group leaves together
  const int64_t slot = slots[b];
  if (bit32(kb.qgaps, i) || bit64(qp.asked + slot * qp.askedWords, i)) {  // :54-58
    if (lane == 0) priority[b * kb.Q + i] = nan("");
    return;
  }
  const int64_t T = kb.T, Tp = kb.Tp, K = kb.K, nTV = Tp >> 2;
  const double *__restrict__ prior = qp.priors + slot * Tp;
  const double *__restrict__ mDi = kb.mD + iLocal * Tp;
  const double *__restrict__ tbl = kb.log2tbl;

  Kahan totW; totW.init(0.0);        // scalar accumulator, every lane carries an identical copy (:89,:134)
  Kahan accL; accL.init();           // one lane of accLack, never reset across answers (:61,:117)
  Kahan avgH, avgV; avgH.init(); avgV.init();  // this thread's lane of accAvgH / accAvgV (:139-172)
  const int64_t nVectorized = (K >> 2) << 2;
  for (int64_t k = 0; k < K; k++) {
    const double *__restrict__ sAik = kb.sA + (iLocal * K + k) * Tp;
    Kahan accW; accW.init();
    for (int64_t v = 0; v < nTV; v++) {                          // pass 1 (:66-87)
      const int64_t j = 4 * v + lane;
      const bool gap = j >= T || bit32(kb.tgaps, j);
      const double invD = gap ? 0.0 : __ddiv_rn(1.0, mDi[j]);
      const double P = gap ? 0.0 : __dmul_rn(sAik[j], invD);
      const double lik = gap ? 0.0 : __dmul_rn(P, prior[j]);
      accW.add(lik);
    }
    const double Wk = group_precise_sum(accW);                   // :88
    totW.add(Wk);
    const double invW = __ddiv_rn(1.0, Wk);                      // :91
    Kahan accH, accV; accH.init(); accV.init();
    for (int64_t v = 0; v < nTV; v++) {                          // pass 2 (:95-128)
      const int64_t j = 4 * v + lane;
      const bool gap = j >= T || bit32(kb.tgaps, j);
      const double invD = gap ? 0.0 : __ddiv_rn(1.0, mDi[j]);
      const double P = gap ? 0.0 : __dmul_rn(sAik[j], invD);
      const double pr = gap ? 0.0 : prior[j];
      const double lik = gap ? 0.0 : __dmul_rn(P, pr);
      const double post = __dmul_rn(lik, invW);                  // :97
      const double l2 = gap ? 0.0 : log2hot(post, tbl);          // :106
      accH.add(__dmul_rn(post, l2));                             // :113-114
      accL.add(gap ? 0.0 : __ddiv_rn(__dmul_rn(invD, invD), l2)); // :116-117
      const double d = __dsub_rn(post, pr);                      // :119
      accV.add(__dmul_rn(d, d));                                 // :126-127
    }
    const double Hk = -group_precise_sum(accH);                  // PairSum (:129-132)
    const double Vk = group_precise_sum(accV);
    // weighted averages: answers 0..nVectorized-1 go to lane k%4, the tail to lane k-nVectorized (:148-172)
    const int tgtLane = (k < nVectorized) ? (int)(k & 3) : (int)(k - nVectorized);
    if (lane == tgtLane) {
      avgH.add(__dmul_rn(Wk, Hk));
      avgV.add(__dmul_rn(Wk, sqrt(Vk)));
    }
    if (lane == 0) {
      if (det.W) det.W[(b * kb.Q + i) * K + k] = Wk;
      if (det.H) det.H[(b * kb.Q + i) * K + k] = Hk;
      if (det.V) det.V[(b * kb.Q + i) * K + k] = Vk;
    }
  }
  const double sumH = group_precise_sum(avgH), sumV = group_precise_sum(avgV);  // :175
  const double lackSum = group_precise_sum(accL);
  if (lane == 0) {
    const double tw = totW.get();
    const double aH = __ddiv_rn(sumH, tw), aV = __ddiv_rn(sumV, tw);            // :176-177
    const double nExp = exp2(aH);                                               // :181
    const double cLnMaxV = 0.34657359027997265470861606072909;                  // SRMath::_cLnSqrt2
    const double lnV = (aV == 0) ? -746.0 : log(aV);                            // :27-29
    const double n1 = (double)(kb.nValidTargets + 1);
    const double vComp = __ddiv_rn(1.0, __dadd_rn(__dsub_rn(cLnMaxV, lnV), __ddiv_rn(cLnMaxV, __dmul_rn(n1, n1))));
    const double lack = -lackSum;                                               // :201
    priority[b * kb.Q + i] = __dmul_rn(__dmul_rn(lack, pow(vComp, 9.0)), pow(nExp, -2.0));  // :207
    if (det.lack) det.lack[b * kb.Q + i] = lack;
  }
}

void launch_eval_staged(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, double *dPriority,
                        const EvalDetail &det, const EvalConfig &cfg, cudaStream_t st);

void launch_eval_questions(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                           double *dPriority, const EvalDetail &det, const EvalConfig &cfg, cudaStream_t st) {
  if (n <= 0) return;
  if (cfg.which == 1) {
    dim3 grid((unsigned)kb.qCount, (unsigned)((n + 31) / 32));
    k_eval_exact<<<grid, 128, 0, st>>>(kb, qp, n, dSlots, dPriority, det);
    count_launch();
  } else {
    launch_eval_staged(kb, qp, n, dSlots, dPriority, det, cfg, st);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Question selection: CpuEngine::NextQuestionSpec (CpuEngine.cpp:337-415) after the evaluation. One CTA per quiz.
//   chunks = CalcSplit(Q, 8W); per chunk a scalar-Kahan running sum over its questions, asked/gap questions repeat
//   the running value (CEEvalQsSubtaskConsider.cpp:54-58,212-214); grand totals = scalar Kahan over the chunks' last
//   values (CpuEngine.cpp:362-374); r = totG * u64 / (2^64-1) (SRDoubleNumber.h:35-39); upper_bound over chunks,
//   upper_bound inside the chunk (:380-400); asked/gap -> nearest available question (BaseEngine.cpp:60-124).
int64_t select_chunk_count(int64_t Q, int W) { return split_count(Q, (int64_t)W * 8); }

// Shared memory: nChunks grand totals. runLength[n*Q] is required (the in-chunk binary search reads it back).
__global__ void __launch_bounds__(256) k_select_question(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                                         const double *__restrict__ priority,
                                                         const uint64_t *__restrict__ randoms, int W,
                                                         double *__restrict__ runLength, double *__restrict__ grandOut,
                                                         int64_t *__restrict__ questions, int setActive) {
  extern __shared__ double sGrand[];
  const int64_t b = blockIdx.x, Q = kb.Q;
  const int64_t nChunks = split_count(Q, (int64_t)W * 8);
  const int64_t chosen = select_question_cta(kb, qp, slots[b], priority + b * Q, questions ? randoms[b] : 0ull, W,
                                             runLength + b * Q, grandOut ? grandOut + b * nChunks : nullptr, questions != nullptr,
                                             setActive, sGrand);
  if (threadIdx.x == 0 && questions) questions[b] = chosen;
}

void launch_select_question(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots,
                            const double *dPriority, const uint64_t *dRandoms, int W, double *dRunLength,
                            double *dGrand, int64_t *dQuestions, int setActive, cudaStream_t st) {
  if (n <= 0) return;
  const int64_t nChunks = select_chunk_count(kb.Q, W);
  int block = 32;
  while (block < nChunks && block < 256) block <<= 1;
  k_select_question<<<(unsigned)n, block, sizeof(double) * (size_t)nChunks, st>>>(
      kb, qp, dSlots, dPriority, dRandoms, W, dRunLength, dGrand, dQuestions, setActive);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// ListTopTargets: CEListTopTargetsAlgorithm::RunHeapifyBased (CEListTopTargetsAlgorithm.cpp:30-97) with the piece
// heaps of CEHeapifyPriorsSubtaskMake (CEHeapifyPriorsSubtaskMake.cpp:42-87). Ties between equal probabilities are
// resolved by the mechanics of libstdc++'s make_heap / pop_heap and SRHeapHelper::Down (SRHeap.h:12-37), so those
// are restated here operation by operation; one thread builds one piece heap, thread 0 runs the merge.
struct Rated { int64_t iTarget; double prob; };      // RatedTarget, Interface/PqaCommon.h:54-61
struct HeadItem { double prob; int64_t iSource; };   // RatingsHeapItem, RatingsHeap.h:11-20

template <typename TItem>
__device__ void heap_push(TItem *first, int64_t hole, int64_t top, TItem value) {  // std::__push_heap
  int64_t parent = (hole - 1) / 2;
  while (hole > top && first[parent].prob < value.prob) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}
template <typename TItem>
__device__ void heap_adjust(TItem *first, int64_t hole, int64_t len, TItem value) {  // std::__adjust_heap
  const int64_t top = hole;
  int64_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (first[child].prob < first[child - 1].prob) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  heap_push(first, hole, top, value);
}
template <typename TItem>
__device__ void heap_make(TItem *first, int64_t len) {  // std::make_heap
  if (len < 2) return;
  for (int64_t parent = (len - 2) / 2;; parent--) {
    const TItem value = first[parent];
    heap_adjust(first, parent, len, value);
    if (parent == 0) return;
  }
}
template <typename TItem>
__device__ void heap_pop(TItem *first, int64_t len) {  // std::pop_heap
  if (len > 1) {
    const TItem value = first[len - 1];
    first[len - 1] = first[0];
    heap_adjust(first, 0, len - 1, value);
  }
}
__device__ void head_down(HeadItem *first, int64_t len) {  // SRHeapHelper::Down, SRHeap.h:16-37
  int64_t cur = 0;
  for (;;) {
    const int64_t c1 = 2 * cur + 1;
    if (c1 >= len) return;
    const int64_t c2 = c1 + 1;
    if (c2 >= len) {
      if (first[cur].prob < first[c1].prob) { const HeadItem t = first[cur]; first[cur] = first[c1]; first[c1] = t; }
      return;
    }
    const int64_t hi = (first[c2].prob < first[c1].prob) ? c1 : c2;
    if (!(first[cur].prob < first[hi].prob)) return;
    const HeadItem t = first[cur]; first[cur] = first[hi]; first[hi] = t;
    cur = hi;
  }
}

// Shared memory: W head items, W starts, W limits, then (if useSmem) T Rated items.
// The piece heaps are built by the whole CTA. Compaction of a piece (CEHeapifyPriorsSubtaskMake.cpp:42-53) is a stable
// filter: one warp per piece, ballot prefix. std::make_heap (:85-86) is a sequence of sift-downs from the last parent to
// the root; the sift-downs of nodes on one tree level touch disjoint subtrees and every deeper level comes first in that
// sequence, so running level by level -- all nodes of a level (of all pieces) in parallel, a barrier between levels --
// leaves exactly the heap the sequential loop leaves, ties included. The k pops of the merge stay sequential (thread 0).
__global__ void __launch_bounds__(256) k_list_top_targets(DeviceKB kb, QuizPool qp, const int64_t *__restrict__ slots,
                                                          int W, int64_t maxCount, Rated *__restrict__ scratch,
                                                          int useSmem, Rated *__restrict__ dest,
                                                          int64_t *__restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smRaw[];
  HeadItem *head = (HeadItem *)smRaw;
  int64_t *starts = (int64_t *)(head + W);
  int64_t *limits = starts + W;
  Rated *ratings = useSmem ? (Rated *)(limits + W) : scratch + (int64_t)blockIdx.x * kb.T;
  __shared__ int sMaxLen;
  const int64_t b = blockIdx.x, slot = slots[b], T = kb.T;
  const double *prior = qp.priors + slot * qp.Tp;
  const int64_t nPieces = split_count(T, W);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nWarps = blockDim.x >> 5;
  if (threadIdx.x == 0) sMaxLen = 0;
  __syncthreads();
  for (int64_t p = warp; p < nPieces; p += nWarps) {      // CEHeapifyPriorsSubtaskMake.cpp:42-53
    const int64_t first = split_start(T, W, p), limit = split_start(T, W, p + 1);
    int64_t sel = first;
    for (int64_t base = first; base < limit; base += 32) {
      const int64_t j = base + lane;
      double prob = 0.0;
      bool keep = false;
      if (j < limit && !bit32(kb.tgaps, j)) { prob = prior[j]; keep = prob > 0; }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int64_t at = sel + __popc(mask & ((1u << lane) - 1u));
        ratings[at].prob = prob; ratings[at].iTarget = j;
      }
      sel += __popc(mask);
    }
    if (lane == 0) { starts[p] = first; limits[p] = sel; atomicMax(&sMaxLen, (int)(sel - first)); }
  }
  __syncthreads();
  // std::make_heap of every piece (:85-86), level by level from the deepest parents up
  const int maxLen = sMaxLen;
  if (maxLen >= 2) {
    int depth = 0;
    while ((2ll << depth) - 1 <= (maxLen - 2) / 2) depth++;          // deepest level that holds a parent
    for (int d = depth; d >= 0; d--) {
      const int64_t perPiece = 1ll << d, total = nPieces * perPiece;
      for (int64_t x = threadIdx.x; x < total; x += blockDim.x) {
        const int64_t p = x >> d, node = perPiece - 1 + (x & (perPiece - 1));
        const int64_t len = limits[p] - starts[p];
        if (len < 2 || node > (len - 2) / 2) continue;
        Rated *first = ratings + starts[p];
        const Rated value = first[node];
        heap_adjust(first, node, len, value);
      }
      __syncthreads();
    }
  }
  if (threadIdx.x != 0) return;
  int64_t nHh = 0;
  for (int64_t p = 0; p < nPieces; p++) {                // CEListTopTargetsAlgorithm.cpp:57-66
    if (limits[p] == starts[p]) continue;
    head[nHh].iSource = p; head[nHh].prob = ratings[starts[p]].prob; nHh++;
  }
  heap_make(head, nHh);
  int64_t listed = maxCount;
  Rated *out = dest + b * maxCount;
  for (int64_t i = 0; i < maxCount; i++) {               // :68-94
    if (nHh == 0) { listed = i; break; }
    const int64_t piece = head[0].iSource, start = starts[piece], lim = limits[piece];
    out[i].prob = head[0].prob;
    out[i].iTarget = ratings[start].iTarget;
    if (start + 1 == lim) { heap_pop(head, nHh); nHh--; continue; }
    heap_pop(ratings + start, lim - start);
    limits[piece]--;
    head[0].prob = ratings[start].prob;
    head_down(head, nHh);
  }
  counts[b] = listed;
}

void launch_list_top_targets(const DeviceKB &kb, const QuizPool &qp, int64_t n, const int64_t *dSlots, int W,
                             int64_t maxCount, void *dScratch, void *dDest, int64_t *dCounts, cudaStream_t st) {
  if (n <= 0) return;
  const size_t fixed = (size_t)W * (sizeof(HeadItem) + 2 * sizeof(int64_t));
  const size_t items = (size_t)kb.T * sizeof(Rated);
  const int useSmem = fixed + items <= 200 * 1024;
  const size_t smem = fixed + (useSmem ? items : 0);
  // function attributes are per device: several engines of one process may sit on different GPUs (ShardGroup)
  static std::atomic<unsigned long long> attrDevices{0};
  int attrDev = 0;
  cudaGetDevice(&attrDev);
  if (!((attrDevices.load(std::memory_order_relaxed) >> (attrDev & 63)) & 1ull)) {
    cudaFuncSetAttribute(k_list_top_targets, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attrDevices.fetch_or(1ull << (attrDev & 63), std::memory_order_relaxed);
  }
  const int block = kb.T >= 4096 ? 256 : 128;     // the whole CTA builds the piece heaps level by level
  k_list_top_targets<<<(unsigned)n, block, smem, st>>>(kb, qp, dSlots, W, maxCount, (Rated *)dScratch, useSmem,
                                                       (Rated *)dDest, dCounts);
  count_launch();
}

// ---------------------------------------------------------------------------------------------------------
// RecordQuizTarget / Train: CETrainOperation::{ProcessOne,Perform1,Perform2} (CETrainOperation.cpp:15-83) with the
// increments of CETrainTaskNumSpec.h:24-32. One thread applies one (question, target) cell group's operations in
// sequence order, so any interleaving of quizzes that hits the same cell reproduces the sequential result.
__global__ void k_train_ops(DeviceKB kb, const TrainOp *__restrict__ ops, const int64_t *__restrict__ groupStart,
                            int64_t nGroups) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGroups) return;
  for (int64_t o = groupStart[g]; o < groupStart[g + 1]; o++) {
    const TrainOp op = ops[o];
    const double b = op.amount;
    const int64_t ql = op.q - kb.qFirst;   // the host only sends operations on questions this device owns
    double *cellD = kb.mD + ql * kb.Tp + op.target;
    double *cellA0 = kb.sA + (ql * kb.K + op.a0) * kb.Tp + op.target;
    if (op.a1 < 0 || op.a1 == op.a0) {
      // ProcessOne (:15-26) with (2b, b^2), or the doubled step (:34-36) with (4b, 4b^2)
      const double twoB = (op.a1 < 0) ? __dmul_rn(2.0, b) : __dmul_rn(4.0, b);
      const double bSq = (op.a1 < 0) ? __dmul_rn(b, b) : __dmul_rn(4.0, __dmul_rn(b, b));
      const double aSq = *cellA0;
      const double addend = __dadd_rn(__dmul_rn(sqrt(aSq), twoB), bSq);
      *cellA0 = __dadd_rn(aSq, addend);
      *cellD = __dadd_rn(*cellD, addend);
    } else {
      // same question, two different answers (:38-47): D receives addend0 twice, as in the reference
      double *cellA1 = kb.sA + (ql * kb.K + op.a1) * kb.Tp + op.target;
      const double twoB = __dmul_rn(2.0, b), bSq = __dmul_rn(b, b);
      const double a0Sq = *cellA0, a1Sq = *cellA1;
      const double add0 = __dadd_rn(__dmul_rn(sqrt(a0Sq), twoB), bSq);
      const double add1 = __dadd_rn(__dmul_rn(sqrt(a1Sq), twoB), bSq);
      *cellA0 = __dadd_rn(a0Sq, add0);
      *cellA1 = __dadd_rn(a1Sq, add1);
      *cellD = __dadd_rn(*cellD, __dadd_rn(add0, add0));
    }
  }
}
void launch_train_ops(const DeviceKB &kb, const TrainOp *dOps, const int64_t *dGroupStart, int64_t nGroups,
                      cudaStream_t st) {
  if (nGroups <= 0) return;
  k_train_ops<<<(unsigned)((nGroups + 127) / 128), 128, 0, st>>>(kb, dOps, dGroupStart, nGroups);
  count_launch();
}

__global__ void k_add_vb(DeviceKB kb, const int64_t *__restrict__ targets, const double *__restrict__ amounts,
                         const int64_t *__restrict__ groupStart, int64_t nGroups) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nGroups) return;
  const int64_t first = groupStart[g], limit = groupStart[g + 1];
  double v = kb.vB[targets[first]];
  for (int64_t o = first; o < limit; o++) v = __dadd_rn(v, amounts[o]);   // CpuEngine.cpp:462
  kb.vB[targets[first]] = v;
}
void launch_add_vb(const DeviceKB &kb, const int64_t *dTargets, const double *dAmounts, const int64_t *dGroupStart,
                   int64_t nGroups, cudaStream_t st) {
  if (nGroups <= 0) return;
  k_add_vb<<<(unsigned)((nGroups + 127) / 128), 128, 0, st>>>(kb, dTargets, dAmounts, dGroupStart, nGroups);
  count_launch();
}

// CUDA loads a kernel's code lazily at its first launch, and that load can wait for running kernels. A barrier kernel
// that is spinning for another engine of the same process must never be what such a load waits for, so the engines of a
// peer-memory exchange load every kernel of the exchanged call sequences up front (cudaFuncGetAttributes loads).
void preload_staged_kernels(int K);
void preload_exchange_kernels(int K) {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, k_update_priors<0>);
  cudaFuncGetAttributes(&a, k_update_priors<1>);
  cudaFuncGetAttributes(&a, k_tshard_ra_partial);
  cudaFuncGetAttributes(&a, k_tshard_ra_finish);
  cudaFuncGetAttributes(&a, k_update_lanes<0>);
  cudaFuncGetAttributes(&a, k_update_lanes<1>);
  cudaFuncGetAttributes(&a, k_update_lanes<2>);
  cudaFuncGetAttributes(&a, k_normalise_rows);
  cudaFuncGetAttributes(&a, k_p2p_barrier);
  cudaFuncGetAttributes(&a, k_p2p_wait);
  cudaFuncGetAttributes(&a, k_p2p_signal);
  cudaFuncGetAttributes(&a, k_p2p_push_prior_rows);
  cudaFuncGetAttributes(&a, k_p2p_pull_prior_rows);
  cudaFuncGetAttributes(&a, k_select_question);
  cudaFuncGetAttributes(&a, k_set_active);
  cudaFuncGetAttributes(&a, k_list_top_targets);
  cudaFuncGetAttributes(&a, k_gather_prior_rows);
  cudaFuncGetAttributes(&a, k_scatter_prior_rows);
  preload_staged_kernels(K);
  (void)cudaGetLastError();
}

} // namespace pqa
