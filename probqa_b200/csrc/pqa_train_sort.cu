// probqa_b200: large RecordQuizTarget / Train batches (BASELINE config 5). The reference applies a quiz' answers in
// sequence (CETrainOperation.cpp:15-83); operations on different (question, target) cells commute, operations on the same
// cell must keep their sequence order. For big batches the grouping by cell is done on the device: a stable radix sort of
// (cell key, sequence number) pairs (radix_sort_pairs below), then one thread per sorted position -- the thread that
// sits at the head of a run of equal keys applies the whole run in sequence order with the same arithmetic as k_train_ops.
// Results are bit-identical to the host-grouped path (tests/test_gpu_parity.py::test_record_quiz_target_and_train_bit_exact
// runs both).
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

#include <stdexcept>

namespace pqa {
void count_launch();

// ---------------------------------------------------------------------------------------------------------
// Stable least-significant-digit radix sort of (key, value) pairs, 8 bits per pass. A block owns a tile of kSortTile
// consecutive elements. Per pass: k_rs_hist counts the tile's digits into hist[digit][block]; k_rs_scan turns the table
// (read digit-major) into exclusive prefix sums = the first output position of every (digit, block); k_rs_scatter walks
// its tile in order, 256 elements at a time, and places every element at (position of its digit for this block) + (number
// of elements with the same digit earlier in the tile): earlier chunks are in the running offsets, earlier warps of the
// chunk in the per-warp counts, earlier lanes of the warp come from __match_any_sync. Equal keys therefore keep their order.
constexpr int kSortTile = 4096, kSortThreads = 256, kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads) k_rs_hist(const int64_t *__restrict__ keys, int64_t n, int shift,
                                                          unsigned *__restrict__ hist, int nBlocks) {
  __shared__ unsigned sHist[256];
  sHist[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t first = (int64_t)blockIdx.x * kSortTile;
  const int64_t limit = (first + kSortTile < n) ? first + kSortTile : n;
  for (int64_t x = first + threadIdx.x; x < limit; x += kSortThreads)
    atomicAdd(&sHist[(unsigned)((uint64_t)keys[x] >> shift) & 255u], 1u);
  __syncthreads();
  hist[(size_t)threadIdx.x * nBlocks + blockIdx.x] = sHist[threadIdx.x];
}

// one block: exclusive prefix sums over m entries in place
__global__ void __launch_bounds__(1024) k_rs_scan(unsigned *__restrict__ a, int64_t m) {
  __shared__ unsigned sPart[1024];
  const int64_t per = (m + 1023) / 1024;
  const int64_t first = (int64_t)threadIdx.x * per;
  const int64_t limit = (first + per < m) ? first + per : m;
  unsigned sum = 0u;
  for (int64_t x = first; x < limit; x++) sum += a[x];
  sPart[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {          // inclusive scan of the 1024 partial sums
    const unsigned v = (threadIdx.x >= (unsigned)d) ? sPart[threadIdx.x - d] : 0u;
    __syncthreads();
    sPart[threadIdx.x] += v;
    __syncthreads();
  }
  unsigned run = sPart[threadIdx.x] - sum;      // exclusive
  for (int64_t x = first; x < limit; x++) { const unsigned v = a[x]; a[x] = run; run += v; }
}

__global__ void __launch_bounds__(kSortThreads) k_rs_scatter(const int64_t *__restrict__ keysIn, const int64_t *__restrict__ valsIn,
                                                             int64_t *__restrict__ keysOut, int64_t *__restrict__ valsOut,
                                                             int64_t n, int shift, const unsigned *__restrict__ offsets,
                                                             int nBlocks) {
  __shared__ unsigned sOff[256];                     // next output position of every digit for this block
  __shared__ unsigned sWarp[kSortWarps][256];        // per chunk: how many elements with the digit every warp holds
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  sOff[threadIdx.x] = offsets[(size_t)threadIdx.x * nBlocks + blockIdx.x];
  const int64_t first = (int64_t)blockIdx.x * kSortTile;
  const int64_t limit = (first + kSortTile < n) ? first + kSortTile : n;
  for (int64_t c0 = first; c0 < limit; c0 += kSortThreads) {
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) sWarp[w][threadIdx.x] = 0u;
    __syncthreads();                                 // also orders the previous chunk's update of sOff before its use below
    const int64_t x = c0 + threadIdx.x;
    const bool valid = x < limit;
    int64_t key = 0, val = 0;
    if (valid) { key = keysIn[x]; val = valsIn[x]; }
    const unsigned digit = valid ? ((unsigned)((uint64_t)key >> shift) & 255u) : 256u + (unsigned)lane;   // invalid lanes match nobody
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
    const unsigned rankInWarp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rankInWarp == 0) sWarp[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      unsigned before = 0u;
      for (int w = 0; w < warp; w++) before += sWarp[w][digit];
      const unsigned pos = sOff[digit] + before + rankInWarp;
      keysOut[pos] = key;
      valsOut[pos] = val;
    }
    __syncthreads();
    unsigned total = 0u;
#pragma unroll
    for (int w = 0; w < kSortWarps; w++) total += sWarp[w][threadIdx.x];
    sOff[threadIdx.x] += total;
  }
}

static int64_t sort_blocks(int64_t n) { return (n + kSortTile - 1) / kSortTile; }
static size_t sort_hist_bytes(int64_t n) { return sizeof(unsigned) * 256 * (size_t)sort_blocks(n); }

// Sorts n pairs by the low `bits` bits of the keys. srcKeys / srcVals are only read; (aKeys, aVals) and (bKeys, bVals) are
// work buffers distinct from the source and from each other; the sorted pairs end in (bKeys, bVals).
static void radix_sort_pairs(const int64_t *srcKeys, const int64_t *srcVals, int64_t *aKeys, int64_t *aVals, int64_t *bKeys,
                             int64_t *bVals, unsigned *hist, int64_t n, int bits, cudaStream_t st) {
  const int passes = (bits + 7) / 8;
  const int nBlocks = (int)sort_blocks(n);
  const int64_t *inK = srcKeys, *inV = srcVals;
  for (int p = 0; p < passes; p++) {
    const bool toB = ((passes - 1 - p) & 1) == 0;      // the last pass writes b, the one before it a, ...
    int64_t *outK = toB ? bKeys : aKeys, *outV = toB ? bVals : aVals;
    k_rs_hist<<<nBlocks, kSortThreads, 0, st>>>(inK, n, 8 * p, hist, nBlocks);
    k_rs_scan<<<1, 1024, 0, st>>>(hist, (int64_t)256 * nBlocks);
    k_rs_scatter<<<nBlocks, kSortThreads, 0, st>>>(inK, inV, outK, outV, n, 8 * p, hist, nBlocks);
    count_launch(); count_launch(); count_launch();
    inK = outK; inV = outV;
  }
}

__global__ void k_train_keys(const TrainOp *__restrict__ ops, int64_t nOps, int64_t T, int64_t *__restrict__ keys,
                             int64_t *__restrict__ seq) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nOps) return;
  keys[x] = ops[x].q * T + ops[x].target;
  seq[x] = x;
}

__device__ __forceinline__ void apply_train_op(const DeviceKB &kb, const TrainOp &op) {
  const double b = op.amount;
  const int64_t ql = op.q - kb.qFirst;
  double *cellD = kb.mD + ql * kb.Tp + op.target;
  double *cellA0 = kb.sA + (ql * kb.K + op.a0) * kb.Tp + op.target;
  if (op.a1 < 0 || op.a1 == op.a0) {   // ProcessOne (CETrainOperation.cpp:15-26) or the doubled step (:34-36)
    const double twoB = (op.a1 < 0) ? __dmul_rn(2.0, b) : __dmul_rn(4.0, b);
    const double bSq = (op.a1 < 0) ? __dmul_rn(b, b) : __dmul_rn(4.0, __dmul_rn(b, b));
    const double aSq = *cellA0;
    const double addend = __dadd_rn(__dmul_rn(sqrt(aSq), twoB), bSq);
    *cellA0 = __dadd_rn(aSq, addend);
    *cellD = __dadd_rn(*cellD, addend);
  } else {                             // same question, two different answers (:38-47)
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

__global__ void k_train_ops_sorted(DeviceKB kb, const TrainOp *__restrict__ ops, const int64_t *__restrict__ keys,
                                   const int64_t *__restrict__ seq, int64_t nOps) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nOps) return;
  const int64_t key = keys[x];
  if (x > 0 && keys[x - 1] == key) return;            // not the head of its run
  for (int64_t y = x; y < nOps && keys[y] == key; y++) apply_train_op(kb, ops[seq[y]]);
}

__global__ void k_vb_keys(const int64_t *__restrict__ targets, int64_t n, int64_t *__restrict__ seq) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n) seq[x] = x;
  (void)targets;
}
__global__ void k_add_vb_sorted(DeviceKB kb, const int64_t *__restrict__ keys, const int64_t *__restrict__ seq,
                                const double *__restrict__ amounts, int64_t n) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const int64_t t = keys[x];
  if (x > 0 && keys[x - 1] == t) return;
  double v = kb.vB[t];
  for (int64_t y = x; y < n && keys[y] == t; y++) v = __dadd_rn(v, amounts[seq[y]]);   // CpuEngine.cpp:462
  kb.vB[t] = v;
}

static int bits_for(int64_t maxKey) {
  int b = 1;
  while (b < 63 && (maxKey >> b) != 0) b++;
  return b;
}

size_t train_sort_scratch_bytes(int64_t n) {
  // layout: keysIn | seqIn | keysA | seqA | keysOut | seqOut | digit table
  return sizeof(int64_t) * 6 * (size_t)n + sort_hist_bytes(n) + 256;
}

// dOps: nOps operations in sequence order (questions owned by this device, targets local); dScratch from
// train_sort_scratch_bytes(nOps).
void launch_train_ops_device_grouped(const DeviceKB &kb, const TrainOp *dOps, int64_t nOps, void *dScratch,
                                     size_t scratchBytes, cudaStream_t st) {
  if (nOps <= 0) return;
  if (scratchBytes < train_sort_scratch_bytes(nOps)) throw std::runtime_error("train sort scratch too small");
  int64_t *keysIn = (int64_t *)dScratch, *seqIn = keysIn + nOps, *keysA = seqIn + nOps, *seqA = keysA + nOps,
          *keysOut = seqA + nOps, *seqOut = keysOut + nOps;
  unsigned *hist = (unsigned *)(((uintptr_t)(seqOut + nOps) + 255) & ~(uintptr_t)255);
  const unsigned grid = (unsigned)((nOps + 255) / 256);
  k_train_keys<<<grid, 256, 0, st>>>(dOps, nOps, kb.T, keysIn, seqIn);
  count_launch();
  radix_sort_pairs(keysIn, seqIn, keysA, seqA, keysOut, seqOut, hist, nOps, bits_for(kb.Q * kb.T), st);   // stable: equal cells keep their sequence order
  k_train_ops_sorted<<<grid, 256, 0, st>>>(kb, dOps, keysOut, seqOut, nOps);
  count_launch();
}

// vB[target] += amount for n (target, amount) pairs in sequence order; dTargets is the key input.
void launch_add_vb_device_grouped(const DeviceKB &kb, const int64_t *dTargets, const double *dAmounts, int64_t n,
                                  void *dScratch, size_t scratchBytes, cudaStream_t st) {
  if (n <= 0) return;
  if (scratchBytes < train_sort_scratch_bytes(n)) throw std::runtime_error("train sort scratch too small");
  int64_t *seqIn = (int64_t *)dScratch + n, *keysA = seqIn + n, *seqA = keysA + n, *keysOut = seqA + n, *seqOut = keysOut + n;
  unsigned *hist = (unsigned *)(((uintptr_t)(seqOut + n) + 255) & ~(uintptr_t)255);
  const unsigned grid = (unsigned)((n + 255) / 256);
  k_vb_keys<<<grid, 256, 0, st>>>(dTargets, n, seqIn);
  count_launch();
  radix_sort_pairs(dTargets, seqIn, keysA, seqA, keysOut, seqOut, hist, n, bits_for(kb.T), st);
  k_add_vb_sorted<<<grid, 256, 0, st>>>(kb, keysOut, seqOut, dAmounts, n);
  count_launch();
}

} // namespace pqa
