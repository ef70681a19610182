// probqa_b200: large RecordQuizTarget / Train batches (BASELINE config 5). The reference applies a quiz' answers in
// sequence (CETrainOperation.cpp:15-83); operations on different (question, target) cells commute, operations on the same
// cell must keep their sequence order. For big batches the grouping by cell is done on the device: a stable radix sort of
// (cell key, sequence number) pairs (CUB, part of the CUDA toolkit), then one thread per sorted position -- the thread that
// sits at the head of a run of equal keys applies the whole run in sequence order with the same arithmetic as k_train_ops.
// Results are bit-identical to the host-grouped path (tests/test_gpu_parity.py::test_record_quiz_target_and_train_bit_exact
// runs both).
#include <cub/device/device_radix_sort.cuh>

#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

namespace pqa {
void count_launch();

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
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const int64_t *)nullptr, (int64_t *)nullptr, (const int64_t *)nullptr,
                                  (int64_t *)nullptr, (int)n, 0, 63);
  // layout: keysIn | seqIn | keysOut | seqOut | cub temp
  return sizeof(int64_t) * 4 * (size_t)n + tmp + 256;
}

// dOps: nOps operations in sequence order (questions owned by this device, targets local); dScratch from
// train_sort_scratch_bytes(nOps).
void launch_train_ops_device_grouped(const DeviceKB &kb, const TrainOp *dOps, int64_t nOps, void *dScratch,
                                     size_t scratchBytes, cudaStream_t st) {
  if (nOps <= 0) return;
  int64_t *keysIn = (int64_t *)dScratch, *seqIn = keysIn + nOps, *keysOut = seqIn + nOps, *seqOut = keysOut + nOps;
  void *tmp = (void *)(((uintptr_t)(seqOut + nOps) + 255) & ~(uintptr_t)255);
  size_t tmpBytes = scratchBytes - ((char *)tmp - (char *)dScratch);
  const unsigned grid = (unsigned)((nOps + 255) / 256);
  k_train_keys<<<grid, 256, 0, st>>>(dOps, nOps, kb.T, keysIn, seqIn);
  count_launch();
  cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, seqIn, seqOut, (int)nOps, 0,
                                  bits_for(kb.Q * kb.T), st);   // stable: equal cells keep their sequence order
  count_launch();
  k_train_ops_sorted<<<grid, 256, 0, st>>>(kb, dOps, keysOut, seqOut, nOps);
  count_launch();
}

// vB[target] += amount for n (target, amount) pairs in sequence order; dTargets doubles as the key input.
void launch_add_vb_device_grouped(const DeviceKB &kb, const int64_t *dTargets, const double *dAmounts, int64_t n,
                                  void *dScratch, size_t scratchBytes, cudaStream_t st) {
  if (n <= 0) return;
  int64_t *seqIn = (int64_t *)dScratch + n, *keysOut = seqIn + n, *seqOut = keysOut + n;
  void *tmp = (void *)(((uintptr_t)(seqOut + n) + 255) & ~(uintptr_t)255);
  size_t tmpBytes = scratchBytes - ((char *)tmp - (char *)dScratch);
  const unsigned grid = (unsigned)((n + 255) / 256);
  k_vb_keys<<<grid, 256, 0, st>>>(dTargets, n, seqIn);
  count_launch();
  cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, dTargets, keysOut, seqIn, seqOut, (int)n, 0, bits_for(kb.T), st);
  count_launch();
  k_add_vb_sorted<<<grid, 256, 0, st>>>(kb, keysOut, seqOut, dAmounts, n);
  count_launch();
}

} // namespace pqa
