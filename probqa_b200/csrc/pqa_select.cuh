// probqa_b200: question selection shared by k_select_question (pqa_kernels.cu) and the fused one-quiz kernel
// (pqa_eval_staged.cu). CpuEngine::NextQuestionSpec (CpuEngine.cpp:337-415) after the evaluation:
//   chunks = CalcSplit(Q, 8W); per chunk a scalar-Kahan running sum over its questions, asked/gap questions repeat
//   the running value (CEEvalQsSubtaskConsider.cpp:54-58,212-214); grand totals = scalar Kahan over the chunks' last
//   values (CpuEngine.cpp:362-374); r = totG * u64 / (2^64-1) (SRDoubleNumber.h:35-39); upper_bound over chunks,
//   upper_bound inside the chunk (:380-400); asked/gap -> nearest available question (BaseEngine.cpp:60-124).
#pragma once
#include "pqa_kernels.cuh"
#include "pqa_device.cuh"

namespace pqa {

__device__ __forceinline__ uint64_t avail_word(const DeviceKB &kb, const uint64_t *asked, int64_t w) {
  // bits of (qgaps | asked) for questions 64w..64w+63, complemented; questions >= Q read as gaps (GapTracker.h:12-15)
  uint64_t g = 0;
  if (kb.qgaps) g = (uint64_t)kb.qgaps[2 * w] | ((uint64_t)kb.qgaps[2 * w + 1] << 32);
  const int64_t rem = kb.Q - 64 * w;
  if (rem < 64) g |= ~0ull << rem;
  return ~(g | asked[w]);
}

__device__ __forceinline__ int64_t find_nearest_question(const DeviceKB &kb, const uint64_t *asked, int64_t iMiddle) {
  const uint32_t dInf = 200;
  const int64_t iPack = iMiddle >> 6;
  const uint32_t iWithin = (uint32_t)(iMiddle & 63);
  const uint64_t available = avail_word(kb, asked, iPack);
  if (available != 0) {
    const uint64_t baseMask = (1ull << iWithin) - 1;
    const uint64_t higher = available & ~baseMask, lower = available & baseMask;
    const uint32_t dHigher = higher ? (uint32_t)(__ffsll((long long)higher) - 1) - iWithin : dInf;
    const uint32_t dLower = lower ? iWithin - (uint32_t)(63 - __clzll((long long)lower)) : dInf;
    if (dHigher < dLower) return iMiddle + dHigher;
    return iMiddle - dLower;
  }
  const int64_t limPack = (kb.Q + 63) >> 6;
  int64_t i = 1;
  while (iPack >= i && iPack + i < limPack) {
    const uint64_t availLeft = avail_word(kb, asked, iPack - i), availRight = avail_word(kb, asked, iPack + i);
    if ((availLeft | availRight) == 0) { i++; continue; }
    const uint32_t dHigher = availRight ? (uint32_t)(__ffsll((long long)availRight) - 1) + 64 - iWithin : dInf;
    const uint32_t dLower = availLeft ? iWithin + 64 - (uint32_t)(63 - __clzll((long long)availLeft)) : dInf;
    if (dHigher < dLower) return iMiddle + dHigher + ((i - 1) << 6);
    return iMiddle - dLower - ((i - 1) << 6);
  }
  while (iPack >= i) {
    const uint64_t availLeft = avail_word(kb, asked, iPack - i);
    if (!availLeft) { i++; continue; }
    const uint32_t dLower = iWithin + 64 - (uint32_t)(63 - __clzll((long long)availLeft));
    return iMiddle - dLower - ((i - 1) << 6);
  }
  while (iPack + i < limPack) {
    const uint64_t availRight = avail_word(kb, asked, iPack + i);
    if (!availRight) { i++; continue; }
    const uint32_t dHigher = (uint32_t)(__ffsll((long long)availRight) - 1) + 64 - iWithin;
    return iMiddle + dHigher + ((i - 1) << 6);
  }
  return -1;
}

__device__ __forceinline__ int64_t upper_bound_d(const double *a, int64_t n, double v) {
  int64_t lo = 0, len = n;
  while (len > 0) {
    const int64_t half = len >> 1;
    if (!(v < a[lo + half])) { lo += half + 1; len -= half + 1; } else len = half;
  }
  return lo;
}


// The whole selection for quiz b (slot `slot`, priorities pri[Q], 64-bit draw `random`), executed by one CTA.
// sGrand: shared memory for nChunks doubles. runLength[Q] is required (the in-chunk binary search reads it back).
// Returns (to thread 0 only) the chosen question or -1; other threads return -2.
__device__ __forceinline__ int64_t select_question_cta(const DeviceKB &kb, const QuizPool &qp, int64_t slot, const double *pri,
                                                       uint64_t random, int W, double *runLength, double *grandOut,
                                                       bool wantQuestion, int setActive, double *sGrand,
                                                       bool *sawAnomaly = nullptr) {
  const int64_t Q = kb.Q;
  const uint64_t *asked = qp.asked + slot * qp.askedWords;
  const int64_t nW = (int64_t)W * 8, nChunks = split_count(Q, nW);
  unsigned bad = 0;      // priorities of available questions that are <= 0 or not finite: :209-211 warns and adds them anyway
  for (int64_t c = threadIdx.x; c < nChunks; c += blockDim.x) {
    const int64_t first = split_start(Q, nW, c), limit = split_start(Q, nW, c + 1);
    Kahan run; run.init(0.0);
    for (int64_t i = first; i < limit; i++) {
      if (!(bit32(kb.qgaps, i) || bit64(asked, i))) {
        const double v = pri[i];
        const long long vb = __double_as_longlong(v);                     // positive and finite <=> 0 < bits < those of +inf
        bad += !(vb > 0ll && vb < 0x7FF0000000000000ll);
        run.add(v);
      }
      runLength[i] = run.get();
    }
    sGrand[c] = run.get();
  }
  if (bad && kb.anomalies) atomicAdd(kb.anomalies + 0, (unsigned long long)bad);
  bool anomaly = __syncthreads_or(bad != 0) != 0;
  if (threadIdx.x != 0) return -2;
  Kahan tot; tot.init(0.0);
  for (int64_t c = 0; c < nChunks; c++) {  // CpuEngine.cpp:362-374
    tot.add(sGrand[c]);
    sGrand[c] = tot.get();
    if (grandOut) grandOut[c] = sGrand[c];
  }
  const double totG = sGrand[nChunks - 1];
  // running totals, once not finite, stay so: the last one tells (CpuEngine.cpp:368-371 checks each and carries on)
  if (!(fabs(totG) < __longlong_as_double(0x7FF0000000000000ll))) { anomaly = true; if (kb.anomalies) atomicAdd(kb.anomalies + 1, 1ull); }
  if (wantQuestion && !(totG > 0.0)) { anomaly = true; if (kb.anomalies) atomicAdd(kb.anomalies + 2, 1ull); }   // :375-377
  if (sawAnomaly) *sawAnomaly = anomaly;
  if (!wantQuestion) return -2;
  // SRDoubleNumber::MakeRandom: upper * rnd / max (left to right, both factors converted to double)
  const double sel = __ddiv_rn(__dmul_rn(totG, __ull2double_rn(random)), __ull2double_rn(~0ull));
  const int64_t iWorker = upper_bound_d(sGrand, nChunks, sel);
  int64_t chosen;
  if (iWorker >= nChunks) {
    chosen = Q - 1;                                                       // :382-386
  } else {
    const double inWorker = __dsub_rn(sel, iWorker == 0 ? 0.0 : sGrand[iWorker - 1]);  // :388
    const int64_t first = split_start(Q, nW, iWorker), limit = split_start(Q, nW, iWorker + 1);
    chosen = first + upper_bound_d(runLength + first, limit - first, inWorker);       // :391
    if (chosen >= limit) chosen = limit - 1;                              // :392-400
  }
  if (bit32(kb.qgaps, chosen) || bit64(asked, chosen)) chosen = find_nearest_question(kb, asked, chosen);  // :404-406
  if (setActive && chosen >= 0) qp.active[slot] = chosen;                 // :412
  return chosen;
}

} // namespace pqa
