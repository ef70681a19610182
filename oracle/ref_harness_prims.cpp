// TEST INFRASTRUCTURE: C-ABI harness around the REFERENCE's own numeric primitives, compiled from
// /root/reference/ProbQA sources (scratch-patched for g++ by build_ref.sh) into oracle/_ref/libpqa_ref.so.
// Used only by tests/ to pin the oracle restatement bit-for-bit. Contains no arithmetic of its own.
#include "../SRPlatform/Interface/SRVectMath.h"
#include "../SRPlatform/Interface/SRAccumVectDbl256.h"
#include "../SRPlatform/Interface/SRAccumulator.h"
#include "../SRPlatform/Interface/SRHeap.h"
#include "../SRPlatform/Interface/SRThreadPool.h"
#include "../SRPlatform/Interface/SRStandardSubtask.h"
#include "../SRPlatform/Interface/SRPoolRunner.h"

using namespace SRPlat;

struct RefRatedTarget { // Interface/PqaCommon.h:54-61
  int64_t _iTarget; double _prob;
  bool operator<(const RefRatedTarget& f) const { return _prob < f._prob; }
};
struct RefHeadItem { // RatingsHeap.h:11-20
  double _prob; int64_t _iSource;
  bool operator<(const RefHeadItem& f) const { return _prob < f._prob; }
};

extern "C" {

void ref_log2hot(const double *x, double *out, int64_t n) { // n must be a multiple of 4
  for (int64_t i = 0; i < n; i += 4) {
    _mm256_storeu_pd(out + i, SRVectMath::Log2Hot(_mm256_loadu_pd(x + i)));
  }
}

// Feed nVects 4-wide vectors into an SRAccumVectDbl256 and report every way of reading it out.
void ref_v4_accumulate(const double *values, int64_t nVects, double *preciseSum, double *fullSum) {
  SRAccumVectDbl256 acc;
  for (int64_t i = 0; i < nVects; i++) acc.Add(_mm256_loadu_pd(values + 4 * i));
  *preciseSum = acc.PreciseSum();
  *fullSum = acc.GetFullSum();
}

void ref_v4_pair(const double *a, const double *b, int64_t nVects, double *sumA, double *sumB) {
  SRAccumVectDbl256 accA, accB;
  for (int64_t i = 0; i < nVects; i++) {
    accA.Add(_mm256_loadu_pd(a + 4 * i));
    accB.Add(_mm256_loadu_pd(b + 4 * i));
  }
  *sumA = accA.PairSum(accB, *sumB);
}

// Scalar-lane Add(at, value) variant followed by PairSum (CEEvalQsSubtaskConsider.cpp:163-175 usage)
void ref_v4_pair_at(const double *a, const double *b, int64_t n, double *sumA, double *sumB) {
  SRAccumVectDbl256 accA, accB;
  const int64_t nVec = (n >> 2) << 2;
  for (int64_t i = 0; i < nVec; i += 4) { accA.Add(_mm256_loadu_pd(a + i)); accB.Add(_mm256_loadu_pd(b + i)); }
  for (int64_t i = nVec; i < n; i++) {
    accA.Add(static_cast<SRVectCompCount>(i - nVec), a[i]);
    accB.Add(static_cast<SRVectCompCount>(i - nVec), b[i]);
  }
  *sumA = accA.PairSum(accB, *sumB);
}

double ref_kahan_scalar(const double *v, int64_t n) {
  SRAccumulator<SRDoubleNumber> acc(SRDoubleNumber::FromDouble(0.0));
  for (int64_t i = 0; i < n; i++) acc.Add(SRDoubleNumber::FromDouble(v[i]));
  return acc.Get().GetValue();
}

// The reference's SRAccumulatorTest KAT objects need direct state injection (test helper is a friend class in
// the reference); emulate by choosing inputs instead -- see tests.

void ref_make_heap(RefRatedTarget *p, int64_t n) { std::make_heap(p, p + n); }
void ref_pop_heap(RefRatedTarget *p, int64_t n) { std::pop_heap(p, p + n); }
void ref_head_make_heap(RefHeadItem *p, int64_t n) { std::make_heap(p, p + n); }
void ref_head_pop_heap(RefHeadItem *p, int64_t n) { std::pop_heap(p, p + n); }
void ref_head_down(RefHeadItem *p, int64_t n) { SRHeapHelper::Down(p, p + n); }

int64_t ref_calc_split(int64_t nItems, int64_t nWorkers, size_t *bounds) {
  SRPoolRunner::Split s = SRPoolRunner::CalcSplit(bounds, (size_t)nItems, (SRSubtaskCount)nWorkers);
  return (int64_t)s._nSubtasks;
}

const double *ref_log2_table_probe(double *out1024) { // table values through Log2Hot cannot be read directly; skip
  (void)out1024; return nullptr;
}

} // extern "C"
