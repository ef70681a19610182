// TEST INFRASTRUCTURE stub: the SRUtils helpers reachable from the files compiled here (the radix-sort top-k branch
// zero-fills its counters, CERadixSortRatingsSubtaskSort.cpp:43; the ResumeQuiz multiply flushes cache lines after a
// block, CEUpdatePriorsSubtaskMul.cpp:86-97 -- a performance hint with no arithmetic, a no-op here).
#pragma once
namespace SRPlat {
class SRUtils {
public:
  template<bool taCache> static void FillZeroVects(__m256i *p, const size_t nVects) { memset(p, 0, nVects * 32); }
  template<bool taFlushLeft, bool taFlushRight> static void FlushCache(const void *, const size_t) {}
};
} // namespace SRPlat
