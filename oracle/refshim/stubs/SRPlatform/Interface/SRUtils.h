// TEST INFRASTRUCTURE stub: the single SRUtils helper reachable from the files compiled here (the radix-sort
// top-k branch zero-fills its counters, CERadixSortRatingsSubtaskSort.cpp:43).
#pragma once
namespace SRPlat {
class SRUtils {
public:
  template<bool taCache> static void FillZeroVects(__m256i *p, const size_t nVects) { memset(p, 0, nVects * 32); }
};
} // namespace SRPlat
