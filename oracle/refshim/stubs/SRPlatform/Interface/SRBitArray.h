// TEST INFRASTRUCTURE stub: only the stateless bit helpers of SRBitArray.h:10-39 that the hot path calls.
#pragma once
#include "../SRPlatform/Interface/SRCast.h"
namespace SRPlat {
class SRBitHelper {
public:
  static bool Test(const __m256i *pArray, const int64_t iBit) {
    return (reinterpret_cast<const uint8_t*>(pArray)[iBit >> 3] >> (iBit & 7)) & 1;
  }
  static bool Set(__m256i *pArray, const int64_t iBit) {
    uint8_t &b = reinterpret_cast<uint8_t*>(pArray)[iBit >> 3];
    const bool old = (b >> (iBit & 7)) & 1;
    b |= uint8_t(1u << (iBit & 7));
    return old;
  }
  template<typename taResult> static const taResult& GetPacked(const __m256i *pArray, const uint64_t iPack) {
    return reinterpret_cast<const taResult*>(pArray)[iPack];
  }
};
} // namespace SRPlat
