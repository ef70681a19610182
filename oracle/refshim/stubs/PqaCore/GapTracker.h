// TEST INFRASTRUCTURE stub of GapTracker.h: a fixed bitmap with the accessors the hot path reads.
// Bits beyond the logical size read as "gap" (reference GapTracker.h:12-15: SRBitArray(0, true)).
#pragma once
#include <vector>
namespace ProbQA {
template <typename taId> class GapTracker {
  std::vector<uint8_t> _bits; // padded with 0xff
  taId _n = 0, _nGaps = 0;
public:
  void Assign(const uint8_t *bits, const taId n) {
    _n = n; _nGaps = 0;
    _bits.assign(size_t(((n + 255) / 256) * 32 + 64), 0xff);
    for (taId i = 0; i < n; i++) {
      const bool g = bits && ((bits[i >> 3] >> (i & 7)) & 1);
      if (g) _nGaps++; else _bits[size_t(i >> 3)] &= uint8_t(~(1u << (i & 7)));
    }
  }
  bool IsGap(const taId at) const { return (_bits[size_t(at >> 3)] >> (at & 7)) & 1; }
  uint8_t GetQuad(const taId iQuad) const { return (_bits[size_t(iQuad >> 1)] >> ((iQuad & 1) << 2)) & 0x0f; }
  template<typename taResult> const taResult& GetPacked(const taId iPack) const {
    return reinterpret_cast<const taResult*>(_bits.data())[iPack];
  }
  taId GetNGaps() const { return _nGaps; }
};
} // namespace ProbQA
