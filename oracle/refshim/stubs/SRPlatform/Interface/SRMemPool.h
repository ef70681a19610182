// TEST INFRASTRUCTURE stub: plain aligned malloc behind the names the reference's drivers use for scratch
// layout (SRMemTotal/SRByteMem/SRMemItem/SRSmartMPP, reference SRMemPool.h:147-256). No pooling.
#pragma once
#include "../SRPlatform/Interface/SRSimd.h"
#include "../SRPlatform/Interface/SRCast.h"
namespace SRPlat {
class SRBaseMemPool {
public:
  void* AllocMem(const size_t nBytes) { return _mm_malloc(nBytes ? nBytes : 32, 32); }
  void ReleaseMem(void *p, const size_t) { _mm_free(p); }
};
template<typename taItem> class SRSmartMPP {
  SRBaseMemPool *_mp; taItem *_p; size_t _n;
public:
  SRSmartMPP(SRBaseMemPool &mp, const size_t nItems) : _mp(&mp), _n(nItems * sizeof(taItem)) {
    _p = static_cast<taItem*>(mp.AllocMem(_n));
  }
  SRSmartMPP(const SRSmartMPP&) = delete;
  ~SRSmartMPP() { if (_p) _mp->ReleaseMem(_p, _n); }
  taItem* Get() const { return _p; }
  taItem* Detach() { taItem *r = _p; _p = nullptr; return r; }
  void EarlyRelease() { if (_p) _mp->ReleaseMem(_p, _n); _p = nullptr; }
};
struct SRMemTotal { size_t _nBytes = 0; };
enum class SRMemPadding : uint8_t { None = 0, Left = 1, Right = 2, Both = 3 };
struct SRByteMem {
  size_t _offs;
  SRByteMem(const size_t nBytes, const SRMemPadding pad, SRMemTotal &mt) {
    const bool left = (uint8_t(pad) & 1) != 0, right = (uint8_t(pad) & 2) != 0;
    _offs = left ? SRSimd::GetPaddedBytes(mt._nBytes) : mt._nBytes;
    mt._nBytes = right ? SRSimd::GetPaddedBytes(_offs + nBytes) : (_offs + nBytes);
  }
  uint8_t* BytePtr(const SRSmartMPP<uint8_t> &b) const { return b.Get() + _offs; }
  template<typename P> P* ToPtr(const SRSmartMPP<uint8_t> &b) const { return reinterpret_cast<P*>(b.Get() + _offs); }
};
template<typename T> struct SRMemItem : public SRByteMem {
  SRMemItem(const size_t nItems, const SRMemPadding pad, SRMemTotal &mt) : SRByteMem(nItems * sizeof(T), pad, mt) {}
  T* Ptr(const SRSmartMPP<uint8_t> &b) const { return ToPtr<T>(b); }
};
} // namespace SRPlat
