// TEST INFRASTRUCTURE: force-included shim that lets g++ parse the MSVC-only reference sources.
// It adds no arithmetic of its own; it only maps MSVC keywords/intrinsics onto GCC equivalents.
#pragma once
#include <immintrin.h>
#include <x86intrin.h>
#include <alloca.h>
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cinttypes>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define __vectorcall
#define __fastcall
#define __cdecl
#define __stdcall
#define __declspec(x)
#define __pragma(x)
#define __forceinline inline
#define __assume(x) __builtin_unreachable()
#define _alloca alloca
#define __int64 long long
#define SRPLATFORM_API
#define PQACORE_API

static inline unsigned char _BitScanForward64(unsigned long *idx, unsigned long long v) {
  if (!v) return 0; *idx = (unsigned long)__builtin_ctzll(v); return 1;
}
static inline unsigned char _BitScanReverse64(unsigned long *idx, unsigned long long v) {
  if (!v) return 0; *idx = (unsigned long)(63 - __builtin_clzll(v)); return 1;
}
static inline unsigned char _BitScanForward(unsigned long *idx, unsigned long v) {
  if (!(uint32_t)v) return 0; *idx = (unsigned long)__builtin_ctz((uint32_t)v); return 1;
}
static inline unsigned char _BitScanReverse(unsigned long *idx, unsigned long v) {
  if (!(uint32_t)v) return 0; *idx = (unsigned long)(31 - __builtin_clz((uint32_t)v)); return 1;
}
static inline unsigned char _bittest64(const long long *a, long long b) {
  return (unsigned char)((((const uint64_t*)a)[b >> 6] >> (b & 63)) & 1);
}
static inline unsigned char _bittestandset64(long long *a, long long b) {
  uint64_t *p = ((uint64_t*)a) + (b >> 6); const uint64_t m = 1ULL << (b & 63);
  const unsigned char r = (*p & m) != 0; *p |= m; return r;
}
static inline unsigned char _bittestandreset64(long long *a, long long b) {
  uint64_t *p = ((uint64_t*)a) + (b >> 6); const uint64_t m = 1ULL << (b & 63);
  const unsigned char r = (*p & m) != 0; *p &= ~m; return r;
}
static inline unsigned char _bittestandcomplement64(long long *a, long long b) {
  uint64_t *p = ((uint64_t*)a) + (b >> 6); const uint64_t m = 1ULL << (b & 63);
  const unsigned char r = (*p & m) != 0; *p ^= m; return r;
}

// MSVC exposes vector lanes as union members (.m256d_f64[i] ...). The sed pass in build_ref.sh rewrites
// those member accesses into these lane-reference helpers.
template<typename E, typename V> static inline E& sr_lane(V& v, size_t i) { return reinterpret_cast<E*>(&v)[i]; }
template<typename E, typename V> static inline const E& sr_lane(const V& v, size_t i) {
  return reinterpret_cast<const E*>(&v)[i];
}
template<typename E, typename V> static inline E* sr_lanes(V& v) { return reinterpret_cast<E*>(&v); }
template<typename E, typename V> static inline const E* sr_lanes(const V& v) { return reinterpret_cast<const E*>(&v); }
