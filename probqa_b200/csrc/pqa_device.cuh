// probqa_b200: device primitives shared by the kernels -- exact (order-preserving) Kahan summation, the
// reference's Log2Hot, split arithmetic, mbarrier / bulk-copy (TMA) wrappers.
// Compile with -fmad=false: every rounding below is explicit; fused multiply-adds appear only where written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pqa {

// ---------------------------------------------------------------------------------------------------------
// Kahan accumulator with the reference's exact operation order (SRPlatform/Interface/SRAccumulator.h:28-35,
// SRAccumVectDbl256.h:40-46): y = v - c; t = s + y; c = (t - s) - y; s = t.
struct Kahan {
  double s, c;
  __device__ __forceinline__ void init(double v = 0.0) { s = v; c = 0.0; }
  __device__ __forceinline__ void add(double v) {
    const double y = __dsub_rn(v, c);
    const double t = __dadd_rn(s, y);
    c = __dsub_rn(__dsub_rn(t, s), y);
    s = t;
  }
  __device__ __forceinline__ void neg() { s = -s; c = -c; }
  __device__ __forceinline__ double get() const { return __dsub_rn(s, c); }
};

// SRAccumVectDbl256::PreciseSum (SRAccumVectDbl256.h:83-91) of four Kahan lanes: scalar Kahan over
// corr[3],corr[2],corr[1],corr[0], negate, then sum[3..0]; result sum - corr. PairSum (:115-132) performs the
// same sequence on two accumulators at once, so this function reproduces both.
__device__ __forceinline__ double precise_sum4(const double s0, const double s1, const double s2, const double s3,
                                               const double c0, const double c1, const double c2, const double c3) {
  Kahan a; a.init(c3);
  a.add(c2); a.add(c1); a.add(c0);
  a.neg();
  a.add(s3); a.add(s2); a.add(s1); a.add(s0);
  return a.get();
}

// Four Kahan lanes held by the four threads of a lane group (lane l = threadIdx & 3 owns element j with j%4==l).
// Gathers the group's lanes with shuffles; every thread of the group returns the same bits. Only the four threads
// of the group need to be converged (the mask names exactly them).
__device__ __forceinline__ double group_precise_sum(const Kahan &k) {
  const unsigned mask = 0xFu << (threadIdx.x & 28u);
  const double s0 = __shfl_sync(mask, k.s, 0, 4), s1 = __shfl_sync(mask, k.s, 1, 4);
  const double s2 = __shfl_sync(mask, k.s, 2, 4), s3 = __shfl_sync(mask, k.s, 3, 4);
  const double c0 = __shfl_sync(mask, k.c, 0, 4), c1 = __shfl_sync(mask, k.c, 1, 4);
  const double c2 = __shfl_sync(mask, k.c, 2, 4), c3 = __shfl_sync(mask, k.c, 3, 4);
  return precise_sum4(s0, s1, s2, s3, c0, c1, c2, c3);
}

// Butterfly sum over the 32 lanes of a warp; every lane returns the total.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 1/x for normal x: MUFU.RCP64H seed (2^-23 relative) + one cubically convergent step (3 DFMA), ~1 ulp.
// Not IEEE-rounded; used only by the tolerance-level kernel.
__device__ __forceinline__ double fast_rcp(const double x) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = __fma_rn(-x, y0, 1.0);
  const double e2 = __fma_rn(e, e, e);
  return __fma_rn(y0, e2, y0);
}
// Same seed + one Newton step (2 DFMA): relative error <= ~2^-46 = 1.4e-14. Enough for the lack sum of the staged
// kernel (tolerance 1e-12), one fp64 instruction per element cheaper.
__device__ __forceinline__ double fast_rcp46(const double x) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = __fma_rn(-x, y0, 1.0);
  return __fma_rn(y0, e, y0);
}

// ---------------------------------------------------------------------------------------------------------
// SRVectMath::Log2Hot (SRPlatform/Interface/SRVectMath.h:87-135), one lane. Bit-exact: every operation is the
// IEEE operation the AVX2 code performs (div, mul, two FMAs, add); the table is the reference's
// (SRVectMath.cpp:30-44) uploaded by the host. x = 0 gives -1023, no special cases.
__device__ __forceinline__ double log2hot(const double x, const double *__restrict__ tbl) {
  const int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  const int e = (hi >> 20) - 1023;                       // arithmetic shift, sign bit not cleared (:96-98)
  const int idx = (hi >> 10) & 1023;                     // top 10 mantissa bits (:101-102)
  const int zhi = (hi & (int)0x800FFFFF) | 0x3FF00000;   // exponent := 0 (:88-89)
  const double z = __hiloint2double(zhi, lo);
  // mid-point of the table bucket: low 42 bits of z replaced by 2^41 (:108)
  const double m = __hiloint2double((zhi & (int)0xFFFFFC00) | 0x00000200, 0);
  const double y = tbl[idx];
  const double t = __ddiv_rn(__dsub_rn(z, m), __dadd_rn(z, m)); // (:111-114)
  const double t2 = __dmul_rn(t, t);
  const double t3 = __dmul_rn(t, t2);
  const double terms01 = __fma_rn(1.0 / 3, t3, t);       // (:117)
  const double l2z = __fma_rn(terms01, 2.8853900817779268147198493620038, y); // (:122)
  return __dadd_rn(l2z, (double)e);                      // (:131-133)
}

// ---------------------------------------------------------------------------------------------------------
// SRPoolRunner::CalcSplit (SRPlatform/Interface/SRPoolRunner.h:96-110): piece p of nItems over nWorkers.
__device__ __host__ __forceinline__ int64_t split_count(int64_t nItems, int64_t nWorkers) {
  return nItems < nWorkers ? nItems : nWorkers;
}
__device__ __host__ __forceinline__ int64_t split_start(int64_t nItems, int64_t nWorkers, int64_t p) {
  const int64_t quot = nItems / nWorkers, rem = nItems % nWorkers;
  return p * quot + (p < rem ? p : rem);
}

__device__ __forceinline__ bool bit32(const uint32_t *bits, int64_t i) {
  return bits != nullptr && ((bits[i >> 5] >> (i & 31)) & 1u);
}
__device__ __forceinline__ bool bit64(const uint64_t *bits, int64_t i) {
  return (bits[i >> 6] >> (i & 63)) & 1ull;
}

// ---------------------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) wrappers.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dstSmem, const void *srcGmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dstSmem)),
               "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// orders this thread's generic-proxy shared-memory accesses before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

} // namespace pqa
