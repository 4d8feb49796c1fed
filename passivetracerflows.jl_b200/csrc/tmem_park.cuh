// Tensor Memory (TMEM, 256 KB/SM on sm_100) used as per-thread PARKING space: tcgen05.st / tcgen05.ld in the 32x32b
// shape give every lane of a warp its own TMEM lane and N consecutive 32-bit columns, i.e. thread-private storage
// that does not occupy registers, shared memory or the LSU/shared-memory pipe.  The fused kernels park the second
// column / row of a pair here while the first one is transformed (DESIGN.md "Pairing through TMEM").
//
// A warp may only touch TMEM lanes [32*(warp%4), 32*(warp%4)+32); warps w and w+4 of a CTA therefore use different
// column ranges of the same lanes.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ptf {
namespace tmem {

// CTA-wide: warp 0 allocates NCOLS columns (power of two >= 32); returns the base address to every thread.
template <int NCOLS>
__device__ __forceinline__ uint32_t alloc_cta(uint32_t* smem_slot) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                     (unsigned long long)__cvta_generic_to_shared(smem_slot)),
                 "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  return *smem_slot;
}

template <int NCOLS>
__device__ __forceinline__ void free_cta(uint32_t base) {
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS));
}

// address of this warp's private column range: `cols_per_warp` columns starting at base
__device__ __forceinline__ uint32_t warp_addr(uint32_t base, int cols_per_warp) {
  const uint32_t warp = threadIdx.x >> 5;
  return base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * (uint32_t)cols_per_warp;
}

// park / fetch 4 double2 (16 words) at column offset `col`
__device__ __forceinline__ void st4(uint32_t addr, double2 a, double2 b, double2 c, double2 d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          addr),
      "r"(__double2loint(a.x)), "r"(__double2hiint(a.x)), "r"(__double2loint(a.y)), "r"(__double2hiint(a.y)),
      "r"(__double2loint(b.x)), "r"(__double2hiint(b.x)), "r"(__double2loint(b.y)), "r"(__double2hiint(b.y)),
      "r"(__double2loint(c.x)), "r"(__double2hiint(c.x)), "r"(__double2loint(c.y)), "r"(__double2hiint(c.y)),
      "r"(__double2loint(d.x)), "r"(__double2hiint(d.x)), "r"(__double2loint(d.y)), "r"(__double2hiint(d.y))
      : "memory");
}

__device__ __forceinline__ void ld4(uint32_t addr, double2& a, double2& b, double2& c, double2& d) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = make_double2(__hiloint2double(r[1], r[0]), __hiloint2double(r[3], r[2]));
  b = make_double2(__hiloint2double(r[5], r[4]), __hiloint2double(r[7], r[6]));
  c = make_double2(__hiloint2double(r[9], r[8]), __hiloint2double(r[11], r[10]));
  d = make_double2(__hiloint2double(r[13], r[12]), __hiloint2double(r[15], r[14]));
}

// single double2 (4 columns): no register-tuple constraints beyond the double2 itself
__device__ __forceinline__ void st1(uint32_t addr, double2 a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(__double2loint(a.x)),
               "r"(__double2hiint(a.x)), "r"(__double2loint(a.y)), "r"(__double2hiint(a.y))
               : "memory");
}
// NOTE: the caller must execute wait_ld() before using values returned by ld1_nowait
__device__ __forceinline__ void ld1_nowait(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ double2 unpack(const uint32_t (&r)[4]) {
  return make_double2(__hiloint2double(r[1], r[0]), __hiloint2double(r[3], r[2]));
}
// fetch N double2 parked at consecutive 4-column slots
template <int N>
__device__ __forceinline__ void ldn(uint32_t addr, double2 (&out)[N]) {
  uint32_t r[N][4];
#pragma unroll
  for (int i = 0; i < N; ++i) ld1_nowait(addr + 4 * i, r[i]);
  wait_ld();
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = unpack(r[i]);
}

__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// park / fetch a whole 16-element register array (64 columns) using compile-time slots
template <class SlotFn>
__device__ __forceinline__ void park16(uint32_t addr, const double2 (&v)[16], SlotFn slot) {
#pragma unroll
  for (int q = 0; q < 4; ++q) st4(addr + 16 * q, v[slot(4 * q)], v[slot(4 * q + 1)], v[slot(4 * q + 2)], v[slot(4 * q + 3)]);
  wait_st();
}
__device__ __forceinline__ void fetch16(uint32_t addr, double2 (&v)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) ld4(addr + 16 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// 256-bit global accesses (LDG.E.256 / STG.E.256): two adjacent double2 in one request
// NOT volatile: the compiler may schedule / batch these freely.  Only for data that no thread of the running kernel
// writes (a pure function of the address for the kernel's lifetime).
__device__ __forceinline__ void ldg256(const double2* p, double2& a, double2& b) {
  asm("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
// ordered (asm volatile) 64-bit streaming load: keeps the compiler from hoisting whole batches of loads
__device__ __forceinline__ double ldg64(const double* p) {
  double v;
  asm("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(double2* p, double2 a, double2 b) {
  asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

}  // namespace tmem
}  // namespace ptf
