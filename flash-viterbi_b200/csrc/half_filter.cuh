// half_filter.cuh — the half-precision filter shared by the pass kernel (flash_persistent.cu: k_flash_persist16)
// and the level kernel (flash_kernels.cu: k_flash_level16): tensor-memory wrappers, the window scan over 256
// chains of 16 sources, and the exact evaluation from the chain-major double table.  The derivation of the window
// is in flash_persistent.cu above k_flash_persist16 and in DESIGN.md §4.
#pragma once

#include <cuda_fp16.h>

#include "trellis_common.cuh"

namespace flashv {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int TMEM_COLS = 512;     // the whole tensor memory of the SM
__device__ __forceinline__ void tmem_alloc(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float4 &v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr)
{
    uint32_t x, y, z, w;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr) : "memory");
    return make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(w));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int NO_CHAIN = 0x7fff;

__device__ __forceinline__ float unford(int o) { return __int_as_float(o >= 0 ? o : (int)((unsigned)(-o) | 0x80000000u)); }

constexpr int H_PAIRS = 4;  // pairs of chains the fast scan holds (lanes 0-15 one chain, 16-31 the other)
constexpr float H_CLAMP = -60000.f, H_TRUST = -30000.f;

__device__ __forceinline__ __half2 u2h(uint32_t w) { return *reinterpret_cast<const __half2 *>(&w); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<const uint32_t *>(&h); }

__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st4_raw(uint32_t taddr, const uint4 &v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}


// Park the first n_it iterations (256 source states each) of a warp's two columns of the tiled half table in its
// tensor-memory slot: columns 8 it .. 8 it + 3 hold the first column's eight halves per lane, 8 it + 4 .. + 7 the
// second's.  Eight 16-byte loads are requested before any of them is stored: the tcgen05.st is volatile, so a load
// written next to its store would wait for memory alone, once per iteration and column (32 round trips per launch).
__device__ __forceinline__ void tmem_fill16(uint32_t tbase, const __half *__restrict__ slab, int ncols, int rs0, int rs1, int n_it,
                                            int lane)
{
    for (int it0 = 0; it0 < n_it; it0 += 4) {
        uint4 va[4], vb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int it = min(it0 + e, n_it - 1);
            va[e] = __ldg(reinterpret_cast<const uint4 *>(slab + ((size_t)it * ncols + rs0) * 256) + lane);
            vb[e] = __ldg(reinterpret_cast<const uint4 *>(slab + ((size_t)it * ncols + rs1) * 256) + lane);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (it0 + e < n_it) {
                tmem_st4_raw(tbase + 8u * (uint32_t)(it0 + e), va[e]);
                tmem_st4_raw(tbase + 8u * (uint32_t)(it0 + e) + 4u, vb[e]);
            }
    }
}

template <int P>
struct Scan16T {
    double la[P];  // pair p: lanes 0-15 hold the elements of chain qa[p], lanes 16-31 those of qb[p]
    int qa[P], qb[P];  // warp-uniform; NO_CHAIN = empty
    uint32_t thr2;       // the window threshold, twice, as half2 bits
    bool live, overflow;  // warp-uniform
};

// Chains inside the window and their doubles, one column; straight-line like scan_fetch() above.
using Scan16 = Scan16T<H_PAIRS>;

template <int P>
__device__ __forceinline__ void scan16_fetch(Scan16T<P> &sc, const __half2 (&m)[4], float tmp, float c, const double *__restrict__ LAc16,
                                             int i, bool have, int lane)
{
    const __half2 mm = __hmax2(__hmax2(m[0], m[1]), __hmax2(m[2], m[3]));
    const float top = unford(__reduce_max_sync(FULL_MASK, ford(fmaxf(__low2float(mm), __high2float(mm)))));
    sc.live = have && top > -INFINITY;  // all estimates -inf: no source has an edge into this state
    const float atop = fabsf(top);
    const float W = 2.1f * 0x1p-10f * atop + 0x1p-20f + 0x1p-20f * (fabsf(tmp) + fabsf(c) + atop);
    // below H_TRUST the clamp may have distorted the estimates: everything is inside the window then
    const __half thr = top >= H_TRUST ? __float2half_rd(top - W) : __float2half_rd(-INFINITY);
    const __half2 thr2 = __half2half2(thr);
    sc.thr2 = h2u(thr2);
    unsigned m8 = 0;  // this lane's chains inside the window: bit 2w + h <-> chain 8 * lane + 2w + h
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const unsigned ge = __hge2_mask(m[w], thr2);
        m8 |= ((ge & 1u) | ((ge >> 15) & 2u)) << (2 * w);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) sc.la[p] = -INFINITY, sc.qa[p] = sc.qb[p] = NO_CHAIN;
    const double *col = LAc16 + (size_t)i * 4096 + (lane & 15);
    int lq = m8 ? 8 * lane + __ffs(m8) - 1 : NO_CHAIN;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int qa = __reduce_min_sync(FULL_MASK, lq);
        if (lq == qa) m8 &= m8 - 1;
        lq = m8 ? 8 * lane + __ffs(m8) - 1 : NO_CHAIN;
        const int qb = __reduce_min_sync(FULL_MASK, lq);
        if (lq == qb) m8 &= m8 - 1;
        lq = m8 ? 8 * lane + __ffs(m8) - 1 : NO_CHAIN;
        if (!sc.live || qa == NO_CHAIN) break;  // warp-uniform
        sc.qa[p] = qa, sc.qb[p] = qb;
        const int q = lane < 16 ? qa : qb;
        if (q != NO_CHAIN) sc.la[p] = __ldg(col + q * 16);
        if (qb == NO_CHAIN) break;
    }
    sc.overflow = __any_sync(FULL_MASK, m8 != 0);
}

// More chains inside the window than the fast scan holds: all of them, two per trip, synchronously.
static __device__ __noinline__ Best scan16_slow(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t thr2, float tmp,
                                         const float *sdelta, const double *__restrict__ LAc16, int K, int i, int lane)
{
    Best acc{-FLT_MAX, 0x7fffffff};
    const uint32_t mw[4] = {m0, m1, m2, m3};
    const double *col = LAc16 + (size_t)i * 4096 + (lane & 15);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const unsigned ge = __hge2_mask(u2h(mw[w]), u2h(thr2));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned hit = __ballot_sync(FULL_MASK, (ge >> (16 * h)) & 1u);
            while (hit) {
                const int la_ = __ffs(hit) - 1;
                hit &= hit - 1;
                int lb_ = -1;
                if (hit) lb_ = __ffs(hit) - 1, hit &= hit - 1;
                const int src = lane < 16 ? la_ : lb_;
                if (src >= 0) {
                    const int q = 8 * src + 2 * w + h, k = q + 256 * (lane & 15);
                    if (k < K) {
                        const float x = exact_cand(__fadd_rn(tmp, sdelta[k]), __ldg(col + q * 16));
                        if (x > -FLT_MAX) best_take(acc, x, k);
                    }
                }
            }
        }
    }
    return acc;
}

template <int P>
__device__ __forceinline__ Best scan16_settle(const Scan16T<P> &sc, const __half2 (&m)[4], float tmp, const float *sdelta,
                                              const double *__restrict__ LAc16, int K, int i, int lane)
{
    Best acc{-FLT_MAX, 0x7fffffff};
    if (sc.live) {
        if (sc.overflow) {
            acc = scan16_slow(h2u(m[0]), h2u(m[1]), h2u(m[2]), h2u(m[3]), sc.thr2, tmp, sdelta, LAc16, K, i, lane);
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (sc.qa[p] == NO_CHAIN) break;  // warp-uniform
                const int q = lane < 16 ? sc.qa[p] : sc.qb[p];
                const int k = q + 256 * (lane & 15);  // every element of a chain inside the window is evaluated exactly
                if (q != NO_CHAIN && k < K) {
                    const float x = exact_cand(__fadd_rn(tmp, sdelta[k]), sc.la[p]);
                    if (x > -FLT_MAX) best_take(acc, x, k);
                }
            }
        }
    }
    const int ox = ford(acc.x);
    const int mo = __reduce_max_sync(FULL_MASK, ox);
    Best b;
    b.k = __reduce_min_sync(FULL_MASK, ox == mo ? acc.k : 0x7fffffff);
    b.x = unford(mo);  // a -0 would come back as +0; it cannot occur (see scan_settle)
    if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
    return b;
}


}  // namespace flashv
