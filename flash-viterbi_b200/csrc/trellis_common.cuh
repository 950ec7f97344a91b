// trellis_common.cuh — device helpers shared by the trellis kernels.
//
// The update every kernel must reproduce bit for bit (F:165-174):
//     pre  = tmp_i (+)f32 delta[k]                      float add, rounded to float
//     cand = (float)( (double)pre (+)f64 logA[k][i] )   double add, rounded to double, then to float
//     strict '>' from (-FLT_MAX, -1): lowest k among equal maxima, dead column -> (-FLT_MAX, -1)
// Kernels stream only hi = (float)logA (4 B per update) and compute the all-float estimate
//     est  = pre (+)f32 hi
// All operands are <= 0, so |logA| <= |pre + logA| and the estimate is within 2 float steps of
// cand.  The true first-argmax is therefore among the k whose estimate lies within
// WINDOW_STEPS float steps of the largest estimate; only those are re-evaluated exactly from the
// double table.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace flashv {

constexpr int WINDOW_STEPS = 4;
constexpr unsigned FULL_MASK = 0xffffffffu;

// Monotone integer image of a float: ord(a) < ord(b) <=> a < b (with -0 == +0).
__device__ __forceinline__ int ford(float f)
{
    int b = __float_as_int(f);
    return b >= 0 ? b : -(b & 0x7fffffff);
}

// The reference's rounding chain for one candidate, F:170.
__device__ __forceinline__ float exact_cand(float pre, double la)
{
    return __double2float_rn(__dadd_rn((double)pre, la));
}

// (value, index) under "larger value wins, then smaller index"; index INT_MAX = none.
struct Best {
    float x;
    int k;
};

__device__ __forceinline__ void best_take(Best &b, float x, int k)
{
    if (x > b.x || (x == b.x && k < b.k)) b.x = x, b.k = k;
}

__device__ __forceinline__ Best warp_best(Best b)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float ox = __shfl_xor_sync(FULL_MASK, b.x, off);
        int ok = __shfl_xor_sync(FULL_MASK, b.k, off);
        best_take(b, ox, ok);
    }
    return b;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, off));
    return v;
}

// Backpointer store: 16-bit rows when K < 65535 (0xFFFF = dead), else 32-bit.
__device__ __forceinline__ void psi_store(void *base, int psi16, size_t idx, int v)
{
    if (psi16)
        reinterpret_cast<uint16_t *>(base)[idx] = (uint16_t)(v < 0 ? 0xFFFFu : (unsigned)v);
    else
        reinterpret_cast<int32_t *>(base)[idx] = v;
}

__device__ __forceinline__ int psi_load(const void *base, int psi16, size_t idx)
{
    if (psi16) {
        unsigned v = reinterpret_cast<const uint16_t *>(base)[idx];
        return v == 0xFFFFu ? -1 : (int)v;
    }
    return reinterpret_cast<const int32_t *>(base)[idx];
}

// Exact resolution of one destination column for one vector, executed by a full warp.
//   chain_max[c] : this lane's largest estimate over its elements k = 4*(lane + 32*u) + c
//   col          : hiT + i*Kp (global), delta: the vector's delta (shared or global), Kp/128 chain length
// Returns (delta'[i], psi[i]) on every lane.
__device__ __forceinline__ Best resolve_column(const float (&chain_max)[4], float tmp, const float *__restrict__ col,
                                               const float *delta, const double *__restrict__ LAd, int K, int Kp,
                                               int i, int lane)
{
    float lane_max = fmaxf(fmaxf(chain_max[0], chain_max[1]), fmaxf(chain_max[2], chain_max[3]));
    float top = warp_max(lane_max);
    Best b{-FLT_MAX, 0x7fffffff};
    if (top > -FLT_MAX) {
        const int thr = ford(top) - WINDOW_STEPS;
        const int chain_len = Kp >> 7;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned hit = __ballot_sync(FULL_MASK, ford(chain_max[c]) >= thr);
            while (hit) {
                const int w = __ffs(hit) - 1;
                hit &= hit - 1;
                for (int u = lane; u < chain_len; u += 32) {
                    const int k = 4 * (w + 32 * u) + c;
                    if (k < K) {
                        const float pre = __fadd_rn(tmp, delta[k]);
                        const float est = __fadd_rn(pre, __ldg(col + k));
                        if (ford(est) >= thr) {
                            const float x = exact_cand(pre, __ldg(LAd + (size_t)k * K + i));
                            if (x > -FLT_MAX) best_take(b, x, k);
                        }
                    }
                }
            }
        }
        b = warp_best(b);
    }
    if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
    return b;
}

}  // namespace flashv
