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

// Window threshold for the sum-first estimate.  For source k let R = tmp + delta[k] + logA[k][i]
// (reals; all three <= 0) and U the float spacing at |R|.  The reference's candidate (F:170) is
//     cand = fl32( fl64( fl32(tmp + delta[k]) + logA ) ):  two float roundings of sums no larger in
//            magnitude than R (U/2 each; U if the last one crosses into the next binade) plus a
//            double rounding (2^-29 U)                                  |cand - R|      <= 1.5 U + eps
// and the kernel's estimate is m2 = fl32(delta[k] + fl32(logA)), again two roundings of at most
// U/2:                                                                  |m2 + tmp - R|  <= 1 U
// so |cand - (m2 + tmp)| <= 2.5 U + eps (1.5 U observed, tests/test_host_logic.py).  If k* is the
// reference's argmax and kt the argmax of m2:
//     m2(k*) + tmp >= cand(k*) - 2.5U >= cand(kt) - 2.5U >= m2(kt) + tmp - 5U.
// Candidates near the top may sit one binade above c = fl32(top + tmp), where the spacing doubles,
// so with G = the spacing at |c|:  m2(k*) >= top - 10 G - eps.  The threshold is top - 12 G, formed
// with one rounding (error <= G/2).  Ties of cand are all inside the window, so the lowest index
// among equal maxima is found by the exact re-evaluation, as the reference's strict '>' does.
__device__ __forceinline__ float filter_threshold(float top, float tmp)
{
    const float c = __fadd_rn(top, tmp);
    int e = (__float_as_int(c) >> 23) & 0xff;
    e = max(e - 23, 1);
    const float G = __int_as_float(e << 23);
    return __fmaf_rn(-12.0f, G, top);
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

// Exact resolution of P (column, vector) pairs at once, executed by a full warp.  The latency of a
// resolution is two dependent memory round trips (the winning chain's estimates from L2, then the
// candidate's double from HBM); doing the pairs one after the other would pay them P times, so the
// three phases below each run over all pairs before the next one starts:
//   1. per pair: largest estimate, window threshold, the (first) chain inside the window; every
//      lane loads its element of that chain
//   2. per pair: the candidate (the common case is exactly one in the whole warp) loads its double
//   3. per pair: exact value on the candidate lane, broadcast; anything unusual (several chains or
//      several candidates inside the window, chains longer than a warp) falls back to
//      resolve_column()
// cm[p] are the lane's chain maxima of pair p, col[p] its hiT column, delta[p] its delta vector
// (shared memory), out[p] receives (delta', psi) on every lane.
template <int P>
__device__ __forceinline__ void resolve_tile(const float (&cm)[P][4], const float (&tmp)[P], const float *const (&col)[P],
                                             const float *const (&delta)[P], const int (&icol)[P],
                                             const double *__restrict__ LAd, int K, int Kp, int lane, Best (&out)[P])
{
    const int chain_len = Kp >> 7;
    float h[P];
    int thr[P], kk[P];
    unsigned slow = 0;  // bit p: pair p needs the general path (warp-uniform)
    unsigned dead = 0;  // bit p: no finite estimate
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const float top = warp_max(fmaxf(fmaxf(cm[p][0], cm[p][1]), fmaxf(cm[p][2], cm[p][3])));
        thr[p] = ford(top) - WINDOW_STEPS;
        h[p] = 0.f, kk[p] = -1;
        if (!(top > -FLT_MAX)) {
            dead |= 1u << p;
            continue;
        }
        int nchain = 0, w = 0, c = 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const unsigned hit = __ballot_sync(FULL_MASK, ford(cm[p][cc]) >= thr[p]);
            if (hit && nchain == 0) w = __ffs(hit) - 1, c = cc;
            nchain += __popc(hit);
        }
        if (nchain != 1 || chain_len > 32) {
            slow |= 1u << p;
            continue;
        }
        const int k = 4 * (w + 32 * lane) + c;
        if (lane < chain_len && k < K) kk[p] = k, h[p] = __ldg(col[p] + k);
    }
    float pre[P];
    double la[P];
    int cand_lane[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        pre[p] = 0.f, la[p] = 0.0, cand_lane[p] = -1;
        if ((slow | dead) >> p & 1u) continue;
        bool in = false;
        if (kk[p] >= 0) {
            pre[p] = __fadd_rn(tmp[p], delta[p][kk[p]]);
            in = ford(__fadd_rn(pre[p], h[p])) >= thr[p];
        }
        const unsigned cands = __ballot_sync(FULL_MASK, in);
        if (__popc(cands) != 1) {
            slow |= 1u << p;
            continue;
        }
        cand_lane[p] = __ffs(cands) - 1;
        if (in) la[p] = __ldg(LAd + (size_t)kk[p] * K + icol[p]);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        if (dead >> p & 1u) {
            out[p] = Best{-FLT_MAX, -1};
        } else if (slow >> p & 1u) {
            out[p] = resolve_column(cm[p], tmp[p], col[p], delta[p], LAd, K, Kp, icol[p], lane);
        } else {
            const float x = exact_cand(pre[p], la[p]);  // meaningful on the candidate lane only
            Best b;
            b.x = __shfl_sync(FULL_MASK, x, cand_lane[p]);
            b.k = __shfl_sync(FULL_MASK, kk[p], cand_lane[p]);
            if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
            out[p] = b;
        }
    }
}

}  // namespace flashv
