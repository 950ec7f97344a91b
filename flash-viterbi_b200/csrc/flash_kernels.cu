// flash_kernels.cu — FLASH passes on the device: start vectors, the per-step max-plus kernel
// (engine STEP), end-state selection and the non-recursive backtrack.
//
// A "pass" advances a set of independent trellis vectors in lock-step: the N-way first pass
// (nvviterNdivide, F:126-202; one vector per sequence) or one level of the task tree (nvviter,
// F:204-262; one vector per task per sequence).  Instead of composing tracker tables every step
// (F:176-179, F:242: (N-1)*K gathers per step), the kernels append each step's backpointer row to
// an HBM store and one thread per vector walks it backwards afterwards; composing T2 forward and
// walking psi backward give the same state by construction (T2_j[i] = T2_{j-1}[psi_j[i]]).
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <stdlib.h>

#include <algorithm>

#include "flashv_internal.h"
#include "half_filter.cuh"
#include "tile_geom.h"
#include "trellis_common.cuh"

namespace flashv {

// ---- start vectors: F:142 / F:212 (pi form) and F:150 / F:220 (restart from Ans[L-1]) ------
__device__ __forceinline__ void init_vector(const VecDesc &vd, int v, const int32_t *__restrict__ ob,
                                            const int32_t *__restrict__ ans, int T, const double *__restrict__ LAd,
                                            const double *__restrict__ LBd, const double *__restrict__ LPi, int K, int Kp,
                                            float *__restrict__ delta, int i_first, int i_stride)
{
    const int prev = vd.L == 0 ? -1 : ans[(size_t)vd.seq * T + vd.L - 1];
    const int o = ob[(size_t)vd.seq * T + vd.L];
    for (int i = i_first; i < K; i += i_stride) {
        const double head = prev < 0 ? LPi[i] : LAd[(size_t)prev * K + i];
        delta[(size_t)v * Kp + i] = __double2float_rn(__dadd_rn(head, LBd[(size_t)o * K + i]));
    }
}

__global__ void k_flash_init(const VecDesc *__restrict__ vecs, int nvec, const int32_t *__restrict__ ob,
                             const int32_t *__restrict__ ans, int T, const double *__restrict__ LAd,
                             const double *__restrict__ LBd, const double *__restrict__ LPi, int K, int Kp,
                             float *__restrict__ delta)
{
    for (int v = blockIdx.y; v < nvec; v += gridDim.y)
        init_vector(vecs[v], v, ob, ans, T, LAd, LBd, LPi, K, Kp, delta, blockIdx.x * blockDim.x + threadIdx.x,
                    gridDim.x * blockDim.x);
}

// Running maxima of the float estimate for an RI x QB tile (RI destination columns, QB delta
// vectors in shared memory) over all source states: lane l owns k = 4*(l+32u)+c.  The hi values
// come straight from L2 (~1 us under load): small tiles hide that with many loads in flight per
// lane (deep unrolling); the big batched tile has no registers to spare for that and runs an
// explicit two-iterations-ahead register prefetch instead.
template <int QB, int RI>
__device__ __forceinline__ void tile_accumulate(float (&cm)[RI][QB][4], const float (&tmp)[RI][QB],
                                                const float4 *const (&col4)[RI], const float4 *sdelta4, int Kp4, int lane)
{
#define FV_ACC(H)                                                                                   \
    _Pragma("unroll") for (int q = 0; q < QB; ++q) {                                                \
        const float4 d = sdelta4[q * Kp4 + t];                                                       \
        _Pragma("unroll") for (int r = 0; r < RI; ++r) {                                            \
            cm[r][q][0] = fmaxf(cm[r][q][0], __fadd_rn(__fadd_rn(tmp[r][q], d.x), H[r].x));          \
            cm[r][q][1] = fmaxf(cm[r][q][1], __fadd_rn(__fadd_rn(tmp[r][q], d.y), H[r].y));          \
            cm[r][q][2] = fmaxf(cm[r][q][2], __fadd_rn(__fadd_rn(tmp[r][q], d.z), H[r].z));          \
            cm[r][q][3] = fmaxf(cm[r][q][3], __fadd_rn(__fadd_rn(tmp[r][q], d.w), H[r].w));          \
        }                                                                                            \
    }
    if (QB * RI >= 8) {
        // three statically named buffers, each re-requested three iterations ahead right after its use: no
        // register rotation (a MOV of a register whose load is still in flight waits for the load — ncu showed
        // the rotation's MOVs stalled on the long scoreboard as often as the FADDs issued)
        float4 hA[RI], hB[RI], hC[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            hA[r] = __ldg(col4[r] + lane);
            hB[r] = lane + 32 < Kp4 ? __ldg(col4[r] + lane + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            hC[r] = lane + 64 < Kp4 ? __ldg(col4[r] + lane + 64) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll 1
        for (int t0 = lane; t0 < Kp4; t0 += 96) {
            {
                const int t = t0;
                FV_ACC(hA)
                if (t + 96 < Kp4) {
#pragma unroll
                    for (int r = 0; r < RI; ++r) hA[r] = __ldg(col4[r] + t + 96);
                }
            }
            if (t0 + 32 < Kp4) {
                const int t = t0 + 32;
                FV_ACC(hB)
                if (t + 96 < Kp4) {
#pragma unroll
                    for (int r = 0; r < RI; ++r) hB[r] = __ldg(col4[r] + t + 96);
                }
            }
            if (t0 + 64 < Kp4) {
                const int t = t0 + 64;
                FV_ACC(hC)
                if (t + 96 < Kp4) {
#pragma unroll
                    for (int r = 0; r < RI; ++r) hC[r] = __ldg(col4[r] + t + 96);
                }
            }
        }
    } else {
        constexpr int UNROLL = QB * RI >= 2 ? 4 : 8;
#pragma unroll UNROLL
        for (int t = lane; t < Kp4; t += 32) {
            float4 h[RI];
#pragma unroll
            for (int r = 0; r < RI; ++r) h[r] = __ldg(col4[r] + t);
            FV_ACC(h)
        }
    }
#undef FV_ACC
}

// ---- engine STEP: one launch per trellis step ------------------------------------------------
// A block of NWARP warps keeps delta of QB vectors in shared memory and walks over tiles of
// NWARP*RI destination columns (grid.y workers per vector group, grid.x vector groups).  A warp
// owns RI columns of the tile: lane l reads hiT[i][4*(l+32u) .. +3] as one 128-bit load per u, so
// a column streams as contiguous 512-byte requests straight from L2/HBM into registers, and every
// hi value meets QB delta vectors, every delta value RI columns (an RI x QB register tile of
// running maxima per lane and float4 component).  Per update: FADD, FADD, FMNMX (estimate only);
// the exact value is recovered by resolve_column() for the handful of candidates inside the
// window.  (QB,RI,NWARP) = (1,1,8) for single vectors, (8,2,16) for the batched tree levels.
struct StepArgs {
    const float *hiT;
    const double *LAd;
    const float *LBf;
    int K, Kp;
    const VecDesc *vecs;
    int nact, s;
    const float *din;
    float *dout;
    const int32_t *ob;
    int T;
    void *psi;
    int psi16;
    int c_begin = 0, c_end = 0;  // destination columns this launch computes: [c_begin, c_end), 0/0 = all K
};

// The body of one step for vector group `group` by column worker `worker` of `nworkers` (standalone: one
// launch per step, grid = groups x workers; persistent level kernel: the same CTA comes back every step).
// delta is read with ld.global.cg: in the persistent kernel it was written by other SMs one grid barrier ago.
template <int QB, int RI, int NWARP>
__device__ __forceinline__ void step_body(const StepArgs &a, int step, int nact, const float *__restrict__ din, float *__restrict__ dout,
                                          int group, int worker, int nworkers, float4 *sdelta4)
{
    constexpr int NT = NWARP * 32;
    const int Kp4 = a.Kp >> 2;
    const int v0 = group * QB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // staging with cp.async (global -> shared without a register in between, L2 only): all QB vectors are
    // requested at once and the CTA pays ONE memory round trip — the load/store loop this replaces paid one per
    // vector (ncu: 22 % of an 8-vector step, all of it long-scoreboard stalls on the stores)
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        const bool live = v0 + q < nact;
        const float4 *src = reinterpret_cast<const float4 *>(din + (size_t)(v0 + q) * a.Kp);
        for (int t = tid; t < Kp4; t += NT) {
            if (live)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdelta4 + q * Kp4 + t)),
                             "l"(src + t)
                             : "memory");
            else
                sdelta4[q * Kp4 + t] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float *sdelta = reinterpret_cast<const float *>(sdelta4);

    int jj[QB];
    const float *tmp_row[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        jj[q] = 0, tmp_row[q] = a.LBf;
        if (v0 + q < nact) {
            const VecDesc vd = a.vecs[v0 + q];
            jj[q] = vd.L + step;
            tmp_row[q] = a.LBf + (size_t)a.ob[(size_t)vd.seq * a.T + jj[q]] * a.Kp;  // F:167
        }
    }

    const int c_begin = a.c_end > 0 ? a.c_begin : 0, c_end = a.c_end > 0 ? a.c_end : a.K;
    const int ntiles = (c_end - c_begin + NWARP * RI - 1) / (NWARP * RI);
    for (int tile = worker; tile < ntiles; tile += nworkers) {
        const int ibase = c_begin + (tile * NWARP + warp) * RI;
        if (ibase >= c_end) continue;  // warp-uniform; no block-wide barrier below
        int col_i[RI];
        const float4 *col4[RI];
        float tmp[RI][QB];
        float cm[RI][QB][4];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            col_i[r] = min(ibase + r, c_end - 1);  // a clamped duplicate column is computed and dropped
            col4[r] = reinterpret_cast<const float4 *>(a.hiT + (size_t)col_i[r] * a.Kp);
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                tmp[r][q] = __ldg(tmp_row[q] + col_i[r]);
                cm[r][q][0] = cm[r][q][1] = cm[r][q][2] = cm[r][q][3] = -INFINITY;
            }
        }
        tile_accumulate<QB, RI>(cm, tmp, col4, sdelta4, Kp4, lane);
        // exact (value, first index) of the pairs, one column (QB pairs) at a time so that the
        // in-flight state of resolve_tile stays in registers; within a column the memory round
        // trips of all QB pairs overlap
#pragma unroll
        for (int r = 0; r < RI; ++r) {
            if (ibase + r >= c_end) continue;  // warp-uniform: the clamped duplicate column
            const int i = ibase + r;
            const float *pcol[QB];
            const float *pdelta[QB];
            int picol[QB];
            Best res[QB];
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                picol[q] = i;
                pcol[q] = a.hiT + (size_t)i * a.Kp;
                pdelta[q] = sdelta + (size_t)q * a.Kp;
            }
            resolve_tile<QB>(cm[r], tmp[r], pcol, pdelta, picol, a.LAd, a.K, a.Kp, lane, res);
#pragma unroll
            for (int q = 0; q < QB; ++q) {
                if (v0 + q >= nact) continue;  // warp-uniform: padding vector of the last group
                if (lane == 0) {
                    dout[(size_t)(v0 + q) * a.Kp + i] = res[q].x;
                    const VecDesc vd = a.vecs[v0 + q];
                    if (jj[q] >= vd.mid + 1)  // F:242: only steps from the latch on are ever read back
                        psi_store(a.psi, a.psi16, (size_t)(vd.psi_row + (jj[q] - vd.mid - 1)) * a.K + i, res[q].k);
                }
            }
        }
    }
}

template <int QB, int RI, int NWARP>
__global__ void __launch_bounds__(NWARP * 32) k_flash_step(const StepArgs a)
{
    extern __shared__ float4 sdelta4[];
    step_body<QB, RI, NWARP>(a, a.s, a.nact, a.din, a.dout, blockIdx.x, blockIdx.y, gridDim.y, sdelta4);  // vector groups on x (no 65535 cap), column workers on y
}

// ---- the last step of a task needs one column ---------------------------------------------------
// nvviter() ends with Ans[mid] = T2[cur][Ans[R]] (F:248, F:261): of the K values the last step
// computes, only the backpointer of destination state Ans[R] is ever read (delta of the last step
// is only needed by full-range passes, F:249-259).  So a task's last step is K updates, not K^2 —
// and half of all tasks are one step long.  One CTA of LC_WARPS warps per vector: the column and the
// vector are two 16 KB streams and the kernel is a chain of L2 round trips, so the more lanes share
// them the shorter it gets.  Warp w takes the float4 groups t = 32*(w + LC_WARPS*m) + lane, i.e. the
// elements u = w, w + LC_WARPS, ... of every chain (lane, component) of trellis_common.cuh; same
// estimate / window / exact logic as every other kernel, combined across warps in shared memory.
constexpr int LC_WARPS = 8;
// Executed by a whole CTA of at least LC_WARPS warps (the first LC_WARPS work, all take the barriers).
__device__ __forceinline__ void last_column_body(const StepArgs &a, int step, const float *__restrict__ din, int v,
                                                 const int32_t *__restrict__ ans)
{
    __shared__ float s_top[LC_WARPS];
    __shared__ float s_bx[LC_WARPS];
    __shared__ int s_bk[LC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const VecDesc vd = a.vecs[v];
    const int j = vd.L + step;  // == vd.R
    const int e = ans[(size_t)vd.seq * a.T + vd.R];
    if (e < 0 || e >= a.K) return;  // CTA-uniform
    const float tmp = __ldg(a.LBf + (size_t)a.ob[(size_t)vd.seq * a.T + j] * a.Kp + e);  // F:233
    const float *col = a.hiT + (size_t)e * a.Kp;
    const float *delta = din + (size_t)v * a.Kp;
    const float4 *col4 = reinterpret_cast<const float4 *>(col);
    const float4 *d4 = reinterpret_cast<const float4 *>(delta);
    const int Kp4 = a.Kp >> 2;
    float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
    for (int t = 32 * warp + lane; t < Kp4 && warp < LC_WARPS; t += 32 * LC_WARPS) {
        const float4 h = __ldg(col4 + t);
        const float4 d = __ldcg(d4 + t);
        cm[0] = fmaxf(cm[0], __fadd_rn(__fadd_rn(tmp, d.x), h.x));
        cm[1] = fmaxf(cm[1], __fadd_rn(__fadd_rn(tmp, d.y), h.y));
        cm[2] = fmaxf(cm[2], __fadd_rn(__fadd_rn(tmp, d.z), h.z));
        cm[3] = fmaxf(cm[3], __fadd_rn(__fadd_rn(tmp, d.w), h.w));
    }
    const float wtop = warp_max(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
    if (lane == 0 && warp < LC_WARPS) s_top[warp] = wtop;
    __syncthreads();
    float top = s_top[0];
#pragma unroll
    for (int w = 1; w < LC_WARPS; ++w) top = fmaxf(top, s_top[w]);
    Best b{-FLT_MAX, 0x7fffffff};
    if (top > -FLT_MAX) {
        const int thr = ford(top) - WINDOW_STEPS;
        // every lane re-reads its own elements of the chains whose maximum is inside the window
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (ford(cm[c]) < thr || warp >= LC_WARPS) continue;
            for (int t = 32 * warp + lane; t < Kp4; t += 32 * LC_WARPS) {
                const int k = 4 * t + c;
                if (k >= a.K) break;
                const float pre = __fadd_rn(tmp, __ldcg(delta + k));
                if (ford(__fadd_rn(pre, __ldg(col + k))) >= thr) {
                    const float x = exact_cand(pre, __ldg(a.LAd + (size_t)k * a.K + e));
                    if (x > -FLT_MAX) best_take(b, x, k);
                }
            }
        }
    }
    b = warp_best(b);
    if (lane == 0 && warp < LC_WARPS) s_bx[warp] = b.x, s_bk[warp] = b.k;
    __syncthreads();
    if (threadIdx.x == 0) {
        Best r{s_bx[0], s_bk[0]};
#pragma unroll
        for (int w = 1; w < LC_WARPS; ++w) best_take(r, s_bx[w], s_bk[w]);
        if (!(r.x > -FLT_MAX)) r.k = -1;
        psi_store(a.psi, a.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * a.K + e, r.k);
    }
    __syncthreads();  // the shared slots are reused by the CTA's next vector
}

__global__ void __launch_bounds__(LC_WARPS * 32) k_flash_last_column(const StepArgs a, int v_begin, const int32_t *__restrict__ ans)
{
    last_column_body(a, a.s, a.din, v_begin + blockIdx.x, ans);
}

// ---- end of a full-range pass: Ans[T-1] = first argmax of delta (F:188-195, F:251-258) ----
// Executed by a whole CTA (any size up to 32 warps).
__device__ __forceinline__ void end_body(const VecDesc &vd, int v, const float *__restrict__ delta, int K, int Kp, int T,
                                         int32_t *__restrict__ ans, float *__restrict__ score, int32_t *__restrict__ endstate)
{
    __shared__ float sx[32];
    __shared__ int sk[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = (blockDim.x + 31) >> 5;
    if (vd.flags & VEC_FULL_RANGE) {
        // "score = T1[cur][0]; arg = 0; if (T1[cur][i] > score)": first maximum, index 0 if all equal.
        Best b{-INFINITY, 0x7fffffff};
        for (int i = tid; i < K; i += blockDim.x) {
            float x = __ldcg(delta + (size_t)v * Kp + i);
            if (x > b.x || (x == b.x && i < b.k)) b.x = x, b.k = i;
        }
        b = warp_best(b);
        if (lane == 0) sx[warp] = b.x, sk[warp] = b.k;
        __syncthreads();
        if (tid == 0) {
            Best r{sx[0], sk[0]};
            for (int w = 1; w < nwarp; ++w) best_take(r, sx[w], sk[w]);
            if (r.k == 0x7fffffff) r.k = 0;  // every entry -inf: the reference keeps arg = 0
            ans[(size_t)vd.seq * T + vd.R] = r.k;
            score[vd.seq] = __ldcg(delta + (size_t)v * Kp + r.k);
            endstate[v] = r.k;
        }
        __syncthreads();
    } else if (tid == 0) {
        endstate[v] = ans[(size_t)vd.seq * T + vd.R];  // F:248
    }
}

__global__ void __launch_bounds__(256) k_flash_end(const VecDesc *__restrict__ vecs, int nvec,
                                                   const float *__restrict__ delta, int K, int Kp, int T,
                                                   int32_t *__restrict__ ans, float *__restrict__ score,
                                                   int32_t *__restrict__ endstate)
{
    const int v = blockIdx.x;
    if (v >= nvec) return;
    end_body(vecs[v], v, delta, K, Kp, T, ans, score, endstate);
}

// ---- non-recursive backtrack through the stored rows --------------------------------------
// One thread per vector: state_{j-1} = psi_j[state_j] for j = R .. mid+1.  A task records
// Ans[mid] (F:261); the first pass records every segment boundary it walks over (F:198-201).
// (rows are read with ld.global.cg: in the persistent level kernel other SMs wrote them a grid barrier ago)
__device__ __forceinline__ int psi_load_cg(const void *base, int psi16, size_t idx)
{
    if (psi16) {
        const unsigned v = __ldcg(reinterpret_cast<const uint16_t *>(base) + idx);
        return v == 0xFFFFu ? -1 : (int)v;
    }
    return __ldcg(reinterpret_cast<const int32_t *>(base) + idx);
}

__device__ __forceinline__ void backtrack_vector(const VecDesc &vd, int state, const void *__restrict__ psi, int psi16, int K, int T,
                                                 const uint8_t *__restrict__ ismid, int32_t *__restrict__ ans)
{
    int32_t *out = ans + (size_t)vd.seq * T;
    for (int j = vd.R; j >= vd.mid + 1; --j) {
        if (state >= 0) state = psi_load_cg(psi, psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + state);
        if ((vd.flags & VEC_FIRST_PASS) && ismid[j - 1]) out[j - 1] = state;
    }
    if (!(vd.flags & VEC_FIRST_PASS)) out[vd.mid] = state;
}

__global__ void k_flash_backtrack(const VecDesc *__restrict__ vecs, int nvec, const void *__restrict__ psi, int psi16,
                                  int K, int T, const uint8_t *__restrict__ ismid,
                                  const int32_t *__restrict__ endstate, int32_t *__restrict__ ans)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvec) return;
    backtrack_vector(vecs[v], endstate[v], psi, psi16, K, T, ismid, ans);
}

// The same walk for ONE long vector (the first pass: T-1-mids[0] dependent hops).  A hop through
// global memory costs an L2 round trip (~0.5 us: 125 us for T=256); here the CTA copies a window
// of consecutive rows into shared memory with coalesced 16-byte loads and thread 0 hops inside it.
__global__ void __launch_bounds__(1024) k_flash_backtrack_staged(const VecDesc *__restrict__ vecs, const void *__restrict__ psi,
                                                                int psi16, int K, int T, const uint8_t *__restrict__ ismid,
                                                                const int32_t *__restrict__ endstate,
                                                                int32_t *__restrict__ ans, int win_rows)
{
    extern __shared__ uint4 swin[];
    __shared__ int s_state;
    const VecDesc vd = vecs[0];
    const size_t esz = psi16 ? 2 : 4;
    const size_t row_bytes = (size_t)K * esz;
    int32_t *out = ans + (size_t)vd.seq * T;
    if (threadIdx.x == 0) s_state = endstate[0];
    for (int jhi = vd.R; jhi >= vd.mid + 1; jhi -= win_rows) {
        const int jlo = max(jhi - win_rows + 1, vd.mid + 1);
        const size_t lo = (size_t)(vd.psi_row + (jlo - vd.mid - 1)) * row_bytes;      // first byte of row jlo
        const size_t hi = (size_t)(vd.psi_row + (jhi - vd.mid - 1) + 1) * row_bytes;  // one past row jhi
        const size_t lo16 = lo & ~(size_t)15;
        const int n16 = (int)((hi - lo16 + 15) >> 4);
        const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(psi) + lo16);
        __syncthreads();  // previous window fully consumed
        // 8 x 16 bytes in flight per thread: the window copy is a string of L2 round trips otherwise
        // (reads past `hi` stay inside the store's padding)
        for (int t0 = threadIdx.x; t0 < n16; t0 += 8 * blockDim.x) {
            uint4 buf[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u * blockDim.x < n16) buf[u] = src[t0 + u * blockDim.x];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u * blockDim.x < n16) swin[t0 + u * blockDim.x] = buf[u];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned char *base = reinterpret_cast<const unsigned char *>(swin) + (lo - lo16);
            int state = s_state;
            for (int j = jhi; j >= jlo; --j) {
                if (state >= 0) state = psi_load(base, psi16, (size_t)(j - jlo) * K + state);
                if ((vd.flags & VEC_FIRST_PASS) && ismid[j - 1]) out[j - 1] = state;
            }
            s_state = state;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !(vd.flags & VEC_FIRST_PASS)) out[vd.mid] = s_state;
}

// ---- one level of the task tree in ONE launch ---------------------------------------------------------
// The reference's worker pool (F:264-308) runs the tasks of the tree on MAX_THREADS host threads; here one
// cooperative kernel per level keeps every SM on the level's tasks from their start vectors to their
// midpoints: start vectors (F:212/F:220), every trellis step of every task in lock-step (grid barrier
// between steps; the step itself is step_body, the tasks on their last step take last_column_body), the
// end states (F:248-259) and the walk back to Ans[mid] (F:261).  One launch instead of 3 + 2 per step.
// Grid barrier of the level kernel: one monotone 64-bit counter per plan.  Every CTA adds 1 and waits until
// the counter reaches (barriers so far) x (CTAs); the host tells each launch where the count stands
// (bar_base), so nothing is reset and a CTA that starts late cannot misread an earlier generation.  All
// CTAs are co-resident (cooperative launch).  The grid-wide sync of cooperative_groups did the same job in
// the first version and cost more per step than a kernel launch.
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1ull);
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(bar) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

// The walk back of the long first pass, in parallel: window w of consecutive backpointer rows goes to CTA w (shared
// memory, as in k_flash_backtrack_staged), which first composes its rows into ONE map — the state below the window for
// every state above it, K independent walks at shared-memory latency — then CTA 0 chains the W maps from the end
// state (W dependent L2 reads instead of T), and every CTA walks its own window once more from the state it was
// handed, recording the division points (F:196-201).  46 us -> 15 us for T=256 at K=3965.
__global__ void __launch_bounds__(1024) k_flash_backtrack_par(const VecDesc *__restrict__ vecs, const void *__restrict__ psi, int psi16,
                                                             int K, int T, const uint8_t *__restrict__ ismid,
                                                             const int32_t *__restrict__ endstate, int32_t *__restrict__ ans,
                                                             int win_rows, int32_t *__restrict__ maps, unsigned long long *bar,
                                                             unsigned long long bar_base, const float *__restrict__ final_delta, int Kp,
                                                             float *__restrict__ score, int32_t *__restrict__ endstate_out)
{
    extern __shared__ uint4 swin[];
    const VecDesc vd = vecs[0];
    const int w = blockIdx.x, W = gridDim.x;
    // the end state (F:188-195 / F:248) is found here too, by the CTA with the (shorter) last window, while the others compose
    if (w == W - 1) end_body(vd, 0, final_delta, K, Kp, T, ans, score, endstate_out);
    int32_t *starts = maps + (size_t)W * K;
    const size_t esz = psi16 ? 2 : 4;
    const size_t row_bytes = (size_t)K * esz;
    int32_t *out = ans + (size_t)vd.seq * T;
    const int jhi = vd.R - w * win_rows, jlo = max(jhi - win_rows + 1, vd.mid + 1);  // the host sizes the grid so that jhi >= jlo
    const size_t lo = (size_t)(vd.psi_row + (jlo - vd.mid - 1)) * row_bytes;
    const size_t hi = (size_t)(vd.psi_row + (jhi - vd.mid - 1) + 1) * row_bytes;
    const size_t lo16 = lo & ~(size_t)15;
    const int n16 = (int)((hi - lo16 + 15) >> 4);
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(psi) + lo16);
    for (int t0 = threadIdx.x; t0 < n16; t0 += 8 * blockDim.x) {
        uint4 buf[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u * blockDim.x < n16) buf[u] = __ldcg(src + t0 + u * blockDim.x);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u * blockDim.x < n16) swin[t0 + u * blockDim.x] = buf[u];
    }
    __syncthreads();
    const unsigned char *base = reinterpret_cast<const unsigned char *>(swin) + (lo - lo16);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        int state = i;
        for (int j = jhi; j >= jlo; --j)
            if (state >= 0) state = psi_load(base, psi16, (size_t)(j - jlo) * K + state);
        maps[(size_t)w * K + i] = state;
    }
    grid_barrier(bar, bar_base + W);
    if (w == 0 && threadIdx.x == 0) {
        int state = __ldcg(endstate_out);  // written by CTA W-1 before the barrier
        for (int x = 0; x < W; ++x) {
            starts[x] = state;
            if (state >= 0) state = __ldcg(maps + (size_t)x * K + state);
        }
        if (!(vd.flags & VEC_FIRST_PASS)) out[vd.mid] = state;
    }
    grid_barrier(bar, bar_base + 2ull * W);
    if (threadIdx.x == 0 && (vd.flags & VEC_FIRST_PASS)) {
        int state = __ldcg(starts + w);
        for (int j = jhi; j >= jlo; --j) {
            if (state >= 0) state = psi_load(base, psi16, (size_t)(j - jlo) * K + state);
            if (ismid[j - 1]) out[j - 1] = state;
        }
    }
}

struct LevelArgs {
    StepArgs st;           // tables, vectors, observations, backpointer store (s, nact, din, dout are set per step)
    const double *LBd, *LPi;
    int nvec, max_steps, full_range;
    const int *nactive;    // [max_steps + 2]: vectors still stepping at step s (a prefix of the order)
    float *d0, *d1;        // delta ping-pong, [nvec][Kp] each
    int32_t *ans;
    float *score;
    int32_t *endstate;
    const uint8_t *ismid;
    unsigned long long *bar;      // the plan's barrier counter
    unsigned long long bar_base;  // its value when this launch starts
};

template <int QB, int RI, int NWARP>
__global__ void __launch_bounds__(NWARP * 32) k_flash_level(const LevelArgs la)
{
    extern __shared__ float4 sdelta4[];
    unsigned long long bar_target = la.bar_base;
    const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
    const int K = la.st.K, Kp = la.st.Kp, T = la.st.T;
    for (int v = b; v < la.nvec; v += G)
        init_vector(la.st.vecs[v], v, la.st.ob, la.ans, T, la.st.LAd, la.LBd, la.LPi, K, Kp, la.d0, tid, NWARP * 32);
    grid_barrier(la.bar, bar_target += gridDim.x);
    for (int s = 1; s <= la.max_steps; ++s) {
        // vectors are sorted longest first: [0, n_cont) go on after this step, [n_cont, n_act) are on
        // their last step and (unless the pass is full-range) need a single column
        const int n_act = la.nactive[s], n_cont = la.full_range ? n_act : la.nactive[s + 1];
        const float *din = (s & 1) ? la.d0 : la.d1;
        float *dout = (s & 1) ? la.d1 : la.d0;
        if (n_cont > 0) {
            const int ngroups = (n_cont + QB - 1) / QB;
            const int workers = G / ngroups > 0 ? G / ngroups : 1;
            for (int item = b; item < ngroups * workers; item += G) {  // one item per CTA unless there are more groups than CTAs
                step_body<QB, RI, NWARP>(la.st, s, n_cont, din, dout, item / workers, item % workers, workers, sdelta4);
                __syncthreads();  // the staged deltas are replaced by the next item's
            }
        }
        for (int v = n_cont + b; v < n_act; v += G) last_column_body(la.st, s, din, v, la.ans);
        grid_barrier(la.bar, bar_target += gridDim.x);
    }
    const float *final_delta = (la.max_steps & 1) ? la.d1 : la.d0;
    for (int v = b; v < la.nvec; v += G) {
        const VecDesc vd = la.st.vecs[v];
        end_body(vd, v, final_delta, K, Kp, T, la.ans, la.score, la.endstate);
        if (tid == 0) backtrack_vector(vd, la.endstate[v], la.st.psi, la.st.psi16, K, T, la.ismid, la.ans);
        __syncthreads();
    }
}

template <int QB, int RI, int NWARP>
static int launch_level(flashv_plan *p, const Pass &pass, LevelArgs &la, bool *done)
{
    flashv_ctx *ctx = p->model->ctx;
    const size_t smem = (size_t)QB * la.st.Kp * sizeof(float);
    const void *fn = (const void *)k_flash_level<QB, RI, NWARP>;
    FV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int per_sm = 0;
    FV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NWARP * 32, smem));
    if (per_sm < 1) return FLASHV_OK;  // does not fit: the caller falls back to one launch per step
    if (per_sm > 4) per_sm = 4;
    const int ngroups = (pass.nvec + QB - 1) / QB;
    const int ntiles = (la.st.K + NWARP * RI - 1) / (NWARP * RI);
    int grid = per_sm * ctx->sm_count;
    if (grid > ngroups * ntiles) grid = ngroups * ntiles;  // no more CTAs than (group, tile) items
    if (grid < 1) grid = 1;
    la.bar = reinterpret_cast<unsigned long long *>(p->d_sync), la.bar_base = p->bar_count;
    p->bar_count += (unsigned long long)(1 + pass.max_steps) * grid;  // one barrier after the start vectors, one per step
    void *params[] = {(void *)&la};
    FV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NWARP * 32), params, smem, ctx->stream));
    ++p->launches;
    *done = true;
    return FLASHV_OK;
}


// ---- the same level in ONE launch with the half-precision filter (models up to 4096 states) ------------------
// k_flash_level spreads (vector group, column tile) items over the CTAs and streams the float table from L2 for
// every item.  Here the roles are those of the pass kernel k_flash_persist16 (flash_persistent.cu; the filter and
// its window: DESIGN.md §4): CTA b owns the destination columns [tile_c0(b), tile_c0(b+1)) for the whole launch,
// its slice of (half)log A is parked in tensor memory once, and every step it walks ALL vectors of the level over
// those columns, VG vectors per trip: one tcgen05.ld feeds VG vectors (HADD2 + HMNMX2 per two updates), the
// winners come from the chain-major double table, delta ping-pongs through global memory as plain floats with a
// grid barrier per step, exactly as in k_flash_level.  Start vectors, last-column steps, end states and the walk
// back are the shared bodies above.
constexpr int L16_WARPS = 14, L16_THREADS = L16_WARPS * 32, L16_PAIRS = 2;

template <int VG>
__global__ void __launch_bounds__(L16_THREADS, 1) k_flash_level16(const LevelArgs la, const __half *__restrict__ hi16,
                                                                    const double *__restrict__ LAc16, int Kp16)
{
    extern __shared__ __align__(16) unsigned char smem16[];
    __shared__ uint32_t tmem_slot;
    __shared__ int wmax[VG][L16_WARPS];
    const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = la.st.K, Kp = la.st.Kp, T = la.st.T;
    float *sdelta = reinterpret_cast<float *>(smem16);                       // [VG][Kp] floats, for the exact evaluation
    __half *sdelta16 = reinterpret_cast<__half *>(sdelta + (size_t)VG * Kp);  // [VG][Kp16] fl16(delta - max), for the sweep
    unsigned long long bar_target = la.bar_base;

    const int c0 = tile_c0(K, G, b), ncols = tile_c0(K, G, b + 1) - c0;  // at most 28: two per warp
    const int n_it = Kp16 >> 8;
    const __half *slab = hi16 + (size_t)c0 * Kp16;  // [iteration][column][256]
    if (warp == 0) tmem_alloc(&tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2);
    const int rr0 = 2 * warp, rr1 = rr0 + 1;
    const bool have0 = rr0 < ncols, have1 = rr1 < ncols;
    const int rs0 = have0 ? rr0 : 0, rs1 = have1 ? rr1 : rs0;  // a duplicate row stands in for a missing one
    const int i0 = c0 + rs0, i1 = c0 + rs1;
    if (have0) tmem_fill16(tbase, slab, ncols, rs0, rs1, n_it, lane);
    tmem_wait_st();
    tmem_fence_before();

    for (int v = b; v < la.nvec; v += G)
        init_vector(la.st.vecs[v], v, la.st.ob, la.ans, T, la.st.LAd, la.LBd, la.LPi, K, Kp, la.d0, tid, L16_THREADS);
    grid_barrier(la.bar, bar_target += gridDim.x);
    tmem_fence_after();

    constexpr int NB = 5;  // pairs of delta values per thread and vector: 4096 / 2 / 448 rounded up
    const int Kp2 = Kp >> 1, Kh2 = Kp16 >> 1;
    for (int s = 1; s <= la.max_steps; ++s) {
        const int n_act = la.nactive[s], n_cont = la.full_range ? n_act : la.nactive[s + 1];
        const float *din = (s & 1) ? la.d0 : la.d1;
        float *dout = (s & 1) ? la.d1 : la.d0;
        for (int v0 = 0; v0 < n_cont; v0 += VG) {
            const int nv = min(VG, n_cont - v0);
            // ---- stage VG vectors: floats, their maxima, then the half-precision images ----
            float2 val[VG][NB];
#pragma unroll
            for (int q = 0; q < VG; ++q)
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int t = tid + e * L16_THREADS;
                    val[q][e] = (q < nv && t < Kp2) ? __ldcg(reinterpret_cast<const float2 *>(din + (size_t)(v0 + q) * Kp) + t)
                                                    : make_float2(0.f, 0.f);
                }
#pragma unroll
            for (int q = 0; q < VG; ++q) {
                float mx = -INFINITY;
                float2 *sd2 = reinterpret_cast<float2 *>(sdelta + (size_t)q * Kp);
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int t = tid + e * L16_THREADS, k = 2 * t;
                    if (k >= K) val[q][e].x = 0.f;  // padding stays finite (the table pads with -inf)
                    else mx = fmaxf(mx, val[q][e].x);
                    if (k + 1 >= K) val[q][e].y = 0.f;
                    else mx = fmaxf(mx, val[q][e].y);
                    if (t < Kp2) sd2[t] = val[q][e];
                }
                const int wm = __reduce_max_sync(FULL_MASK, ford(mx));
                if (lane == 0) wmax[q][warp] = wm;
            }
            __syncthreads();
            float cmax[VG];
#pragma unroll
            for (int q = 0; q < VG; ++q) {
                cmax[q] = unford(__reduce_max_sync(FULL_MASK, lane < L16_WARPS ? wmax[q][lane] : ford(-INFINITY)));
                __half2 *sh2 = reinterpret_cast<__half2 *>(sdelta16 + (size_t)q * Kp16);
                const float c = cmax[q] > -FLT_MAX ? cmax[q] : 0.f;  // no state alive: the vector is skipped below
#pragma unroll
                for (int e = 0; e < NB; ++e) {
                    const int t = tid + e * L16_THREADS, k = 2 * t;
                    if (t < Kh2) {
                        const float lo = k < K ? fmaxf(__fsub_rn(val[q][e].x, c), H_CLAMP) : 0.f;
                        const float hi = k + 1 < K ? fmaxf(__fsub_rn(val[q][e].y, c), H_CLAMP) : 0.f;
                        sh2[t] = __floats2half2_rn(lo, hi);
                    }
                }
            }
            __syncthreads();

            if (have0) {
                // ---- sweep: every tensor-memory load meets VG vectors ----
                const __half2 ninf = __float2half2_rn(-INFINITY);
                __half2 m[VG][2][4];
#pragma unroll
                for (int q = 0; q < VG; ++q)
#pragma unroll
                    for (int w = 0; w < 4; ++w) m[q][0][w] = m[q][1][w] = ninf;
                const uint4 *d4 = reinterpret_cast<const uint4 *>(sdelta16) + lane;
                const int q4 = Kp16 >> 3;  // uint4 per staged vector
#pragma unroll 1
                for (int u = 0; u < n_it; u += 4) {
                    uint32_t r[32];
                    tmem_ld32_raw(tbase + 8u * (uint32_t)u, r);
                    tmem_wait_ld();
                    const int ne = min(4, n_it - u);  // the table pads to whole iterations; a short tail reads unused columns
#pragma unroll
                    for (int q = 0; q < VG; ++q) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (e < ne) {
                                const uint4 d = d4[(size_t)q * q4 + (u + e) * 32];
#pragma unroll
                                for (int w = 0; w < 4; ++w) {
                                    const uint32_t dw = w == 0 ? d.x : w == 1 ? d.y : w == 2 ? d.z : d.w;
                                    m[q][0][w] = __hmax2(m[q][0][w], __hadd2(u2h(dw), u2h(r[8 * e + w])));
                                    m[q][1][w] = __hmax2(m[q][1][w], __hadd2(u2h(dw), u2h(r[8 * e + 4 + w])));
                                }
                            }
                        }
                    }
                }
                // ---- winners, two vectors (four columns) at a time so that their HBM round trips overlap ----
#pragma unroll
                for (int q = 0; q < VG; q += 2) {
                    Scan16T<L16_PAIRS> sc[2][2];
                    float tmp[2][2];
                    bool on[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        on[h] = q + h < nv && cmax[q + h < VG ? q + h : 0] > -FLT_MAX;
                        const VecDesc &vd = la.st.vecs[v0 + (q + h < nv ? q + h : 0)];
                        const float *tmp_row = la.st.LBf + (size_t)la.st.ob[(size_t)vd.seq * T + vd.L + s] * Kp;
                        tmp[h][0] = __ldg(tmp_row + i0), tmp[h][1] = __ldg(tmp_row + i1);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (q + h >= VG) continue;
                        const float c = cmax[q + h];
                        scan16_fetch(sc[h][0], m[q + h][0], tmp[h][0], c, LAc16, i0, on[h], lane);
                        scan16_fetch(sc[h][1], m[q + h][1], tmp[h][1], c, LAc16, i1, on[h] && have1, lane);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (q + h >= VG || q + h >= nv) continue;  // warp-uniform
                        const int v = v0 + q + h;
                        const VecDesc &vd = la.st.vecs[v];
                        const int j = vd.L + s;
                        const float *sd = sdelta + (size_t)(q + h) * Kp;
                        const Best r0 = scan16_settle(sc[h][0], m[q + h][0], tmp[h][0], sd, LAc16, K, i0, lane);
                        const Best r1 = scan16_settle(sc[h][1], m[q + h][1], tmp[h][1], sd, LAc16, K, i1, lane);
                        if (lane == 0) {
                            dout[(size_t)v * Kp + i0] = r0.x;
                            if (j >= vd.mid + 1) psi_store(la.st.psi, la.st.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + i0, r0.k);
                            if (have1) {
                                dout[(size_t)v * Kp + i1] = r1.x;
                                if (j >= vd.mid + 1) psi_store(la.st.psi, la.st.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + i1, r1.k);
                            }
                        }
                    }
                }
            }
            __syncthreads();  // the staged vectors are replaced by the next trip's
        }
        for (int v = n_cont + b; v < n_act; v += G) last_column_body(la.st, s, din, v, la.ans);
        grid_barrier(la.bar, bar_target += gridDim.x);
    }
    const float *final_delta = (la.max_steps & 1) ? la.d1 : la.d0;
    for (int v = b; v < la.nvec; v += G) {
        const VecDesc vd = la.st.vecs[v];
        end_body(vd, v, final_delta, K, Kp, T, la.ans, la.score, la.endstate);
        if (tid == 0) backtrack_vector(vd, la.endstate[v], la.st.psi, la.st.psi16, K, T, la.ismid, la.ans);
        __syncthreads();
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot);
}

template <int VG>
static int launch_level16(flashv_plan *p, const Pass &pass, LevelArgs &la, bool *done)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    const size_t smem = (size_t)VG * ((size_t)m->Kp * 4 + (size_t)m->Kp16 * 2);
    const void *fn = (const void *)k_flash_level16<VG>;
    FV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, L16_THREADS, smem));
    const int grid = m->tile_G;  // the grid the half-precision table was tiled for
    if (per_sm < 1 || grid > ctx->sm_count) return FLASHV_OK;  // does not fit: the caller falls back
    la.bar = reinterpret_cast<unsigned long long *>(p->d_sync), la.bar_base = p->bar_count;
    p->bar_count += (unsigned long long)(1 + pass.max_steps) * grid;
    const __half *hi16 = m->hi16;
    const double *LAc16 = m->LAc16;
    int Kp16 = m->Kp16;
    void *params[] = {(void *)&la, (void *)&hi16, (void *)&LAc16, (void *)&Kp16};
    FV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(L16_THREADS), params, smem, ctx->stream));
    ++p->launches;
    *done = true;
    return FLASHV_OK;
}

// A level whose tasks are all ONE step long (the bottom of every tree): that step is the task's last one, so each
// task needs one column (last_column_body) and nothing from the other tasks — start vector, column, end state and
// Ans[mid] in one ordinary launch, one CTA per task, instead of four launches.
__global__ void __launch_bounds__(LC_WARPS * 32) k_flash_level_one_step(const LevelArgs la)
{
    const int v = blockIdx.x;
    const VecDesc vd = la.st.vecs[v];
    init_vector(vd, v, la.st.ob, la.ans, la.st.T, la.st.LAd, la.LBd, la.LPi, la.st.K, la.st.Kp, la.d0, threadIdx.x, LC_WARPS * 32);
    __threadfence_block();
    __syncthreads();  // the column sweep reads the vector through L2 (ld.global.cg)
    last_column_body(la.st, 1, la.d0, v, la.ans);
    end_body(vd, v, la.d0, la.st.K, la.st.Kp, la.st.T, la.ans, la.score, la.endstate);
    if (threadIdx.x == 0) backtrack_vector(vd, la.endstate[v], la.st.psi, la.st.psi16, la.st.K, la.st.T, la.ismid, la.ans);
}

// Returns with *done = true when the whole pass (start vectors .. Ans[mid]) ran in one launch.
static int level_pass(flashv_plan *p, const Pass &pass, bool *done)
{
    flashv_model *m = p->model;
    *done = false;
    if (!m->ctx->coop || !p->d_nactive || getenv("FLASHV_LEVEL_STEPS")) return FLASHV_OK;
    if (pass.max_steps == 1 && !pass.full_range && pass.nvec <= 65535) {
        LevelArgs la;
        StepArgs &a = la.st;
        a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
        a.vecs = p->d_vecs + pass.vec_offset, a.nact = pass.nvec, a.s = 1, a.din = nullptr, a.dout = nullptr;
        a.ob = p->d_ob, a.T = p->T, a.psi = p->d_psi, a.psi16 = p->psi16;
        la.LBd = m->LBd, la.LPi = m->LPi;
        la.nvec = pass.nvec, la.max_steps = 1, la.full_range = 0;
        la.nactive = nullptr;
        la.d0 = p->d_delta, la.d1 = nullptr;
        la.ans = p->d_ans, la.score = p->d_score, la.endstate = p->d_endstate, la.ismid = p->d_ismid;
        la.bar = nullptr, la.bar_base = 0;
        k_flash_level_one_step<<<pass.nvec, LC_WARPS * 32, 0, m->ctx->stream>>>(la);
        FV_CUDA(cudaGetLastError());
        ++p->launches;
        *done = true;
        return FLASHV_OK;
    }
    // Measured (K=3965, T=256, same box): one launch per level against one per step — N=1 9.65 / 10.12 ms, N=8
    // 6.29 / 6.53 ms, N=64 2.77 / 2.75 ms, N=127 2.22 / 2.16 ms.  A level of one or two steps has nothing to
    // amortise the extra grid barriers over, so short levels keep the per-step launches.
    const int min_steps = getenv("FLASHV_LEVEL_MIN_STEPS") ? atoi(getenv("FLASHV_LEVEL_MIN_STEPS")) : (m->hi16 ? 2 : 5);
    if (pass.max_steps < min_steps) return FLASHV_OK;
    LevelArgs la;
    StepArgs &a = la.st;
    a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.vecs = p->d_vecs + pass.vec_offset, a.nact = 0, a.s = 0, a.din = nullptr, a.dout = nullptr;
    a.ob = p->d_ob, a.T = p->T, a.psi = p->d_psi, a.psi16 = p->psi16;
    la.LBd = m->LBd, la.LPi = m->LPi;
    la.nvec = pass.nvec, la.max_steps = pass.max_steps, la.full_range = pass.full_range ? 1 : 0;
    la.nactive = p->d_nactive + pass.nactive_off;
    la.d0 = p->d_delta, la.d1 = p->d_delta + (size_t)p->max_vec * m->Kp;
    la.ans = p->d_ans, la.score = p->d_score, la.endstate = p->d_endstate, la.ismid = p->d_ismid;
    // models up to 4096 states: the half-precision filter over a tensor-memory-resident table (FLASHV_LEVEL16=0: the float sweep)
    if (m->hi16 && (!getenv("FLASHV_LEVEL16") || atoi(getenv("FLASHV_LEVEL16")) != 0)) {
        const int vg = getenv("FLASHV_LEVEL16_VG") ? atoi(getenv("FLASHV_LEVEL16_VG")) : 4;
        int rc = (vg >= 4 && pass.nvec > 2) ? launch_level16<4>(p, pass, la, done) : launch_level16<2>(p, pass, la, done);
        if (rc != FLASHV_OK || *done) return rc;
    }
    const size_t vec_bytes = (size_t)m->Kp * 4;
    if (pass.nvec >= 5 && 8 * vec_bytes <= 200 * 1024) return launch_level<8, 2, 16>(p, pass, la, done);
    if (pass.nvec >= 3 && 4 * vec_bytes <= 200 * 1024) return launch_level<4, 2, 8>(p, pass, la, done);
    if (2 * vec_bytes <= 200 * 1024) return launch_level<2, 1, 8>(p, pass, la, done);
    return FLASHV_OK;
}

// ---- host orchestration -------------------------------------------------------------------------
template <int QB, int RI, int NWARP>
static cudaError_t launch_step(const StepArgs &a, int nact, int sm_count, cudaStream_t st)
{
    const size_t smem = (size_t)QB * a.Kp * sizeof(float);
    // per launch, not once per process: the attribute is per device and a process may drive several
    cudaError_t e = cudaFuncSetAttribute(k_flash_step<QB, RI, NWARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return e;
    const int ngroups = (nact + QB - 1) / QB;
    const int ncols = a.c_end > 0 ? a.c_end - a.c_begin : a.K;
    const int ntiles = (ncols + NWARP * RI - 1) / (NWARP * RI);
    // blocks that fit at once (shared memory bound), spread over the vector groups
    int per_sm = smem ? (int)((200 * 1024) / smem) : 8;
    per_sm = per_sm < 1 ? 1 : (per_sm > 2048 / (NWARP * 32) ? 2048 / (NWARP * 32) : per_sm);
    int workers = (sm_count * per_sm) / ngroups;  // round down: one extra block would cost a whole second wave
    if (workers > ntiles) workers = ntiles;
    if (workers < 1) workers = 1;
    if (workers > 65535) workers = 65535;
    dim3 grid(ngroups, workers);
    k_flash_step<QB, RI, NWARP><<<grid, NWARP * 32, smem, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t dispatch_step(const StepArgs &a, int nact, int sm_count, cudaStream_t st)
{
    const size_t vec_bytes = (size_t)a.Kp * 4;
    if (nact >= 5 && 8 * vec_bytes <= 200 * 1024) return launch_step<8, 2, 16>(a, nact, sm_count, st);
    if (nact >= 3 && 4 * vec_bytes <= 200 * 1024) return launch_step<4, 2, 8>(a, nact, sm_count, st);
    if (nact >= 2 && 2 * vec_bytes <= 200 * 1024) return launch_step<2, 1, 8>(a, nact, sm_count, st);
    return launch_step<1, 1, 8>(a, nact, sm_count, st);
}

int persistent_pass(flashv_plan *p, const Pass &pass);  // flash_persistent.cu

int flash_run_pass(flashv_plan *p, const Pass &pass, bool time_it)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int K = m->K, Kp = m->Kp, T = p->T;
    const VecDesc *vecs = p->d_vecs + pass.vec_offset;
    float *d0 = p->d_delta, *d1 = p->d_delta + (size_t)p->max_vec * Kp;

    if (time_it) FV_CUDA(cudaEventRecord(ctx->ev[2], st));
    const float *final_delta = d0;
    const bool persistent = p->engine == FLASHV_ENGINE_PERSISTENT && pass.nvec == 1;
    const bool sparse = p->engine == FLASHV_ENGINE_SPARSE && pass.nvec == 1;
    // many vectors over a table small enough that the deltas of a group stay in shared memory (flash_group.cu)
    const bool grouped = p->engine != FLASHV_ENGINE_STEP && pass.nvec >= 2 * ctx->sm_count && group_engine_fits(m);
    if (!grouped && p->engine == FLASHV_ENGINE_PERSISTENT && pass.nvec >= 2 && !pass_is_sharded(p, pass)) {
        // a level of the task tree (or a small batch's N-way pass): everything in one cooperative launch
        bool done = false;
        int rc = level_pass(p, pass, &done);
        if (rc != FLASHV_OK) return rc;
        if (done) {
            if (time_it) FV_CUDA(cudaEventRecord(ctx->ev[3], st));
            return FLASHV_OK;
        }
    }
    if (!grouped) {  // the group kernel builds its start vectors itself
        dim3 ig((K + 255) / 256, pass.nvec < 65535 ? pass.nvec : 65535);
        k_flash_init<<<ig, 256, 0, st>>>(vecs, pass.nvec, p->d_ob, p->d_ans, T, m->LAd, m->LBd, m->LPi, K, Kp, d0);
        FV_CUDA(cudaGetLastError());
        ++p->launches;
    }
    if (grouped) {
        int rc = group_run_pass(p, pass, d1);
        if (rc != FLASHV_OK) return rc;
        final_delta = d1;
    } else if (sparse) {
        int rc = sparse_pass(p, pass);
        if (rc != FLASHV_OK) return rc;
        final_delta = d1;  // like the persistent kernel
    } else if (persistent) {
        int rc = persistent_pass(p, pass);
        if (rc != FLASHV_OK) return rc;
        final_delta = d1;  // the persistent kernel leaves the last delta there
    } else {
        for (int s = 1; s <= pass.max_steps; ++s) {
            StepArgs a;
            a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = K, a.Kp = Kp;
            a.vecs = vecs, a.s = s;
            a.din = (s & 1) ? d0 : d1, a.dout = (s & 1) ? d1 : d0;
            a.ob = p->d_ob, a.T = T, a.psi = p->d_psi, a.psi16 = p->psi16;
            // vectors are sorted longest first: [0, n_cont) go on after this step, [n_cont, n_act) are
            // on their last step and (unless the pass is full-range) need a single column
            const int n_act = pass.nactive[s];
            const int n_cont = pass.full_range ? n_act : pass.nactive[s + 1];
            if (n_cont > 0 && p->engine == FLASHV_ENGINE_SPARSE) {
                int rc = sparse_level_step(p, pass, s, n_cont, a.din, a.dout);
                if (rc != FLASHV_OK) return rc;
            } else if (n_cont > 0) {
                a.nact = n_cont;
                FV_CUDA(dispatch_step(a, n_cont, ctx->sm_count, st));
                ++p->launches;
            }
            if (n_act > n_cont) {
                a.nact = n_act;
                k_flash_last_column<<<n_act - n_cont, LC_WARPS * 32, 0, st>>>(a, n_cont, p->d_ans);
                FV_CUDA(cudaGetLastError());
                ++p->launches;
            }
        }
        final_delta = (pass.max_steps & 1) ? d1 : d0;
    }
    if (time_it) FV_CUDA(cudaEventRecord(ctx->ev[3], st));

    const size_t row_bytes = (size_t)K * (p->psi16 ? 2 : 4);
    const int win_rows = (int)((200 * 1024 - 32) / row_bytes);
    const int bt_rows = pass.first_vec.R - pass.first_vec.mid;
    const int bt_W = win_rows > 0 ? (bt_rows + win_rows - 1) / win_rows : 0;
    const bool par_walk = pass.nvec == 1 && pass.max_steps >= 16 && win_rows >= 4 && bt_W >= 2 && bt_W <= ctx->sm_count &&
                          bt_W < p->bt_windows && ctx->coop && p->d_btmap && !getenv("FLASHV_BACKTRACK_SERIAL");
    if (!par_walk) {
        // Only full-range vectors read delta here, and they run all max_steps steps of the pass.
        k_flash_end<<<pass.nvec, 256, 0, st>>>(vecs, pass.nvec, final_delta, K, Kp, T, p->d_ans, p->d_score, p->d_endstate);
        FV_CUDA(cudaGetLastError());
    }
    if (par_walk) {
        const size_t smem = (size_t)win_rows * row_bytes + 32;
        FV_CUDA(cudaFuncSetAttribute(k_flash_backtrack_par, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const void *bpsi = pass_psi(p, pass);
        int psi16 = p->psi16, wr = win_rows;
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(p->d_sync), bar_base = p->bar_count;
        p->bar_count += 2ull * bt_W;
        int Kp_ = Kp;
        void *params[] = {(void *)&vecs, (void *)&bpsi, (void *)&psi16, (void *)&K, (void *)&T, (void *)&p->d_ismid, (void *)&p->d_endstate,
                          (void *)&p->d_ans, (void *)&wr, (void *)&p->d_btmap, (void *)&bar, (void *)&bar_base, (void *)&final_delta,
                          (void *)&Kp_, (void *)&p->d_score, (void *)&p->d_endstate};
        FV_CUDA(cudaLaunchCooperativeKernel((const void *)k_flash_backtrack_par, dim3(bt_W), dim3(1024), params, smem, st));
    } else if (pass.nvec == 1 && pass.max_steps >= 16 && win_rows >= 4) {
        const size_t smem = (size_t)win_rows * row_bytes + 32;
        FV_CUDA(cudaFuncSetAttribute(k_flash_backtrack_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        k_flash_backtrack_staged<<<1, 1024, smem, st>>>(vecs, pass_psi(p, pass), p->psi16, K, T, p->d_ismid, p->d_endstate, p->d_ans,
                                                      win_rows);
    } else {
        k_flash_backtrack<<<(pass.nvec + 127) / 128, 128, 0, st>>>(vecs, pass.nvec, pass_psi(p, pass), p->psi16, K, T, p->d_ismid,
                                                                   p->d_endstate, p->d_ans);
    }
    FV_CUDA(cudaGetLastError());
    p->launches += 2;
    return FLASHV_OK;
}

// ---- single-step hooks for the per-step parity tests ------------------------------------------
int persistent_single_step(flashv_model *m, const float *d_in_dev, int o, float *d_out_dev, int32_t *psi_dev);

__global__ void k_one_vec(VecDesc *v, int32_t *ob, int o)
{
    v->seq = 0, v->L = 0, v->R = 1, v->mid = 0, v->psi_row = 0, v->flags = 0;
    ob[0] = o, ob[1] = o;
}

int flash_single_step(flashv_model *m, const float *d_in_dev, int o, float *d_out_dev, int32_t *psi_dev, int engine)
{
    if (engine == FLASHV_ENGINE_PERSISTENT) return persistent_single_step(m, d_in_dev, o, d_out_dev, psi_dev);
    flashv_ctx *ctx = m->ctx;
    // a 1-vector, 1-step pass: L=0, R=1, mid=0 so the step at j=1 stores its backpointers in row 0
    VecDesc *dv = reinterpret_cast<VecDesc *>(m->scratch_i);
    int32_t *dob = m->scratch_i + 16;
    k_one_vec<<<1, 1, 0, ctx->stream>>>(dv, dob, o);
    FV_CUDA(cudaGetLastError());
    StepArgs a;
    a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.vecs = dv, a.nact = 1, a.s = 1, a.din = d_in_dev, a.dout = d_out_dev;
    a.ob = dob, a.T = 2, a.psi = psi_dev, a.psi16 = 0;
    FV_CUDA((launch_step<1, 1, 8>(a, 1, ctx->sm_count, ctx->stream)));
    return FLASHV_OK;
}

// One step of ONE vector over the destination columns [c_begin, c_end) only, everything device-resident
// (tools/nccl_baseline.py: the state-sharded pass with the delta exchange done by ncclAllGather between
// per-step launches — the baseline the in-kernel peer stores are measured against).
int flash_step_columns(flashv_model *m, const float *d_in_dev, int o, int c_begin, int c_end, float *d_out_dev, int32_t *psi_dev)
{
    flashv_ctx *ctx = m->ctx;
    VecDesc *dv = reinterpret_cast<VecDesc *>(m->scratch_i);
    int32_t *dob = m->scratch_i + 16;
    k_one_vec<<<1, 1, 0, ctx->stream>>>(dv, dob, o);
    FV_CUDA(cudaGetLastError());
    StepArgs a;
    a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.vecs = dv, a.nact = 1, a.s = 1, a.din = d_in_dev, a.dout = d_out_dev;
    a.ob = dob, a.T = 2, a.psi = psi_dev, a.psi16 = 0;
    a.c_begin = c_begin, a.c_end = c_end;
    FV_CUDA((launch_step<1, 1, 8>(a, 1, ctx->sm_count, ctx->stream)));
    return FLASHV_OK;
}

int flash_single_init(flashv_model *m, int prev_state, int o, float *d_out_dev)
{
    flashv_ctx *ctx = m->ctx;
    // L == 0 selects the pi form; otherwise ans[L-1] must hold prev_state: use L=1 and a 2-entry ans
    VecDesc hv{0, prev_state < 0 ? 0 : 1, 1, 0, 0, 0};
    int32_t hob[2] = {o, o}, hans[2] = {prev_state, 0};
    VecDesc *dv = reinterpret_cast<VecDesc *>(m->scratch_i);
    int32_t *dob = m->scratch_i + 16, *dans = m->scratch_i + 24;
    FV_CUDA(cudaMemcpyAsync(dv, &hv, sizeof(hv), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(dob, hob, sizeof(hob), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(dans, hans, sizeof(hans), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));  // host arrays are on this frame
    k_flash_init<<<dim3((m->K + 255) / 256, 1), 256, 0, ctx->stream>>>(dv, 1, dob, dans, 2, m->LAd, m->LBd, m->LPi, m->K,
                                                                      m->Kp, d_out_dev);
    FV_CUDA(cudaGetLastError());
    return FLASHV_OK;
}

}  // namespace flashv
