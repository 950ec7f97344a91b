// flash_sparse.cu — engine SPARSE: the single-vector FLASH pass over the in-edge lists of the
// transition graph instead of the dense K x K table.
//
// The reference's update (F:165-174) starts every destination at (-FLT_MAX, -1) and replaces it on
// a strict '>': a source k with A[k][i] == 0 contributes log(0) = -inf and can never win.  The
// HMMs of generate_data/data_script.py have Binomial(K, p) out-edges per row (D:14-26), so a
// fraction 1-p of the table is dead weight: 89 % at the headline p = 0.112.  Dropping those entries
// changes no result and leaves so little data (K=3965: 1.76 M edges) that
//   * a CTA's share of the graph — for each of its ~27 destination columns the ascending list of
//     sources k (16 bit) and the DOUBLE log A[k][i] — is about 120 KB and stays in shared memory
//     for the whole pass: after the prologue a step reads no table byte from L2 or HBM at all;
//   * with the doubles at hand there is no float estimate, no window and no second look: every
//     edge is evaluated with the reference's exact rounding chain
//     (float)((double)(float)(tmp + delta[k]) + logA) — B200's FP64 pipe is half the FP32 rate —
//     and the first maximum is kept lane-locally in ascending k, then reduced with "larger value,
//     then smaller index".
// One warp owns one destination column per round; steps hand over through the same self-validating
// {value, epoch|step} 64-bit words as the dense persistent engine (no grid barrier).  What is left
// of a step is the hand-over latency itself.
//
// This engine is opt-in (FLASHV_ENGINE_SPARSE) and reported separately from the dense roofline:
// its algorithmic bytes are not K^2 * 4 per step.
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c   D: = generate_data/data_script.py
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace flashv {

constexpr int SP_THREADS = 1024;  // 32 warps: one destination column each per round
constexpr unsigned long long SP_WATCHDOG_NS = 4000000000ull;

struct SparseArgs {
    const int *cptr;           // [K+1] in-edge list of column i: entries cptr[i] .. cptr[i+1]-1, ascending k
    const uint16_t *ck;        // [nnz] source state
    const double *cla;         // [nnz] log A[k][i] (the host libm's value, same as LAd)
    const float *LBf;
    int K, Kp;
    const int32_t *ob;
    int L, nsteps, mid, psi_row;
    const float *d_init;
    float *d_final;
    unsigned long long *xch;   // [2][Kp] exchange words, see flash_persistent.cu
    unsigned epoch;
    void *psi;
    int psi16;
    int resident;              // the CTA's lists fit in shared memory
    int max_cta_nnz;           // largest per-CTA edge count (sizes the resident buffers)
};

__device__ __forceinline__ unsigned long long sp_now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Stage delta_{s-1} in shared memory: the plain start vector for s == 1, else the exchange words of
// step s-1, polled until every one carries that step's tag (all of a thread's words are requested
// before any is inspected: one L2 round trip when the data is already there).  Measured
// alternatives, both slower than polling the words themselves (0.84 ms per 255-step pass): polling
// one sentinel word per producer CTA first and reading the vector afterwards (1.03 ms: one more
// dependent round trip), and a separate array of per-CTA flags (2.01 ms: 148 CTAs spinning on the
// same five cache lines serialise in one L2 slice).  Dropping the end-of-step block barrier (ping-pong
// delta buffers) is also slower, 2.2 ms: warps that finish early start polling at once and their
// traffic slows the CTAs that still have to publish — the barrier is the back-off.
__device__ __forceinline__ void sp_delta_load(const SparseArgs &a, int s, float *sdelta, int tid)
{
    if (s == 1) {
        for (int k = tid; k < a.K; k += SP_THREADS) sdelta[k] = __ldcg(a.d_init + k);
    } else {
        const unsigned long long *x = a.xch + (size_t)((s - 1) & 1) * a.Kp;
        const unsigned want = (a.epoch << 16) | (unsigned)(s - 1);
        constexpr int NB = 4;
        for (int k0 = tid; k0 < a.K; k0 += SP_THREADS * NB) {
            unsigned long long w[NB];
            unsigned pending = 0;
#pragma unroll
            for (int e = 0; e < NB; ++e)
                if (k0 + e * SP_THREADS < a.K) pending |= 1u << e;
            unsigned long long t0 = 0;
            for (unsigned spins = 0; pending; ++spins) {
#pragma unroll
                for (int e = 0; e < NB; ++e)
                    if (pending >> e & 1u)
                        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[e]) : "l"(x + k0 + e * SP_THREADS) : "memory");
#pragma unroll
                for (int e = 0; e < NB; ++e)
                    if ((pending >> e & 1u) && (unsigned)(w[e] >> 32) == want) pending &= ~(1u << e);
                if (pending && (spins & 255u) == 255u) {  // a protocol bug must trap, not hang the device
                    const unsigned long long t = sp_now_ns();
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > SP_WATCHDOG_NS) __trap();
                }
            }
#pragma unroll
            for (int e = 0; e < NB; ++e)
                if (k0 + e * SP_THREADS < a.K) sdelta[k0 + e * SP_THREADS] = __uint_as_float((unsigned)w[e]);
        }
    }
    __syncthreads();
}

// dynamic shared memory: float sdelta[Kp]; then (RES) double sla[max_cta_nnz]; uint16_t sk[max_cta_nnz]
template <bool RES>
__global__ void __launch_bounds__(SP_THREADS, 1) k_flash_sparse_pass(const SparseArgs a)
{
    extern __shared__ __align__(16) unsigned char sp_smem[];
    float *sdelta = reinterpret_cast<float *>(sp_smem);
    double *sla = reinterpret_cast<double *>(sp_smem + (size_t)a.Kp * 4);
    uint16_t *sk = reinterpret_cast<uint16_t *>(sla + a.max_cta_nnz);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int c0 = (int)((long long)b * a.K / G), c1 = (int)((long long)(b + 1) * a.K / G);  // owned columns
    const int e_base = a.cptr[c0];
    if (RES) {
        const int n = a.cptr[c1] - e_base;
        for (int e = tid; e < n; e += SP_THREADS) sla[e] = a.cla[e_base + e], sk[e] = a.ck[e_base + e];
    }
    const double *la = RES ? sla - e_base : a.cla;  // indexed by the global entry number either way
    const uint16_t *ks = RES ? sk - e_base : a.ck;
    __syncthreads();

    unsigned long long *const xbuf[2] = {a.xch, a.xch + a.Kp};
    for (int s = 1; s <= a.nsteps; ++s) {
        sp_delta_load(a, s, sdelta, tid);
        const int j = a.L + s;
        const float *tmp_row = a.LBf + (size_t)a.ob[j] * a.Kp;  // F:167
        const bool keep = j >= a.mid + 1, last = s == a.nsteps;
        const unsigned tag = (a.epoch << 16) | (unsigned)s;
        for (int i = c0 + warp; i < c1; i += SP_THREADS / 32) {
            const float tmp = __ldg(tmp_row + i);
            const int e0 = a.cptr[i], e1 = a.cptr[i + 1];
            Best b{-FLT_MAX, 0x7fffffff};
#pragma unroll 4
            for (int e = e0 + lane; e < e1; e += 32) {
                const int k = ks[e];
                const float x = exact_cand(__fadd_rn(tmp, sdelta[k]), la[e]);  // F:170
                if (x > b.x) b.x = x, b.k = k;  // ascending k within a lane: the first maximum stays
            }
            b = warp_best(b);  // larger value, then smaller index
            if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
            if (lane == 0) {
                const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(b.x);
                asm volatile("st.global.u64 [%0], %1;" ::"l"(xbuf[s & 1] + i), "l"(w) : "memory");
                if (last) a.d_final[i] = b.x;
                if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i, b.k);
            }
        }
        __syncthreads();  // sdelta is overwritten by the next step's load
    }
}

// ---- one step of a level of the task tree (many vectors) over the same edge lists -----------------
// grid (vector groups, column workers): a CTA stages delta of SQ vectors in shared memory and its
// warps walk destination columns; a column's edge list is read once (coalesced) and applied to all
// SQ vectors with the exact chain, first maximum kept lane-locally in ascending k.
constexpr int SQ = 8;  // vectors per group

struct SparseStepArgs {
    const int *cptr;
    const uint16_t *ck;
    const double *cla;
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
};

__global__ void __launch_bounds__(512) k_flash_sparse_step(const SparseStepArgs a)
{
    extern __shared__ float4 sp_sdelta4[];
    float *sdelta = reinterpret_cast<float *>(sp_sdelta4);  // [SQ][Kp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int v0 = blockIdx.x * SQ;  // vector groups on x (no 65535 cap), column workers on y
    const int Kp4 = a.Kp >> 2;
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
        const bool live = v0 + q < a.nact;
        const float4 *src = reinterpret_cast<const float4 *>(a.din + (size_t)(v0 + q) * a.Kp);
        for (int t = tid; t < Kp4; t += blockDim.x) sp_sdelta4[q * Kp4 + t] = live ? src[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int jj[SQ], row[SQ];
    const float *tmp_row[SQ];
#pragma unroll
    for (int q = 0; q < SQ; ++q) {
        jj[q] = 0, row[q] = -1, tmp_row[q] = a.LBf;
        if (v0 + q < a.nact) {
            const VecDesc vd = a.vecs[v0 + q];
            jj[q] = vd.L + a.s;
            tmp_row[q] = a.LBf + (size_t)a.ob[(size_t)vd.seq * a.T + jj[q]] * a.Kp;  // F:233
            if (jj[q] >= vd.mid + 1) row[q] = vd.psi_row + (jj[q] - vd.mid - 1);     // F:242
        }
    }
    __syncthreads();
    for (int i = blockIdx.y * nwarp + warp; i < a.K; i += gridDim.y * nwarp) {
        float tmp[SQ];
        Best b[SQ];
#pragma unroll
        for (int q = 0; q < SQ; ++q) tmp[q] = __ldg(tmp_row[q] + i), b[q] = Best{-FLT_MAX, 0x7fffffff};
        const int e0 = a.cptr[i], e1 = a.cptr[i + 1];
#pragma unroll 2
        for (int e = e0 + lane; e < e1; e += 32) {
            const int k = __ldg(a.ck + e);
            const double la = __ldg(a.cla + e);
#pragma unroll
            for (int q = 0; q < SQ; ++q) {
                const float x = exact_cand(__fadd_rn(tmp[q], sdelta[q * a.Kp + k]), la);  // F:236
                if (x > b[q].x) b[q].x = x, b[q].k = k;
            }
        }
#pragma unroll
        for (int q = 0; q < SQ; ++q) {
            if (v0 + q >= a.nact) continue;  // warp-uniform: padding vector of the last group
            Best r = warp_best(b[q]);
            if (!(r.x > -FLT_MAX)) r.x = -FLT_MAX, r.k = -1;
            if (lane == 0) {
                a.dout[(size_t)(v0 + q) * a.Kp + i] = r.x;
                if (row[q] >= 0) psi_store(a.psi, a.psi16, (size_t)row[q] * a.K + i, r.k);
            }
        }
    }
}

int sparse_level_step(flashv_plan *p, const Pass &pass, int s, int nact, const float *din, float *dout)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    SparseStepArgs a;
    a.cptr = m->csc_ptr, a.ck = m->csc_k, a.cla = m->csc_la, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.vecs = p->d_vecs + pass.vec_offset, a.nact = nact, a.s = s, a.din = din, a.dout = dout;
    a.ob = p->d_ob, a.T = p->T, a.psi = p->d_psi, a.psi16 = p->psi16;
    const size_t smem = (size_t)SQ * m->Kp * sizeof(float);
    if (smem > (size_t)ctx->smem_optin) {
        set_error("sparse engine: K=%d does not fit the level kernel's shared memory", m->K);
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaFuncSetAttribute(k_flash_sparse_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ngroups = (nact + SQ - 1) / SQ;
    int per_sm = (int)((size_t)ctx->smem_optin / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    int workers = (ctx->sm_count * per_sm) / ngroups;
    const int max_workers = (m->K + 15) / 16;  // 16 warps per CTA, at least one column each
    workers = workers < 1 ? 1 : (workers > max_workers ? max_workers : workers);
    k_flash_sparse_step<<<dim3(ngroups, workers), 512, smem, ctx->stream>>>(a);
    FV_CUDA(cudaGetLastError());
    ++p->launches;
    return FLASHV_OK;
}

// ---- the edge lists, built once per model ON THE DEVICE from the double log table ---------------------
// (the host only computes logarithms; everything that is a re-arrangement of them runs here, so a
// model whose rows were computed by several ranks and assembled over NVLink gets the same lists)
//
// In-edge lists (by destination i, ascending source k): the source axis is cut into CS slices; thread
// (i, slice) counts, an exclusive scan over the (i-major, slice-minor) counts gives every thread its
// first entry — which is already the list order — and a second walk fills.
constexpr int CS = 32;

__global__ void k_csc_count(const double *__restrict__ LAd, int K, int kslice, int *__restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, sl = blockIdx.y;
    if (i >= K) return;
    const int k1 = min(K, (sl + 1) * kslice);
    int c = 0;
    for (int k = sl * kslice; k < k1; ++k) c += LAd[(size_t)k * K + i] > -INFINITY;
    cnt[(size_t)i * CS + sl] = c;
}

__global__ void k_csc_fill(const double *__restrict__ LAd, int K, int kslice, const int *__restrict__ off,
                           uint16_t *__restrict__ ck, double *__restrict__ cla, int *__restrict__ ptr, int nnz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, sl = blockIdx.y;
    if (i >= K) return;
    int e = off[(size_t)i * CS + sl];
    if (sl == 0) ptr[i] = e;
    if (sl == 0 && i == K - 1) ptr[K] = nnz;
    const int k1 = min(K, (sl + 1) * kslice);
    for (int k = sl * kslice; k < k1; ++k) {
        const double v = LAd[(size_t)k * K + i];
        if (v > -INFINITY) ck[e] = (uint16_t)k, cla[e] = v, ++e;
    }
}

// Out-edge lists (by source k, ascending destination i), one warp per row.
__global__ void k_csr_count(const double *__restrict__ LAd, int K, int *__restrict__ cnt)
{
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= K) return;
    int c = 0;
    for (int i = lane; i < K; i += 32) c += LAd[(size_t)k * K + i] > -INFINITY;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if (lane == 0) cnt[k] = c;
}

__global__ void k_csr_fill(const double *__restrict__ LAd, int K, const int *__restrict__ rowstart,
                           uint16_t *__restrict__ ri, double *__restrict__ rla)
{
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= K) return;
    int e = rowstart[k];
    for (int i0 = 0; i0 < K; i0 += 32) {
        const int i = i0 + lane;
        const double v = i < K ? LAd[(size_t)k * K + i] : -INFINITY;
        const unsigned m = __ballot_sync(FULL_MASK, v > -INFINITY);
        if (v > -INFINITY) {
            const int at = e + __popc(m & ((1u << lane) - 1u));
            ri[at] = (uint16_t)i, rla[at] = v;
        }
        e += __popc(m);
    }
}

// cut[k][q] = first entry of row k whose destination is >= q*per8 (q = 0..8): the ranges a cluster's CTAs own
__global__ void k_csr_cuts(const uint16_t *__restrict__ ri, const int *__restrict__ rowstart, int K, int nnz, int per8,
                           int *__restrict__ cut)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * 9) return;
    const int k = t / 9, q = t - 9 * k;
    int lo = rowstart[k], hi = k + 1 < K ? rowstart[k + 1] : nnz;
    const int want = q * per8;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)ri[mid] < want) lo = mid + 1;
        else hi = mid;
    }
    cut[t] = lo;
}

static int exclusive_scan(int *d_in, int *d_out, int n, cudaStream_t st)
{
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    FV_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_in, d_out, n, st));
    FV_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_in, d_out, n, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    FV_CUDA(e);
    return FLASHV_OK;
}

int sparse_build(flashv_model *m)
{
    const int K = m->K;
    if (K >= 65536) return FLASHV_OK;  // 16-bit source indices: larger models keep the dense engines only
    if (const char *e = getenv("FLASHV_NO_SPARSE"))
        if (atoi(e) != 0) return FLASHV_OK;  // skip the edge lists (saves their build time and memory on very large models)
    flashv_ctx *ctx = m->ctx;
    cudaStream_t st = ctx->stream;
    const int kslice = (K + CS - 1) / CS;
    const size_t ncnt = (size_t)K * CS;
    int *d_cnt = nullptr, *d_off = nullptr;
    FV_CUDA(cudaMalloc(&d_cnt, (ncnt + K) * sizeof(int)));
    FV_CUDA(cudaMalloc(&d_off, (ncnt + K) * sizeof(int)));
    int rc = FLASHV_OK;
    auto done = [&](int code) {
        cudaFree(d_cnt), cudaFree(d_off);
        return code;
    };
    k_csc_count<<<dim3((K + 127) / 128, CS), 128, 0, st>>>(m->LAd, K, kslice, d_cnt);
    if ((rc = exclusive_scan(d_cnt, d_off, (int)ncnt, st)) != FLASHV_OK) return done(rc);
    int last[2];
    cudaError_t e = cudaMemcpy(&last[0], d_off + ncnt - 1, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&last[1], d_cnt + ncnt - 1, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return done(cuda_fail(e, "edge count", __FILE__, __LINE__));
    const size_t nnz = (size_t)last[0] + (size_t)last[1];
    if (nnz == 0 || nnz > (size_t)K * K / 2) return done(FLASHV_OK);  // dense enough that the dense engines are the better fit

#define SB_CUDA(call)                                                                  \
    if ((e = (call)) != cudaSuccess) return done(cuda_fail(e, #call, __FILE__, __LINE__));
    SB_CUDA(cudaMalloc(&m->csc_ptr, ((size_t)K + 1) * sizeof(int)));
    SB_CUDA(cudaMalloc(&m->csc_k, nnz * sizeof(uint16_t)));
    SB_CUDA(cudaMalloc(&m->csc_la, nnz * sizeof(double)));
    k_csc_fill<<<dim3((K + 127) / 128, CS), 128, 0, st>>>(m->LAd, K, kslice, d_off, m->csc_k, m->csc_la, m->csc_ptr, (int)nnz);
    SB_CUDA(cudaGetLastError());
    m->csc_nnz = (long long)nnz;
    // the largest per-CTA edge count for the grid the pass kernel uses
    std::vector<int> ptr((size_t)K + 1);
    SB_CUDA(cudaMemcpyAsync(ptr.data(), m->csc_ptr, ptr.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    const int G = ctx->sm_count < K ? ctx->sm_count : K;
    int worst = 0;
    for (int b = 0; b < G; ++b) {
        const int c0 = (int)((long long)b * K / G), c1 = (int)((long long)(b + 1) * K / G);
        worst = std::max(worst, ptr[c1] - ptr[c0]);
    }
    m->csc_max_cta_nnz = worst;
    m->bytes += ((size_t)K + 1) * sizeof(int) + nnz * (sizeof(uint16_t) + sizeof(double));

    // The same edges by SOURCE for FLASH-BS (a step only looks at the out-edges of the B beam states),
    // with every row cut at the 8 destination-range boundaries q*ceil(K/8) a cluster's CTAs own.
    int *d_rowcnt = d_cnt + ncnt, *d_rowstart = d_off + ncnt;
    k_csr_count<<<(K + 7) / 8, 256, 0, st>>>(m->LAd, K, d_rowcnt);
    SB_CUDA(cudaGetLastError());
    if ((rc = exclusive_scan(d_rowcnt, d_rowstart, K, st)) != FLASHV_OK) return done(rc);
    SB_CUDA(cudaMalloc(&m->csr_cut, (size_t)K * 9 * sizeof(int)));
    SB_CUDA(cudaMalloc(&m->csr_i, nnz * sizeof(uint16_t)));
    SB_CUDA(cudaMalloc(&m->csr_la, nnz * sizeof(double)));
    k_csr_fill<<<(K + 7) / 8, 256, 0, st>>>(m->LAd, K, d_rowstart, m->csr_i, m->csr_la);
    SB_CUDA(cudaGetLastError());
    k_csr_cuts<<<(K * 9 + 255) / 256, 256, 0, st>>>(m->csr_i, d_rowstart, K, (int)nnz, (K + 7) / 8, m->csr_cut);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaStreamSynchronize(st));
#undef SB_CUDA
    m->bytes += (size_t)K * 9 * sizeof(int) + nnz * (sizeof(uint16_t) + sizeof(double));
    return done(FLASHV_OK);
}

bool sparse_engine_available(const flashv_model *m) { return m->csc_ptr != nullptr; }

int sparse_pass(flashv_plan *p, const Pass &pass)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    const VecDesc &vd = pass.first_vec;  // the pass has exactly one vector (batch == 1)
    SparseArgs a;
    a.cptr = m->csc_ptr, a.ck = m->csc_k, a.cla = m->csc_la, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.ob = p->d_ob;
    a.L = vd.L, a.nsteps = vd.R - vd.L, a.mid = vd.mid, a.psi_row = vd.psi_row;
    a.d_init = p->d_delta, a.d_final = p->d_delta + (size_t)p->max_vec * m->Kp;
    a.xch = reinterpret_cast<unsigned long long *>(p->d_delta + (size_t)2 * p->max_vec * m->Kp);
    a.psi = p->d_psi, a.psi16 = p->psi16;
    a.epoch = (++p->run_epoch) & 0xffffu;
    if (a.epoch == 0) a.epoch = (++p->run_epoch) & 0xffffu;  // tag 0 is what a fresh buffer holds
    if (a.nsteps >= 65536) {
        set_error("sparse engine: more than 65535 steps in one pass");
        return FLASHV_ERR_ARG;
    }
    a.max_cta_nnz = (m->csc_max_cta_nnz + 7) & ~7;
    const size_t res_bytes = (size_t)m->Kp * 4 + (size_t)a.max_cta_nnz * (sizeof(double) + sizeof(uint16_t)) + 16;
    a.resident = res_bytes <= (size_t)ctx->smem_optin && !(getenv("FLASHV_SPARSE_RESIDENT") && atoi(getenv("FLASHV_SPARSE_RESIDENT")) == 0);
    const size_t smem = a.resident ? res_bytes : (size_t)m->Kp * 4 + 16;
    if (!a.resident) a.max_cta_nnz = 0;
    const void *fn = a.resident ? (const void *)k_flash_sparse_pass<true> : (const void *)k_flash_sparse_pass<false>;
    FV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = ctx->sm_count < m->K ? ctx->sm_count : m->K;  // the grid sparse_build sized the buffers for
    void *params[] = {(void *)&a};
    // cooperative launch: every CTA polls data the others produce, so all must be co-resident
    FV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(SP_THREADS), params, smem, ctx->stream));
    ++p->launches;
    return FLASHV_OK;
}

}  // namespace flashv
