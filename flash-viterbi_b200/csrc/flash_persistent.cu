// flash_persistent.cu — engine PERSISTENT: a whole single-vector FLASH pass (the N-way first pass
// of nvviterNdivide, F:126-202, or a full-length task of nvviter, F:204-262) in ONE cooperative
// launch, one CTA per SM.
//
//   * Each CTA owns a contiguous range of destination columns of hiT (K/gridDim of them), i.e. one
//     contiguous ~K*Kp*4/gridDim byte slab that it re-reads every step — from L2 when the table
//     fits (62.9 MB at K=3965 against 126 MB of L2).
//   * Warp NW is the producer: one lane streams the slab through an NSTAGE-deep shared-memory ring
//     with bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx), running ahead across step
//     boundaries since the table does not change.
//   * Warps 0..NW-1 are consumers: a warp takes one column at a time; lane l owns k = 4*(l+32u)+c,
//     keeps four running maxima of the float estimate (FADD, FADD, FMNMX per update), then the
//     warp resolves the exact (value, first index) from the double table for the few candidates
//     inside the window (trellis_common.cuh).  delta lives in registers when Kp <= 4096.
//   * Steps are separated by a grid-wide barrier (release/acquire counter in global memory); the
//     next delta is re-read from L2 with ld.global.cg.
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace flashv {

constexpr int NW = 7;                 // consumer warps (7 + producer = 8 warps: 2 per SM sub-partition, up to 255 registers each)
constexpr int NCONS = NW * 32;        // consumer threads
constexpr int NTHREADS = NCONS + 32;  // + producer warp
constexpr int MAX_STAGES = 16;

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends in hardware for a bounded time; the outer loop is the watchdog
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (TMA, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy, bool hint)
{
    if (hint)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                smem_u32(dst_smem)),
            "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
            : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                     : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PersistArgs {
    const float *hiT;
    const double *LAd;
    const float *LBf;
    int K, Kp;
    const int32_t *ob;  // observations of the sequence this vector walks
    int L, nsteps, mid, psi_row;
    float *d0, *d1;  // step s reads (s odd ? d0 : d1) and writes the other
    void *psi;
    int psi16;
    unsigned *bar;  // zeroed before the launch
    int chunk;      // floats per ring stage (multiple of 128, divides Kp)
    int nstage;
    int l2_hint;
};

// Grid-wide barrier for the consumer threads of all CTAs: `epoch` counts from 1.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned epoch, int ctid)
{
    named_bar_sync(1, NCONS);
    if (ctid == 0) {
        __threadfence();
        red_release_gpu(bar, 1u);
        const unsigned want = epoch * gridDim.x;
        for (uint32_t spins = 0; ld_acquire_gpu(bar) < want; ++spins)
            if (spins > (1u << 28)) __trap();
        __threadfence();
    }
    named_bar_sync(1, NCONS);
}

// NF4 > 0: delta in registers (NF4 float4 per lane, Kp <= 128*NF4, one chunk per column).
// NF4 == 0: delta read from shared memory, columns streamed in `chunk`-float pieces.
template <int NF4>
__global__ void __launch_bounds__(NTHREADS, 1) k_flash_persist(const PersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Kp4 = a.Kp >> 2;
    const int chunk4 = a.chunk >> 2;
    const int nchunks = a.Kp / a.chunk;
    const uint32_t stage_bytes = (uint32_t)a.chunk * 4u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    float4 *sdelta4 = reinterpret_cast<float4 *>(smem_raw + 2 * MAX_STAGES * sizeof(uint64_t));
    unsigned char *ring = reinterpret_cast<unsigned char *>(sdelta4 + Kp4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int c0 = (int)((long long)b * a.K / G), c1 = (int)((long long)(b + 1) * a.K / G);
    const int ncols = c1 - c0;

    if (tid == 0) {
        for (int s = 0; s < a.nstage; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ---------------- producer: the slab of this CTA, once per step, through the ring --------
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();
            const unsigned char *slab = reinterpret_cast<const unsigned char *>(a.hiT + (size_t)c0 * a.Kp);
            uint32_t item = 0;
            for (int s = 1; s <= a.nsteps; ++s)
                for (int w = 0; w < ncols * nchunks; ++w, ++item) {
                    const uint32_t st = item % (uint32_t)a.nstage, use = item / (uint32_t)a.nstage;
                    if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                    mbar_expect_tx(&full[st], stage_bytes);
                    bulk_g2s(ring + (size_t)st * stage_bytes, slab + (size_t)w * stage_bytes, stage_bytes, &full[st], pol,
                             a.l2_hint != 0);
                }
        }
        return;
    }

    // ---------------- consumers ---------------------------------------------------------------
    const float *sdelta = reinterpret_cast<const float *>(sdelta4);
    for (int s = 1; s <= a.nsteps; ++s) {
        const float *din = (s & 1) ? a.d0 : a.d1;
        float *dout = (s & 1) ? a.d1 : a.d0;
        const int j = a.L + s;
        {
            const float4 *din4 = reinterpret_cast<const float4 *>(din);
            for (int t = tid; t < Kp4; t += NCONS) sdelta4[t] = __ldcg(din4 + t);
        }
        named_bar_sync(1, NCONS);
        float4 dreg[NF4 > 0 ? NF4 : 1];
        if (NF4 > 0) {
#pragma unroll
            for (int u = 0; u < NF4; ++u)
                dreg[u] = (lane + 32 * u) < Kp4 ? sdelta4[lane + 32 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float *tmp_row = a.LBf + (size_t)__ldg(a.ob + j) * a.Kp;  // F:167
        const bool keep = j >= a.mid + 1;                                // F:242

        for (int n = warp; n < ncols; n += NW) {
            const int i = c0 + n;
            const float tmp = __ldg(tmp_row + i);
            float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            for (int ch = 0; ch < nchunks; ++ch) {
                const uint32_t item = (uint32_t)((s - 1) * ncols + n) * (uint32_t)nchunks + (uint32_t)ch;
                const uint32_t st = item % (uint32_t)a.nstage, use = item / (uint32_t)a.nstage;
                mbar_wait(&full[st], use & 1);
                const float4 *st4 = reinterpret_cast<const float4 *>(ring + (size_t)st * stage_bytes);
                if (NF4 > 0) {
#pragma unroll
                    for (int u = 0; u < NF4; ++u) {
                        if (lane + 32 * u < chunk4) {
                            const float4 h = st4[lane + 32 * u];
                            const float4 d = dreg[u];
                            cm[0] = fmaxf(cm[0], __fadd_rn(__fadd_rn(tmp, d.x), h.x));
                            cm[1] = fmaxf(cm[1], __fadd_rn(__fadd_rn(tmp, d.y), h.y));
                            cm[2] = fmaxf(cm[2], __fadd_rn(__fadd_rn(tmp, d.z), h.z));
                            cm[3] = fmaxf(cm[3], __fadd_rn(__fadd_rn(tmp, d.w), h.w));
                        }
                    }
                } else {
                    const float4 *d4 = sdelta4 + (size_t)ch * chunk4;
#pragma unroll 4
                    for (int t = lane; t < chunk4; t += 32) {
                        const float4 h = st4[t];
                        const float4 d = d4[t];
                        cm[0] = fmaxf(cm[0], __fadd_rn(__fadd_rn(tmp, d.x), h.x));
                        cm[1] = fmaxf(cm[1], __fadd_rn(__fadd_rn(tmp, d.y), h.y));
                        cm[2] = fmaxf(cm[2], __fadd_rn(__fadd_rn(tmp, d.z), h.z));
                        cm[3] = fmaxf(cm[3], __fadd_rn(__fadd_rn(tmp, d.w), h.w));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
            }
            const Best r = resolve_column(cm, tmp, a.hiT + (size_t)i * a.Kp, sdelta, a.LAd, a.K, a.Kp, i, lane);
            if (lane == 0) {
                dout[i] = r.x;
                if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i, r.k);
            }
        }
        grid_barrier(a.bar, (unsigned)s, tid);
    }
}

// ---- host side ---------------------------------------------------------------------------------
static size_t persist_smem(int Kp, int chunk, int nstage)
{
    return 2 * MAX_STAGES * sizeof(uint64_t) + (size_t)Kp * 4 + (size_t)nstage * chunk * 4;
}

static int launch_persist(flashv_ctx *ctx, PersistArgs &a)
{
    const int Kp = a.Kp;
    const bool in_regs = Kp <= 4096;
    a.chunk = in_regs ? Kp : 4096;
    while (Kp % a.chunk) a.chunk -= 128;  // Kp is a multiple of 128, so this ends at >= 128
    const size_t fixed = persist_smem(Kp, 0, 0);
    int nstage = (int)(((size_t)ctx->smem_optin - fixed) / ((size_t)a.chunk * 4));
    if (nstage > MAX_STAGES) nstage = MAX_STAGES;
    if (nstage < 2) {
        set_error("persistent engine: K=%d does not fit shared memory (%d bytes)", a.K, ctx->smem_optin);
        return FLASHV_ERR_ARG;
    }
    a.nstage = nstage;
    const size_t smem = persist_smem(Kp, a.chunk, nstage);
    const void *fn;
    if (!in_regs)
        fn = (const void *)k_flash_persist<0>;
    else if (Kp <= 1024)
        fn = (const void *)k_flash_persist<8>;
    else if (Kp <= 2048)
        fn = (const void *)k_flash_persist<16>;
    else
        fn = (const void *)k_flash_persist<32>;
    FV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NTHREADS, smem));
    if (per_sm < 1) {
        set_error("persistent engine: kernel does not fit one CTA per SM");
        return FLASHV_ERR_CUDA;
    }
    int grid = ctx->sm_count;
    if (grid > a.K) grid = a.K;
    FV_CUDA(cudaMemsetAsync(a.bar, 0, sizeof(unsigned), ctx->stream));
    void *params[] = {(void *)&a};
    FV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NTHREADS), params, smem, ctx->stream));
    return FLASHV_OK;
}

static int l2_hint_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("FLASHV_L2_HINT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

int persistent_pass(flashv_plan *p, const Pass &pass)
{
    flashv_model *m = p->model;
    const VecDesc &vd = pass.first_vec;  // the pass has exactly one vector (batch == 1)
    PersistArgs a;
    a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.ob = p->d_ob;  // batch == 1 for single-vector passes of sequence 0
    a.L = vd.L, a.nsteps = vd.R - vd.L, a.mid = vd.mid, a.psi_row = vd.psi_row;
    a.d0 = p->d_delta, a.d1 = p->d_delta + (size_t)p->max_vec * m->Kp;
    a.psi = p->d_psi, a.psi16 = p->psi16, a.bar = p->d_sync;
    a.l2_hint = l2_hint_enabled();
    int rc = launch_persist(m->ctx, a);
    if (rc == FLASHV_OK) p->launches += 1;
    return rc;
}

int persistent_single_step(flashv_model *m, const float *d_in_dev, int o, float *d_out_dev, int32_t *psi_dev)
{
    flashv_ctx *ctx = m->ctx;
    int32_t *dob = m->scratch_i + 16;
    int32_t hob[2] = {o, o};
    FV_CUDA(cudaMemcpyAsync(dob, hob, sizeof(hob), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    PersistArgs a;
    a.hiT = m->hiT, a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.ob = dob, a.L = 0, a.nsteps = 1, a.mid = 0, a.psi_row = 0;
    a.d0 = const_cast<float *>(d_in_dev), a.d1 = d_out_dev;
    a.psi = psi_dev, a.psi16 = 0;
    a.bar = reinterpret_cast<unsigned *>(m->scratch_i + 32);
    a.l2_hint = l2_hint_enabled();
    return launch_persist(ctx, a);
}

}  // namespace flashv
