// flash_persistent.cu — engine PERSISTENT: a whole single-vector FLASH pass (the N-way first pass
// of nvviterNdivide, F:126-202, or a full-length task of nvviter, F:204-262) in ONE cooperative
// launch, one CTA per SM.
//
//   * Each CTA owns a contiguous range of destination columns (K/gridDim of them) and re-reads their
//     log-A entries every step — from L2 when the table fits (62.9 MB at K=3965 against 126 MB of
//     L2; ncu: 97 % L2 hit rate, 0.25 % DRAM).  The table is stored CTA-tiled (hiC, tile_geom.h):
//     per CTA, per chunk of TILE_CH source states, the chunk of every owned column back to back,
//     so one step of one CTA is one linear stream and every ring stage is one contiguous bulk copy.
//   * Warp NCW is the producer: one lane streams the slab through an NSTAGE-deep shared-memory ring
//     with bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx), running ahead across step
//     boundaries since the table does not change — the ring refills while the CTAs hand over.
//   * Warps 0..NCW-1 are consumers and ALL of them work on every stage: warp w owns CPW fixed
//     columns of the round (rows 2w, 2w+1 of the stage), lane l owns k = 4*(l+32u)+c and keeps four
//     running maxima per column of the float estimate (FADD, FADD, FMNMX per update); a stage is
//     released after NCW arrivals, so the ring is a true FIFO and no warp ever holds a stage longer
//     than one pass over its two rows.  The exact (value, first index) comes from the double table
//     for the few candidates inside the window (trellis_common.cuh): the winning chain is re-read
//     from the tiled table (L2), the double loads of both columns are issued before either is
//     consumed.
//   * Steps hand over through the delta vector itself ({value, step} words polled from L2, see
//     delta_wait_load): no grid barrier, no atomics.
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "flashv_internal.h"
#include "tile_geom.h"
#include "trellis_common.cuh"

namespace flashv {

constexpr int NCW = TILE_RW / 2;      // consumer warps (14)
constexpr int CPW = 2;                // columns a consumer warp owns per round
constexpr int NCONS = NCW * 32;       // consumer threads
constexpr int NTHREADS = NCONS + 32;  // + producer warp
constexpr int MAX_STAGES = 64;
constexpr int TRACE_STEPS = 64, TRACE_PTS = 7;
constexpr int CTRL_BYTES = 2 * MAX_STAGES * 8 + 64;  // full[], empty[], the TMEM base address slot

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
// Every spin loop of this kernel is bounded in wall time: a protocol bug must end in a trap within
// seconds, not hang the device.
constexpr unsigned long long WATCHDOG_NS = 4000000000ull;
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void watchdog(uint32_t spins, unsigned long long &t0, unsigned long long limit_ns = WATCHDOG_NS)
{
    if ((spins & 1023u) == 1023u) {
        const unsigned long long t = now_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > limit_ns) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) watchdog(spins, t0);
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
// ---- tensor memory as resident table storage ---------------------------------------------------
// A CTA re-reads the same ~428 KB of log A every step, and what bounds the stream phase is the
// rate at which L2 can refill the ring (~60 GB/s per SM while all 148 SMs pull).  Blackwell's
// 256 KB of tensor memory per SM is idle in this kernel, so the first TM_RES_CHUNKS chunks (2048
// source states) of every owned column are parked there ONCE, at kernel start (tcgen05.st), and
// every step reads them back with tcgen05.ld; only the rest of each column goes through the ring
// — little enough that the ring refills it completely while the CTAs hand over.
// Layout: a warp can only touch the 32 TMEM lanes of its quadrant (warp % 4) and owns two fixed
// columns, so warp w keeps its own data in columns [128*(w/4), +128) of its quadrant: for
// iteration u (128 source states) TMEM lane l, columns 8u..8u+3 / 8u+4..8u+7 hold the float4 that
// lane l needs for its first / second column.
constexpr int TMEM_COLS = 512;     // the whole tensor memory of the SM
constexpr int TM_RES_CHUNKS = 8;   // TILE_CH-sized chunks of every owned column kept resident (2048 states)
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

struct PersistArgs {
    const float *hiC;  // CTA-tiled (float)log A of the columns [col_begin, col_begin+ncol), tile_geom.h
    int col_begin, ncol;  // destination columns this GPU owns (all of them unless the pass is state-sharded)
    const double *LAd;
    const float *LBf;
    int K, Kp;
    const int32_t *ob;  // observations of the sequence this vector walks
    int L, nsteps, mid, psi_row;
    const float *d_init;      // delta of the start vector (plain floats, written by k_flash_init)
    float *d_final;           // delta after the last step (plain floats, for k_flash_end)
    unsigned long long *xch;  // [2][Kp] exchange buffers of {value, step} words, zeroed before the launch
    // state-sharded pass: every GPU publishes its slice of delta and of the backpointer
    // rows into the buffers of ALL GPUs (peer stores over NVLink) and polls only its own copy
    unsigned epoch;                    // run counter: step tags are (epoch << 16 | step), so stale words never match
    int npeer;                         // GPUs taking part, this one included (1 = not sharded)
    unsigned long long *xch_peer[8];   // their exchange buffers (entry `rank` is xch itself)
    void *psi_peer[8];                 // their backpointer stores
    void *psi;
    int psi16;
    int nstage;
    int l2_hint;
    unsigned long long watchdog_ns;  // how long a poll for delta words may last before the kernel traps (longer across GPUs)
    long long *trace;  // optional [TRACE_STEPS][grid][2 warps][TRACE_PTS] clock64 samples (FLASHV_TRACE_FILE), else null
};

// ---- step-to-step hand-over without a barrier ------------------------------------------------
// The only data a step needs from the other CTAs is the previous delta vector, so the vector itself
// carries the flag: every entry is published as ONE 64-bit store {float bits, step number} into
// an exchange buffer, and every CTA polls the buffer of the previous step (ld.volatile, served by
// L2) until all K entries show that step number, staging the values in shared memory as it goes.
// A 64-bit store is single-copy atomic, so no fence, no atomic and no ordering between different
// stores is needed.  Two buffers ping-pong: step s writes X[s&1] and reads X[(s-1)&1]; overwriting
// X[s&1] at step s is safe because its previous content (step s-2) was read at the start of step
// s-1, and a CTA can only be in step s once every CTA has published its step s-1 output, i.e. has
// finished that read.  The step number is tagged with a per-run epoch, so words left over from an
// earlier run (or not yet overwritten by a slower GPU) never match and nothing has to be cleared.
__device__ __forceinline__ void ld_volatile_2x64(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void publish_delta(unsigned long long *xch, int i, float v, int step)
{
    const unsigned long long w = ((unsigned long long)(unsigned)step << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.global.u64 [%0], %1;" ::"l"(xch + i), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned step_tag(const PersistArgs &a, int step) { return (a.epoch << 16) | (unsigned)step; }
// Sharded form: the word goes to every GPU (plain 64-bit stores; the tag makes each one self-
// validating).  Only the LAST step of a pass needs ordering: the final hand-shake promises that
// once a GPU's last delta word is visible, so are its backpointer entries, hence the system-scope
// fence between them there.
__device__ __forceinline__ void publish_delta_peers(const PersistArgs &a, int parity, int i, float v, int step, bool last)
{
    const unsigned long long w = ((unsigned long long)step_tag(a, step) << 32) | (unsigned long long)__float_as_uint(v);
    if (last) __threadfence_system();
    for (int r = 0; r < a.npeer; ++r)
        asm volatile("st.global.u64 [%0], %1;" ::"l"(a.xch_peer[r] + (size_t)parity * a.Kp + i), "l"(w) : "memory");
}

// Stage delta_{s-1} in shared memory: from the plain start vector for s == 1, else from the
// exchange buffer, waiting for every entry to carry step s-1.
__device__ __forceinline__ void delta_wait_load(const PersistArgs &a, int s, float *sdelta, int ctid)
{
    const int Kp2 = a.Kp >> 1;
    float2 *sd2 = reinterpret_cast<float2 *>(sdelta);
    if (s == 1) {
        const float2 *in2 = reinterpret_cast<const float2 *>(a.d_init);
        for (int t = ctid; t < Kp2; t += NCONS) {
            float2 v = __ldcg(in2 + t);
            if (2 * t >= a.K) v.x = 0.f;  // padding lanes stay finite (hiT pads with -inf)
            if (2 * t + 1 >= a.K) v.y = 0.f;
            sd2[t] = v;
        }
    } else {
        // All of a thread's entries are requested before any is inspected, so a step pays one L2
        // round trip here when the data is already there, not one per entry.
        const unsigned long long *x = a.xch + (size_t)((s - 1) & 1) * a.Kp;
        const unsigned want = step_tag(a, s - 1);
        constexpr int NB = 8;  // entries (pairs of delta values) in flight per thread
        for (int t0 = ctid; t0 < Kp2; t0 += NCONS * NB) {
            unsigned long long w0[NB], w1[NB];
            unsigned pending = 0;
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const int t = t0 + e * NCONS;
                if (t < Kp2 && 2 * t < a.K) pending |= 1u << e;
            }
            unsigned long long tw = 0;
            for (uint32_t spins = 0; pending; ++spins) {
#pragma unroll
                for (int e = 0; e < NB; ++e)
                    if (pending >> e & 1u) ld_volatile_2x64(x + 2 * (t0 + e * NCONS), w0[e], w1[e]);
#pragma unroll
                for (int e = 0; e < NB; ++e)
                    if (pending >> e & 1u) {
                        const int k = 2 * (t0 + e * NCONS);
                        if ((unsigned)(w0[e] >> 32) == want && (k + 1 >= a.K || (unsigned)(w1[e] >> 32) == want))
                            pending &= ~(1u << e);
                    }
                if (pending) watchdog(spins * 64u + 63u, tw, a.watchdog_ns);
            }
#pragma unroll
            for (int e = 0; e < NB; ++e) {
                const int t = t0 + e * NCONS;
                if (t < Kp2) {
                    const int k = 2 * t;
                    float2 v = make_float2(0.f, 0.f);  // padding lanes stay finite (the table pads with -inf)
                    if (k < a.K) v.x = __uint_as_float((unsigned)w0[e]);
                    if (k + 1 < a.K) v.y = __uint_as_float((unsigned)w1[e]);
                    sd2[t] = v;
                }
            }
        }
    }
    named_bar_sync(1, NCONS);
}

// Candidates of one column inside the window, gathered lane-locally: the newest candidate's double
// load stays in flight (`la`), older ones of the same lane are folded into `acc` on arrival.
struct Pending {
    Best acc;
    double la;
    float pre;
    int k;
    bool has;
};

__device__ __forceinline__ void pending_push(Pending &p, float pre, int k, const double *__restrict__ la_ptr)
{
    if (p.has) {
        const float x = exact_cand(p.pre, p.la);
        if (x > -FLT_MAX) best_take(p.acc, x, p.k);
    }
    p.pre = pre, p.k = k, p.la = __ldg(la_ptr), p.has = true;
}

__device__ __forceinline__ Best pending_finish(Pending &p)
{
    if (p.has) {
        const float x = exact_cand(p.pre, p.la);
        if (x > -FLT_MAX) best_take(p.acc, x, p.k);
    }
    Best b = warp_best(p.acc);
    if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
    return b;
}

// Window scan of one column, split in two so that the L2 reads of several columns overlap:
// scan_fetch() finds the chains whose maximum lies inside the window and loads, per lane, the
// estimate inputs of (up to) SCAN_SLOTS chain elements from the tiled table; scan_commit() compares
// them against the window and starts the exact double loads.  Anything that does not fit the slots
// (several chains inside the window, or chains longer than 32 elements) takes the slow path
// inside scan_commit(), which re-reads synchronously.
constexpr int SCAN_SLOTS = 2;
struct Scan {
    float hi0, hi1;  // slot s holds element `lane` of the s-th chain inside the window
    int k0, k1;      // its source state, or -1 if this lane has no element in that slot
    int thr;         // window threshold (ordinal), warp-uniform
    unsigned overflow;  // warp-uniform: some chain did not fit the slots
    bool dead;       // warp-uniform: no finite estimate in the column
};

__device__ __forceinline__ void scan_fetch(Scan &sc, const float (&cm)[4], const float *__restrict__ round_base,
                                           int ncr, int rr, int K, int Kp, int lane)
{
    sc.k0 = sc.k1 = -1, sc.hi0 = sc.hi1 = 0.f, sc.overflow = 0;
    const float top = warp_max(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
    sc.dead = !(top > -FLT_MAX);
    sc.thr = ford(top) - WINDOW_STEPS;
    if (sc.dead) return;
    const int chain_len = Kp >> 7;
    int used = 0;  // warp-uniform count of chains taken
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned hit = __ballot_sync(FULL_MASK, ford(cm[c]) >= sc.thr);
        while (hit) {
            const int w = __ffs(hit) - 1;
            hit &= hit - 1;
            if (used >= SCAN_SLOTS || chain_len > 32) {
                sc.overflow = 1;
                continue;
            }
            const int k = 4 * (w + 32 * lane) + c;  // element `lane` of chain (w, c)
            // `used` is warp-uniform: branch, so that the load lands in its slot without being
            // consumed (a select on the loaded value would wait for it right here)
            const bool mine = lane < chain_len && k < K;
            const float *src = round_base + tile_round_off(Kp, ncr, rr, mine ? k : 0);
            if (used == 0) {
                if (mine) sc.k0 = k, sc.hi0 = __ldg(src);
            } else {
                if (mine) sc.k1 = k, sc.hi1 = __ldg(src);
            }
            ++used;
        }
    }
}

__device__ __forceinline__ void scan_commit(Pending &p, const Scan &sc, const float (&cm)[4], float tmp,
                                            const float *__restrict__ round_base, int ncr, int rr, const float *sdelta,
                                            const double *__restrict__ LAd, int K, int Kp, int i, int lane)
{
    p.acc = Best{-FLT_MAX, 0x7fffffff};
    p.has = false;
    p.la = 0.0, p.pre = 0.f, p.k = 0;
    if (sc.dead) return;
    if (!sc.overflow) {
        if (sc.k0 >= 0) {
            const float pre = __fadd_rn(tmp, sdelta[sc.k0]);
            if (ford(__fadd_rn(pre, sc.hi0)) >= sc.thr) pending_push(p, pre, sc.k0, LAd + (size_t)sc.k0 * K + i);
        }
        if (sc.k1 >= 0) {
            const float pre = __fadd_rn(tmp, sdelta[sc.k1]);
            if (ford(__fadd_rn(pre, sc.hi1)) >= sc.thr) pending_push(p, pre, sc.k1, LAd + (size_t)sc.k1 * K + i);
        }
        return;
    }
    // slow path: every chain inside the window, element by element
    const int chain_len = Kp >> 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned hit = __ballot_sync(FULL_MASK, ford(cm[c]) >= sc.thr);
        while (hit) {
            const int w = __ffs(hit) - 1;
            hit &= hit - 1;
            for (int u = lane; u < chain_len; u += 32) {
                const int k = 4 * (w + 32 * u) + c;
                if (k < K) {
                    const float pre = __fadd_rn(tmp, sdelta[k]);
                    const float est = __fadd_rn(pre, __ldg(round_base + tile_round_off(Kp, ncr, rr, k)));
                    if (ford(est) >= sc.thr) pending_push(p, pre, k, LAd + (size_t)k * K + i);
                }
            }
        }
    }
}

// TM: the first TM_RES_CHUNKS chunks of every owned column live in tensor memory (single-round CTAs only).
template <bool TM>
__global__ void __launch_bounds__(NTHREADS, 1) k_flash_persist(const PersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Kp4 = a.Kp >> 2;
    const int nk = (a.Kp + TILE_CH - 1) / TILE_CH;  // chunks per column
    constexpr uint32_t STAGE_BYTES = TILE_RW * TILE_CH * 4u;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(empty + MAX_STAGES);  // TMEM base address from tcgen05.alloc
    float4 *sdelta4 = reinterpret_cast<float4 *>(smem_raw + CTRL_BYTES);
    unsigned char *ring = reinterpret_cast<unsigned char *>(sdelta4 + Kp4);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int c0 = tile_c0(a.ncol, G, b), ncols = tile_c0(a.ncol, G, b + 1) - c0;  // local column indices
    const int nrounds = (ncols + TILE_RW - 1) / TILE_RW;
    const float *slab = a.hiC + (size_t)c0 * a.Kp;  // this CTA's columns: ncols*Kp floats, in stream order

    if (tid == 0) {
        for (int s = 0; s < a.nstage; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], NCW);
        mbar_fence_init();
    }
    if (TM && warp == 0) tmem_alloc(tmem_slot);
    if (TM) tmem_fence_before();
    __syncthreads();
    if (TM) tmem_fence_after();
    // this warp's window into tensor memory: its quadrant's lanes, its own 128 columns
    const uint32_t tbase = TM ? (*tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2)) : 0u;
    const int nk_res = TM ? min(TM_RES_CHUNKS, a.Kp / TILE_CH) : 0;  // resident chunks (full ones only)

    if (warp == NCW) {
        // ---------------- producer: the slab, once per step, linearly through the ring -----------
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();
            int st = 0;
            uint32_t use = 0;  // how often the ring has wrapped: stage st is being filled for the use-th time
            for (int s = 1; s <= a.nsteps; ++s) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(slab);
                for (int rho = 0; rho < nrounds; ++rho) {
                    const int ncr = min(TILE_RW, ncols - rho * TILE_RW);
                    src += (size_t)nk_res * TILE_CH * ncr * sizeof(float);  // resident chunks are not streamed
                    for (int u = nk_res; u < nk; ++u) {
                        if (use > 0) mbar_wait(&empty[st], (use - 1) & 1);
                        const uint32_t bytes = (uint32_t)(ncr * min(TILE_CH, a.Kp - u * TILE_CH)) * 4u;
                        mbar_expect_tx(&full[st], bytes);
                        bulk_g2s(ring + (size_t)st * STAGE_BYTES, src, bytes, &full[st], pol, a.l2_hint != 0);
                        src += bytes;
                        if (++st == a.nstage) st = 0, ++use;
                    }
                }
            }
        }
        return;
    }

    // ---------------- consumers ---------------------------------------------------------------
    const float *sdelta = reinterpret_cast<const float *>(sdelta4);
    if (TM) {
        // park the resident chunks of this warp's two columns (round 0 is the only round in TM mode)
        const int ncr = ncols, rr0 = warp * CPW, rr1 = rr0 + 1;
        if (rr0 < ncr) {
            const int ra = rr0, rb = rr1 < ncr ? rr1 : rr0;
            for (int u = 0; u < nk_res * (TILE_CH / 128); ++u) {
                const int k = 4 * (lane + 32 * u);
                tmem_st4(tbase + 8u * (uint32_t)u, __ldg(reinterpret_cast<const float4 *>(slab + tile_round_off(a.Kp, ncr, ra, k))));
                tmem_st4(tbase + 8u * (uint32_t)u + 4u, __ldg(reinterpret_cast<const float4 *>(slab + tile_round_off(a.Kp, ncr, rb, k))));
            }
        }
        tmem_wait_st();
        tmem_fence_before();
        named_bar_sync(1, NCONS);
        tmem_fence_after();
    }
    int st = 0;           // every consumer warp visits every ring item, in the producer's order
    uint32_t parity = 0;  // parity of the ring wrap count = phase parity to wait for
    const int nk_full = a.Kp / TILE_CH;  // chunks of exactly TILE_CH states; at most one shorter chunk follows
    for (int s = 1; s <= a.nsteps; ++s) {
        unsigned long long *xout = a.xch + (size_t)(s & 1) * a.Kp;
        const bool last_step = s == a.nsteps;
        const int j = a.L + s;
        const float *tmp_row = a.LBf + (size_t)__ldg(a.ob + j) * a.Kp;  // F:167
        const bool keep = j >= a.mid + 1;                                // F:242
        const bool tracing = a.trace != nullptr && s <= TRACE_STEPS && lane == 0 && (warp == 0 || warp == NCW - 1);
        long long *tr = tracing ? a.trace + ((((size_t)(s - 1) * G + b) * 2 + (warp == 0 ? 0 : 1)) * TRACE_PTS) : nullptr;
        if (tracing) tr[0] = clock64();
        delta_wait_load(a, s, reinterpret_cast<float *>(sdelta4), tid);
        if (tracing) tr[1] = clock64();

        for (int rho = 0; rho < nrounds; ++rho) {
            const int ncr = min(TILE_RW, ncols - rho * TILE_RW);
            const int rr0 = warp * CPW, rr1 = rr0 + 1;
            const bool have0 = rr0 < ncr, have1 = rr1 < ncr;
            const int i0 = a.col_begin + c0 + rho * TILE_RW + (have0 ? rr0 : 0);  // global state indices
            const int i1 = a.col_begin + c0 + rho * TILE_RW + (have1 ? rr1 : 0);
            const float tmp0 = __ldg(tmp_row + i0), tmp1 = __ldg(tmp_row + i1);
            float cm0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            float cm1[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            // row offsets inside a full stage, in float4 units; a duplicate row stands in for a missing one
            const int row0 = (have0 ? rr0 : 0) * (TILE_CH >> 2) + lane;
            const int row1 = (have1 ? rr1 : (have0 ? rr0 : 0)) * (TILE_CH >> 2) + lane;
            const float4 *d4 = sdelta4 + lane;
#define FV_ACC2(D, H0, H1)                                                  \
    cm0[0] = fmaxf(cm0[0], __fadd_rn(__fadd_rn(tmp0, (D).x), (H0).x));      \
    cm0[1] = fmaxf(cm0[1], __fadd_rn(__fadd_rn(tmp0, (D).y), (H0).y));      \
    cm0[2] = fmaxf(cm0[2], __fadd_rn(__fadd_rn(tmp0, (D).z), (H0).z));      \
    cm0[3] = fmaxf(cm0[3], __fadd_rn(__fadd_rn(tmp0, (D).w), (H0).w));      \
    cm1[0] = fmaxf(cm1[0], __fadd_rn(__fadd_rn(tmp1, (D).x), (H1).x));      \
    cm1[1] = fmaxf(cm1[1], __fadd_rn(__fadd_rn(tmp1, (D).y), (H1).y));      \
    cm1[2] = fmaxf(cm1[2], __fadd_rn(__fadd_rn(tmp1, (D).z), (H1).z));      \
    cm1[3] = fmaxf(cm1[3], __fadd_rn(__fadd_rn(tmp1, (D).w), (H1).w));
            // Tensor memory and shared memory are read through different pipes (TMEM ~64 B/clk, smem
            // 128 B/clk per SM) and either one alone bounds its part, so odd warps walk the ring stages
            // first and the resident chunks second: both pipes are busy for the whole phase.  The ring
            // holds (almost) a whole step of streamed chunks, so the order does not stall the producer.
            const bool ring_first = TM && (warp & 1);
            const float4 *dring = d4 + nk_res * (TILE_CH >> 2);
            for (int ph = 0; ph < 2; ++ph) {
                const bool tm_now = TM && ((ph == 0) != ring_first);
                if (tm_now) {
                    if (have0) {
                    // the resident chunks: table operands from tensor memory, four iterations per trip so
                    // that one tcgen05.wait covers eight loads; the common full-residency case is unrolled
                    // completely so every tcgen05.ld address is the base plus an immediate
                    if (nk_res == TM_RES_CHUNKS) {
#pragma unroll
                        for (int u = 0; u < TM_RES_CHUNKS * (TILE_CH / 128); u += 4) {
                            float4 ha[4], hb[4], dd[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                ha[e] = tmem_ld4(tbase + 8u * (uint32_t)(u + e));
                                hb[e] = tmem_ld4(tbase + 8u * (uint32_t)(u + e) + 4u);
                                dd[e] = d4[(u + e) * 32];
                            }
                            tmem_wait_ld();
#pragma unroll
                            for (int e = 0; e < 4; ++e) { FV_ACC2(dd[e], ha[e], hb[e]) }
                        }
                    } else {
                        for (int u = 0; u < nk_res * (TILE_CH / 128); u += 2) {
                            const float4 ha0 = tmem_ld4(tbase + 8u * (uint32_t)u), hb0 = tmem_ld4(tbase + 8u * (uint32_t)u + 4u);
                            const float4 ha1 = tmem_ld4(tbase + 8u * (uint32_t)u + 8u), hb1 = tmem_ld4(tbase + 8u * (uint32_t)u + 12u);
                            const float4 da = d4[u * 32], db = d4[u * 32 + 32];
                            tmem_wait_ld();
                            FV_ACC2(da, ha0, hb0)
                            FV_ACC2(db, ha1, hb1)
                        }
                    }
                }
                } else if (TM || ph == 0) {
                    for (int u = nk_res; u < nk; ++u) {
                    const float4 *stage4 = reinterpret_cast<const float4 *>(ring + (size_t)st * STAGE_BYTES);
                    mbar_wait(&full[st], parity);
                    if (have0) {
                        if (u < nk_full) {
#pragma unroll
                            for (int it = 0; it < TILE_CH / 128; ++it) {
                                const float4 d = dring[it * 32];
                                const float4 h0 = stage4[row0 + it * 32];
                                const float4 h1 = stage4[row1 + it * 32];
                                FV_ACC2(d, h0, h1)
                            }
                        } else {  // the short last chunk: rows are len4 float4 apart
                            const int len4 = (a.Kp - u * TILE_CH) >> 2;
                            const float4 *p0 = stage4 + (size_t)rr0 * len4;
                            const float4 *p1 = have1 ? stage4 + (size_t)rr1 * len4 : p0;
                            for (int t = lane; t < len4; t += 32) {
                                const float4 d = dring[t - lane];
                                const float4 h0 = p0[t];
                                const float4 h1 = p1[t];
                                FV_ACC2(d, h0, h1)
                            }
                        }
                    }
                    dring += TILE_CH >> 2;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[st]);
                    if (++st == a.nstage) st = 0, parity ^= 1;
                }
                }
            }
            if (tracing) tr[2] = clock64();
            if (!have0) continue;  // warp-uniform
            const float *round_base = slab + (size_t)rho * TILE_RW * a.Kp;
            Pending q0, q1;
            Scan s0, s1;
            scan_fetch(s0, cm0, round_base, ncr, rr0, a.K, a.Kp, lane);
            if (have1) scan_fetch(s1, cm1, round_base, ncr, rr1, a.K, a.Kp, lane);
            scan_commit(q0, s0, cm0, tmp0, round_base, ncr, rr0, sdelta, a.LAd, a.K, a.Kp, i0, lane);
            if (have1) scan_commit(q1, s1, cm1, tmp1, round_base, ncr, rr1, sdelta, a.LAd, a.K, a.Kp, i1, lane);
            if (tracing) tr[3] = clock64();
            const Best r0 = pending_finish(q0);
            if (lane == 0) {
                if (a.npeer > 1) {
                    if (keep)
                        for (int r = 0; r < a.npeer; ++r)
                            psi_store(a.psi_peer[r], a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i0, r0.k);
                    publish_delta_peers(a, s & 1, i0, r0.x, s, last_step);
                } else {
                    publish_delta(xout, i0, r0.x, (int)step_tag(a, s));
                    if (last_step) a.d_final[i0] = r0.x;
                    if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i0, r0.k);
                }
            }
            if (have1) {
                const Best r1 = pending_finish(q1);
                if (lane == 0) {
                    if (a.npeer > 1) {
                        if (keep)
                            for (int r = 0; r < a.npeer; ++r)
                                psi_store(a.psi_peer[r], a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i1, r1.k);
                        publish_delta_peers(a, s & 1, i1, r1.x, s, last_step);
                    } else {
                        publish_delta(xout, i1, r1.x, (int)step_tag(a, s));
                        if (last_step) a.d_final[i1] = r1.x;
                        if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i1, r1.k);
                    }
                }
            }
        }
        if (tracing) tr[4] = clock64();
        named_bar_sync(1, NCONS);  // sdelta is overwritten by the next step's load
        if (tracing) tr[5] = clock64();
    }
#undef FV_ACC2
    if (a.npeer > 1) {
        // Final hand-shake of a sharded pass: wait until every GPU's last delta slice has arrived
        // here (their release stores order their backpointer entries before it), acquire, and
        // leave the complete final vector where k_flash_end expects it.
        delta_wait_load(a, a.nsteps + 1, reinterpret_cast<float *>(sdelta4), tid);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (b == 0)
            for (int i = tid; i < a.K; i += NCONS) a.d_final[i] = sdelta[i];
    }
    if (TM) {
        tmem_fence_before();
        named_bar_sync(1, NCONS);
        if (warp == 0) tmem_dealloc(*tmem_slot);
    }
}

// ---- the whole table slice on chip ---------------------------------------------------------------
// k_flash_resident: the same pass for models small enough that a CTA's slice of the float table
// (~27 columns x Kp floats, 428 KB at K=3965 on 148 SMs) fits the SM's on-chip memories TOGETHER:
// source states [0, S) of every owned column live in tensor memory (S <= 2048: 4 columns x 2048 x 4 B
// per warp, 224 KB), states [S, Kp) in shared memory (207 KB at K=3965).  Both are loaded once; after
// that a step reads no table byte from L2 — the ring, its TMA refill traffic (which cost as much
// shared-memory bandwidth as the reads themselves) and the per-stage mbarrier traffic are gone, and
// so is the L2 round trip of the window scan: the winning chain is re-read from the on-chip copy.
//
// Work split: warp pair p (warps p and p+RP) owns the four columns 4p..4p+3 of the CTA.  Warp p
// sweeps the tensor-memory half of the source axis for all four (64 B/clk read path), warp p+RP the
// shared-memory half (128 B/clk) — both pipes busy for the whole phase, every delta value read meets
// four columns (half the shared-memory delta traffic of the two-column layout).  The halves meet in
// shared memory twice: once for the column maxima (window threshold), once for the exact results.
constexpr int RP = 7;                 // warp pairs
constexpr int RCOLS = 4;              // columns per pair
constexpr int RES_MAX_COLS = RP * RCOLS;
static_assert(RES_MAX_COLS == TILE_RW, "the resident kernel uses the tiling of the ring kernel (one round of TILE_RW columns)");
constexpr int RES_TM_STATES = 2048;   // 4 columns x 2048 states x 4 B = 256 tensor-memory columns per warp

struct PairSlot {  // what the two warps of a pair hand each other, per column
    float top[2][RCOLS];
    float bx[2][RCOLS];
    int bk[2][RCOLS];
};

// One candidate chain of a tensor-memory warp: the 16 elements k = 4*(w+32u)+c of column j sit in TMEM
// lane w, so lane w walks them alone (every lane executes the loads; only lane w's values matter).
template <int NU>
__device__ __forceinline__ void rescan_tmem(Pending &p, uint32_t tbase, int j, int w, int c, int nu, float tmp, int thr,
                                            const float *sdelta, const double *__restrict__ LAd, int K, int i, int lane)
{
    float v[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        uint32_t x;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(tbase + 16u * (uint32_t)u + 4u * (uint32_t)j + (uint32_t)c) : "memory");
        v[u] = __uint_as_float(x);
    }
    tmem_wait_ld();
    if (lane == w) {
#pragma unroll
        for (int u = 0; u < NU; ++u)
            if (u < nu) {
                const int k = 4 * (w + 32 * u) + c;
                if (k < K) {
                    const float pre = __fadd_rn(tmp, sdelta[k]);
                    if (ford(__fadd_rn(pre, v[u])) >= thr) pending_push(p, pre, k, LAd + (size_t)k * K + i);
                }
            }
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_flash_resident(const PersistArgs a, int S)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Kp4 = a.Kp >> 2;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(full + 2 * MAX_STAGES);
    float4 *sdelta4 = reinterpret_cast<float4 *>(smem_raw + CTRL_BYTES);
    PairSlot *spair = reinterpret_cast<PairSlot *>(sdelta4 + Kp4);
    float *srest = reinterpret_cast<float *>(spair + RP);  // [chunk][row][state] exactly as the tiled table stores it

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int c0 = tile_c0(a.ncol, G, b), ncols = tile_c0(a.ncol, G, b + 1) - c0;  // <= RES_MAX_COLS: one round
    const float *slab = a.hiC + (size_t)c0 * a.Kp;
    const int rest_states = a.Kp - S;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    if (warp == 2 * RP) {
        // loader: the shared-memory half of the slice, once; the copy engine runs while the other warps
        // park their half in tensor memory
        if (lane == 0 && rest_states > 0) {
            const uint32_t total = (uint32_t)rest_states * (uint32_t)ncols * 4u;
            mbar_expect_tx(&full[0], total);
            const unsigned char *src = reinterpret_cast<const unsigned char *>(slab) + (size_t)S * ncols * 4;
            unsigned char *dst = reinterpret_cast<unsigned char *>(srest);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t n = total - off < 32768u ? total - off : 32768u;
                bulk_g2s(dst + off, src + off, n, &full[0], 0, false);
            }
        }
        return;
    }

    // ---------------- consumers: warps 0..RP-1 tensor-memory half, RP..2RP-1 shared-memory half ------
    const bool tm_half = warp < RP;
    const int pair = tm_half ? warp : warp - RP, half = tm_half ? 0 : 1;
    const float *sdelta = reinterpret_cast<const float *>(sdelta4);
    const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16) + 256u * (uint32_t)(warp >> 2);
    const int nu_t = S >> 7;            // 128-state iterations of the tensor-memory half (<= 16)
    const int nu_s = rest_states >> 7;  // ... of the shared-memory half
    int rr[RCOLS];                      // rows (local columns) of this pair; a missing one is a duplicate of the pair's first
    bool have[RCOLS];
#pragma unroll
    for (int j = 0; j < RCOLS; ++j) {
        have[j] = pair * RCOLS + j < ncols;
        rr[j] = have[j] ? pair * RCOLS + j : (pair * RCOLS < ncols ? pair * RCOLS : 0);
    }
    const bool pair_live = pair * RCOLS < ncols;  // warp-uniform
    if (tm_half && pair_live) {
        for (int u = 0; u < nu_t; ++u) {
            const int k = 4 * (lane + 32 * u);
#pragma unroll
            for (int j = 0; j < RCOLS; ++j)
                tmem_st4(tbase + 16u * (uint32_t)u + 4u * (uint32_t)j,
                         __ldg(reinterpret_cast<const float4 *>(slab + tile_round_off(a.Kp, ncols, rr[j], k))));
        }
        tmem_wait_st();
    }
    tmem_fence_before();
    named_bar_sync(1, NCONS);
    tmem_fence_after();
    if (rest_states > 0) mbar_wait(&full[0], 0);

    // address of (row, state k >= S) in srest: chunks of TILE_CH states, the last one possibly shorter
    auto rest_ptr = [&](int row, int k) -> const float * {
        const int kr = k - S, cu = kr / TILE_CH, len = min(TILE_CH, rest_states - cu * TILE_CH);
        return srest + (size_t)cu * TILE_CH * ncols + (size_t)row * len + (kr - cu * TILE_CH);
    };

    for (int s = 1; s <= a.nsteps; ++s) {
        unsigned long long *xout = a.xch + (size_t)(s & 1) * a.Kp;
        const bool last_step = s == a.nsteps;
        const int jstep = a.L + s;
        const float *tmp_row = a.LBf + (size_t)__ldg(a.ob + jstep) * a.Kp;  // F:167
        const bool keep = jstep >= a.mid + 1;                                // F:242
        const bool tracing = a.trace != nullptr && s <= TRACE_STEPS && lane == 0 && (warp == 0 || warp == 2 * RP - 1);
        long long *tr = tracing ? a.trace + ((((size_t)(s - 1) * G + b) * 2 + (warp == 0 ? 0 : 1)) * TRACE_PTS) : nullptr;
        if (tracing) tr[0] = clock64();
        delta_wait_load(a, s, reinterpret_cast<float *>(sdelta4), tid);
        if (tracing) tr[1] = clock64();
        if (!pair_live) {  // warp-uniform: this pair owns no column of the CTA (only when ncols < 4*RP - 3)
            named_bar_sync(1, NCONS);
            continue;
        }
        int icol[RCOLS];
        float tmp[RCOLS];
        float cm[RCOLS][4];
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            icol[j] = a.col_begin + c0 + rr[j];
            tmp[j] = __ldg(tmp_row + icol[j]);
            cm[j][0] = cm[j][1] = cm[j][2] = cm[j][3] = -INFINITY;
        }
#define FV_ACC4(D, H)                                                                    \
    _Pragma("unroll") for (int j = 0; j < RCOLS; ++j) {                                  \
        cm[j][0] = fmaxf(cm[j][0], __fadd_rn(__fadd_rn(tmp[j], (D).x), (H)[j].x));        \
        cm[j][1] = fmaxf(cm[j][1], __fadd_rn(__fadd_rn(tmp[j], (D).y), (H)[j].y));        \
        cm[j][2] = fmaxf(cm[j][2], __fadd_rn(__fadd_rn(tmp[j], (D).z), (H)[j].z));        \
        cm[j][3] = fmaxf(cm[j][3], __fadd_rn(__fadd_rn(tmp[j], (D).w), (H)[j].w));        \
    }
        if (tm_half) {
            const float4 *d4 = sdelta4 + lane;
            if (nu_t == RES_TM_STATES / 128) {
#pragma unroll
                for (int u = 0; u < RES_TM_STATES / 128; u += 2) {  // two iterations per tcgen05.wait: eight loads in flight
                    float4 h0[RCOLS], h1[RCOLS];
#pragma unroll
                    for (int j = 0; j < RCOLS; ++j) {
                        h0[j] = tmem_ld4(tbase + 16u * (uint32_t)u + 4u * (uint32_t)j);
                        h1[j] = tmem_ld4(tbase + 16u * (uint32_t)(u + 1) + 4u * (uint32_t)j);
                    }
                    const float4 da = d4[u * 32], db = d4[(u + 1) * 32];
                    tmem_wait_ld();
                    FV_ACC4(da, h0)
                    FV_ACC4(db, h1)
                }
            } else {
                for (int u = 0; u < nu_t; ++u) {
                    float4 h0[RCOLS];
#pragma unroll
                    for (int j = 0; j < RCOLS; ++j) h0[j] = tmem_ld4(tbase + 16u * (uint32_t)u + 4u * (uint32_t)j);
                    const float4 da = d4[u * 32];
                    tmem_wait_ld();
                    FV_ACC4(da, h0)
                }
            }
        } else {
            const float4 *d4 = sdelta4 + (S >> 2) + lane;
            const float4 *r4 = reinterpret_cast<const float4 *>(srest);
            const int nfull = rest_states / TILE_CH;  // whole chunks; at most one shorter chunk (128 states... any multiple of 128) follows
            for (int cu = 0; cu < nfull; ++cu) {
                const float4 *ch = r4 + (size_t)cu * (TILE_CH >> 2) * ncols + lane;
#pragma unroll
                for (int it = 0; it < TILE_CH / 128; ++it) {
                    float4 h0[RCOLS];
#pragma unroll
                    for (int j = 0; j < RCOLS; ++j) h0[j] = ch[rr[j] * (TILE_CH >> 2) + it * 32];
                    const float4 da = d4[(cu * (TILE_CH / 128) + it) * 32];
                    FV_ACC4(da, h0)
                }
            }
            const int tail = rest_states - nfull * TILE_CH;  // states of the short last chunk
            if (tail > 0) {
                const float4 *ch = r4 + (size_t)nfull * (TILE_CH >> 2) * ncols + lane;
                const int len4 = tail >> 2;
                for (int it = 0; it < tail / 128; ++it) {
                    float4 h0[RCOLS];
#pragma unroll
                    for (int j = 0; j < RCOLS; ++j) h0[j] = ch[rr[j] * len4 + it * 32];
                    const float4 da = d4[(nfull * (TILE_CH / 128) + it) * 32];
                    FV_ACC4(da, h0)
                }
            }
        }
#undef FV_ACC4
        if (tracing) tr[2] = clock64();

        // ---- column maxima of both halves -> window thresholds --------------------------------------
        PairSlot &ps = spair[pair];
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            const float t = warp_max(fmaxf(fmaxf(cm[j][0], cm[j][1]), fmaxf(cm[j][2], cm[j][3])));
            if (lane == 0) ps.top[half][j] = t;
        }
        named_bar_sync(2 + pair, 64);
        int thr[RCOLS];
        bool dead[RCOLS];
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            const float top = fmaxf(ps.top[0][j], ps.top[1][j]);
            dead[j] = !(top > -FLT_MAX);
            thr[j] = ford(top) - WINDOW_STEPS;
        }
        // ---- chains inside the window: re-read them from the on-chip copy, start the exact loads -----
        Pending q[RCOLS];
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            q[j].acc = Best{-FLT_MAX, 0x7fffffff};
            q[j].has = false, q[j].la = 0.0, q[j].pre = 0.f, q[j].k = 0;
            if (dead[j] || !have[j]) continue;  // warp-uniform
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                unsigned hit = __ballot_sync(FULL_MASK, ford(cm[j][c]) >= thr[j]);
                while (hit) {
                    const int w = __ffs(hit) - 1;
                    hit &= hit - 1;
                    if (tm_half) {
                        rescan_tmem<RES_TM_STATES / 128>(q[j], tbase, j, w, c, nu_t, tmp[j], thr[j], sdelta, a.LAd, a.K, icol[j], lane);
                    } else {
                        for (int u = lane; u < nu_s; u += 32) {
                            const int k = S + 4 * (w + 32 * u) + c;
                            if (k < a.K) {
                                const float pre = __fadd_rn(tmp[j], sdelta[k]);
                                if (ford(__fadd_rn(pre, *rest_ptr(rr[j], k))) >= thr[j])
                                    pending_push(q[j], pre, k, a.LAd + (size_t)k * a.K + icol[j]);
                            }
                        }
                    }
                }
            }
        }
        if (tracing) tr[3] = clock64();
        // ---- exact (value, first index) of each half, then of the pair ------------------------------
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            const Best r = pending_finish(q[j]);  // (-FLT_MAX, -1) when this half holds no candidate
            if (lane == 0) ps.bx[half][j] = r.x, ps.bk[half][j] = r.k;
        }
        named_bar_sync(2 + pair, 64);
        // the tensor-memory warp publishes columns 0 and 1, the shared-memory warp 2 and 3: lanes 0/1 one each
#pragma unroll
        for (int j = 0; j < RCOLS; ++j) {
            if ((j >> 1) != half || (j & 1) != lane || !have[j]) continue;
            Best r{ps.bx[0][j], ps.bk[0][j]};
            if (r.k < 0) r.k = 0x7fffffff;
            const int k1 = ps.bk[1][j];
            best_take(r, ps.bx[1][j], k1 < 0 ? 0x7fffffff : k1);
            if (!(r.x > -FLT_MAX)) r.x = -FLT_MAX, r.k = -1;
            const int i = icol[j];
            publish_delta(xout, i, r.x, (int)step_tag(a, s));
            if (last_step) a.d_final[i] = r.x;
            if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (jstep - a.mid - 1)) * a.K + i, r.k);
        }
        if (tracing) tr[4] = clock64();
        named_bar_sync(1, NCONS);  // sdelta is overwritten by the next step's load; the pair slots by the next step
        if (tracing) tr[5] = clock64();
    }
    tmem_fence_before();
    named_bar_sync(1, NCONS);
    if (warp == 0) tmem_dealloc(*tmem_slot);
}

// ---- host side ---------------------------------------------------------------------------------
static size_t persist_smem(int Kp, int nstage)
{
    return CTRL_BYTES + (size_t)Kp * 4 + (size_t)nstage * TILE_RW * TILE_CH * 4;
}

// Largest padded K whose delta vector plus two ring stages fit the shared memory of one CTA.
bool persistent_engine_fits(const flashv_ctx *ctx, int Kp)
{
    return ctx->coop && persist_smem(Kp, 2) <= (size_t)ctx->smem_optin;
}

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int launch_persist(flashv_model *m, PersistArgs &a)
{
    flashv_ctx *ctx = m->ctx;
    const int Kp = a.Kp;
    const size_t fixed = persist_smem(Kp, 0);
    if (fixed + 2 * (size_t)TILE_RW * TILE_CH * 4 > (size_t)ctx->smem_optin) {
        set_error("persistent engine: K=%d does not fit shared memory (%d bytes)", a.K, ctx->smem_optin);
        return FLASHV_ERR_ARG;
    }
    int nstage = (int)(((size_t)ctx->smem_optin - fixed) / ((size_t)TILE_RW * TILE_CH * 4));
    if (nstage > MAX_STAGES) nstage = MAX_STAGES;
    const int cap = env_int("FLASHV_STAGES", 0);
    if (cap >= 2 && cap < nstage) nstage = cap;
    a.nstage = nstage;
    a.l2_hint = env_int("FLASHV_L2_HINT", 1);
    size_t smem = persist_smem(Kp, nstage);
    // tensor-memory residency needs every CTA to own at most one round of columns
    const int grid = ctx->sm_count < a.ncol ? ctx->sm_count : a.ncol;  // the grid the tiled table was laid out for
    const int cols_max = (a.ncol + grid - 1) / grid;
    const bool use_tmem = cols_max <= TILE_RW && Kp >= TILE_CH && env_int("FLASHV_TMEM", 1) != 0;
    // the whole slice on chip (k_flash_resident): single GPU, one round of columns, and the part of the slice
    // that tensor memory cannot take must fit shared memory beside the delta vector
    int res_S = Kp / TILE_CH * TILE_CH;
    if (res_S > RES_TM_STATES) res_S = RES_TM_STATES;
    const size_t res_smem = CTRL_BYTES + (size_t)Kp * 4 + RP * sizeof(PairSlot) + (size_t)(Kp - res_S) * cols_max * 4;
    const bool use_res = a.npeer == 1 && cols_max <= RES_MAX_COLS && res_smem <= (size_t)ctx->smem_optin &&
                         env_int("FLASHV_RESIDENT", 1) != 0;
    if (use_res) smem = res_smem;
    const void *fn = use_res ? (const void *)k_flash_resident
                             : (use_tmem ? (const void *)k_flash_persist<true> : (const void *)k_flash_persist<false>);
    FV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NTHREADS, smem));
    if (per_sm < 1) {
        set_error("persistent engine: kernel does not fit one CTA per SM");
        return FLASHV_ERR_CUDA;
    }
    // developer aid: FLASHV_TRACE_FILE=<path> dumps per-step phase timestamps of the last launch
    const char *trace_path = a.nsteps >= TRACE_STEPS ? getenv("FLASHV_TRACE_FILE") : nullptr;
    static long long *d_trace = nullptr;
    const size_t trace_n = (size_t)TRACE_STEPS * grid * 2 * TRACE_PTS;
    a.trace = nullptr;
    if (trace_path) {
        if (!d_trace) FV_CUDA(cudaMalloc(&d_trace, trace_n * sizeof(long long)));
        FV_CUDA(cudaMemsetAsync(d_trace, 0, trace_n * sizeof(long long), ctx->stream));
        a.trace = d_trace;
    }
    void *params[] = {(void *)&a, (void *)&res_S};  // the ring kernels take the first argument only
    // cooperative launch: every CTA polls data the others produce, so all must be co-resident;
    // the grid is the one the tiled table was laid out for
    FV_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NTHREADS), params, smem, ctx->stream));
    if (trace_path) {
        std::vector<long long> h(trace_n);
        FV_CUDA(cudaMemcpyAsync(h.data(), d_trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaStreamSynchronize(ctx->stream));
        if (FILE *fp = fopen(trace_path, "wb")) {
            const int hdr[4] = {TRACE_STEPS, grid, 2, TRACE_PTS};
            fwrite(hdr, sizeof(hdr), 1, fp);
            fwrite(h.data(), sizeof(long long), trace_n, fp);
            fclose(fp);
        }
    }
    return FLASHV_OK;
}

static unsigned long long watchdog_limit_ns(bool cross_gpu)
{
    // A rank that launches late (first-call module load, a slow host) must not trap its peers.
    const int ms = env_int("FLASHV_WATCHDOG_MS", cross_gpu ? 30000 : 4000);
    return (unsigned long long)(ms < 100 ? 100 : ms) * 1000000ull;
}

int persistent_pass(flashv_plan *p, const Pass &pass)
{
    flashv_model *m = p->model;
    const VecDesc &vd = pass.first_vec;  // the pass has exactly one vector (batch == 1)
    PersistArgs a;
    a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.ob = p->d_ob;
    a.L = vd.L, a.nsteps = vd.R - vd.L, a.mid = vd.mid, a.psi_row = vd.psi_row;
    a.d_init = p->d_delta, a.d_final = p->d_delta + (size_t)p->max_vec * m->Kp;
    a.psi16 = p->psi16;
    // Only the plan's pass 0 is sharded (the N-way pass, or the root task): it is the one long
    // single-vector pass.  It uses the plan's shard region — exchange words and backpointer rows that
    // no local pass touches — and tags its words with the run count, which is the same on every rank.
    if (pass_is_sharded(p, pass)) {
        // state-sharded: this GPU owns the columns [shard_c0, shard_c0 + shard_ncol) and publishes
        // into every GPU's region (same offsets in every plan: identical plan parameters)
        if (a.nsteps >= 65536) {
            set_error("state-sharded pass: more than 65535 steps");
            return FLASHV_ERR_ARG;
        }
        a.epoch = p->shard_run % 65535u + 1u;
        a.hiC = p->hiC_shard, a.col_begin = p->shard_c0, a.ncol = p->shard_ncol, a.npeer = p->shard_world;
        a.xch = reinterpret_cast<unsigned long long *>(p->shard_region);
        a.psi = p->shard_region + p->shard_psi_off;
        for (int r = 0; r < p->shard_world; ++r) {
            a.xch_peer[r] = reinterpret_cast<unsigned long long *>(p->peer_region[r]);
            a.psi_peer[r] = p->peer_region[r] + p->shard_psi_off;
        }
        a.watchdog_ns = watchdog_limit_ns(true);
    } else {
        a.epoch = (++p->run_epoch) & 0xffffu;
        if (a.epoch == 0) a.epoch = (++p->run_epoch) & 0xffffu;  // tag 0 is what a fresh buffer holds
        a.xch = reinterpret_cast<unsigned long long *>(p->d_delta + (size_t)2 * p->max_vec * m->Kp);
        a.psi = p->d_psi;
        a.hiC = m->hiC, a.col_begin = 0, a.ncol = m->K, a.npeer = 1;
        a.xch_peer[0] = a.xch, a.psi_peer[0] = a.psi;
        a.watchdog_ns = watchdog_limit_ns(false);
    }
    int rc = launch_persist(m, a);
    if (rc == FLASHV_OK) p->launches += 1;
    return rc;
}

// ---- task-tree levels of a state-sharded plan: every rank ran its share of the level's tasks; now every
// rank needs all of the level's Ans entries (the next level restarts from Ans[L-1] and ends in Ans[R],
// F:220, F:248).  Same idea as the delta exchange: each entry travels as ONE self-validating 64-bit
// word {Ans value, run|level tag}, stored by its owner into every rank's region and polled locally.
struct AnsXchArgs {
    const int32_t *mids;  // midpoints of the level's tasks, all ranks', in level order; task t belongs to rank t % world
    int cnt, rank, world;
    int32_t *ans;
    unsigned tag;
    unsigned long long *xch_peer[8];  // every rank's Ans exchange words [T] (entry `rank` is the local one)
    unsigned long long watchdog_ns;
};

__global__ void __launch_bounds__(256) k_ans_exchange(const AnsXchArgs a)
{
    for (int t = threadIdx.x; t < a.cnt; t += blockDim.x)
        if (t % a.world == a.rank) {
            const int mid = a.mids[t];
            const unsigned long long w = ((unsigned long long)a.tag << 32) | (unsigned long long)(unsigned)a.ans[mid];
            for (int r = 0; r < a.world; ++r) asm volatile("st.global.u64 [%0], %1;" ::"l"(a.xch_peer[r] + mid), "l"(w) : "memory");
        }
    const unsigned long long *mine = a.xch_peer[a.rank];
    for (int t = threadIdx.x; t < a.cnt; t += blockDim.x) {
        const int mid = a.mids[t];
        unsigned long long w, t0 = 0;
        for (uint32_t spins = 0;; ++spins) {
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(mine + mid) : "memory");
            if ((unsigned)(w >> 32) == a.tag) break;
            watchdog(spins * 16u + 15u, t0, a.watchdog_ns);
        }
        a.ans[mid] = (int32_t)(unsigned)w;
    }
}

int shard_ans_exchange(flashv_plan *p, int level)
{
    if (level < 0 || level >= (int)p->lvl_cnt.size()) return FLASHV_OK;
    AnsXchArgs a;
    a.mids = p->d_lvl_mid + p->lvl_off[level], a.cnt = p->lvl_cnt[level];
    a.rank = p->shard_rank, a.world = p->shard_world, a.ans = p->d_ans;
    a.tag = ((p->shard_run & 0x3ffffffu) << 6) | (unsigned)(level + 1);
    for (int r = 0; r < p->shard_world; ++r)
        a.xch_peer[r] = reinterpret_cast<unsigned long long *>(p->peer_region[r] + p->shard_ans_off);
    a.watchdog_ns = watchdog_limit_ns(true);
    k_ans_exchange<<<1, 256, 0, p->model->ctx->stream>>>(a);
    FV_CUDA(cudaGetLastError());
    ++p->launches;
    return FLASHV_OK;
}

// ---- state sharding: this GPU's column slice of the tiled table ------------------------
int shard_build_table(flashv_plan *p)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    const int K = m->K, Kp = m->Kp;
    p->shard_c0 = (int)((long long)p->shard_rank * K / p->shard_world);
    p->shard_ncol = (int)((long long)(p->shard_rank + 1) * K / p->shard_world) - p->shard_c0;
    if (p->shard_ncol < 1) {
        set_error("state sharding: rank %d of %d owns no column of K=%d", p->shard_rank, p->shard_world, K);
        return FLASHV_ERR_ARG;
    }
    const int grid = ctx->sm_count < p->shard_ncol ? ctx->sm_count : p->shard_ncol;
    FV_CUDA(cudaMalloc(&p->hiC_shard, (size_t)p->shard_ncol * Kp * sizeof(float)));
    p->bytes += (size_t)p->shard_ncol * Kp * sizeof(float);
    build_tiled_slice(m->LAd, p->hiC_shard, K, Kp, p->shard_c0, p->shard_ncol, grid, ctx->stream);
    FV_CUDA(cudaGetLastError());
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    return FLASHV_OK;
}

int persistent_single_step(flashv_model *m, const float *d_in_dev, int o, float *d_out_dev, int32_t *psi_dev)
{
    flashv_ctx *ctx = m->ctx;
    int32_t *dob = m->scratch_i + 16;
    int32_t hob[2] = {o, o};
    FV_CUDA(cudaMemcpyAsync(dob, hob, sizeof(hob), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    PersistArgs a;
    a.hiC = m->hiC, a.col_begin = 0, a.ncol = m->K, a.npeer = 1;
    a.LAd = m->LAd, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
    a.ob = dob, a.L = 0, a.nsteps = 1, a.mid = 0, a.psi_row = 0;
    static unsigned hook_epoch = 0;
    a.epoch = (++hook_epoch & 0xffffu) ? (hook_epoch & 0xffffu) : (++hook_epoch & 0xffffu);
    a.d_init = d_in_dev, a.d_final = d_out_dev;
    a.xch = reinterpret_cast<unsigned long long *>(m->scratch_x);
    a.psi = psi_dev, a.psi16 = 0;
    a.xch_peer[0] = a.xch, a.psi_peer[0] = a.psi;
    a.watchdog_ns = watchdog_limit_ns(false);
    return launch_persist(m, a);
}

}  // namespace flashv
