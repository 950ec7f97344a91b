// flash_persistent.cu — engine PERSISTENT: a whole single-vector FLASH pass (the N-way first pass
// of nvviterNdivide, F:126-202, or a full-length task of nvviter, F:204-262) in ONE cooperative
// launch, one CTA per SM.  Two kernels:
//
//   k_flash_persist16 (models up to 4096 states, unsharded — the headline pass; second half of this file)
//     the CTA's whole slice of (half)log A lives in tensor memory for the launch; the sweep is a half-precision
//     filter (HADD2 + HMNMX2 per two updates) with a rigorous window, and every chain inside the window is
//     evaluated exactly from a chain-major double table.  See the comment above the kernel and DESIGN.md §4.
//
//   k_flash_persist<TM, PIN, PEERS> (wider models, state-sharded passes, FLASHV_F16=0)
//   * Each CTA owns a contiguous range of destination columns (K/gridDim of them) and re-reads their
//     float log-A entries every step.  The table is stored CTA-tiled (hiC, tile_geom.h): per CTA, per chunk of
//     TILE_CH source states, the chunk of every owned column back to back, so one step of one CTA is one linear
//     stream and every ring stage is one contiguous bulk copy.  Wide models stream it from HBM at the roofline
//     (K=16384: 6.38 TB/s); single-round CTAs keep the first 2048 states of every column in tensor memory (TM)
//     and, when the rest fits shared memory whole, load it once (PIN).
//   * Warp NCW is the producer: one lane streams the slab through an NSTAGE-deep shared-memory ring
//     with bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx), running ahead across step
//     boundaries since the table does not change — the ring refills while the CTAs hand over.
//   * Warps 0..NCW-1 are consumers and ALL of them work on every stage: warp w owns CPW fixed
//     columns of the round (rows 2w, 2w+1 of the stage), lane l owns k = 4*(l+32u)+c and keeps four
//     running maxima per column of the float estimate (FADD, FADD, FMNMX per update); a stage is
//     released after NCW arrivals, so the ring is a true FIFO and no warp ever holds a stage longer
//     than one pass over its two rows.  The exact (value, first index) comes from a chain-major double
//     table for the chains inside the window (scan_fetch / scan_settle / scan_long below).
//   * Steps hand over through the delta vector itself ({value, step} words polled from L2, see
//     delta_wait_load): no grid barrier, no atomics.  PEERS: the words (and the backpointer rows) go to
//     every GPU of a state-sharded pass with in-kernel peer stores.
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cuda_fp16.h>

#include "flashv_internal.h"
#include "tile_geom.h"
#include "trellis_common.cuh"
#include "half_filter.cuh"

namespace flashv {

constexpr int NCW = TILE_RW / 2;      // consumer warps (14)
constexpr int CPW = 2;                // columns a consumer warp owns per round
constexpr int NCONS = NCW * 32;       // consumer threads
constexpr int NTHREADS = NCONS + 32;  // + producer warp
constexpr int MAX_STAGES = 64;
constexpr int TRACE_STEPS = 64, TRACE_PTS = 7;
constexpr int CTRL_BYTES = 2 * MAX_STAGES * 8 + 64;  // full[], empty[], the TMEM base address slot

// ---- PTX wrappers -------------------------------------------------------------------------

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
constexpr int TM_RES_CHUNKS = 8;   // TILE_CH-sized chunks of every owned column kept resident (2048 states)
// 32 consecutive tensor-memory columns of this thread's lane in ONE instruction: four 128-state iterations of
// both columns a warp owns (layout below: columns 8u..8u+3 first column, 8u+4..8u+7 second).  A tcgen05.ld
// costs the tensor-memory read path a fixed ~8 cycles plus its bytes at ~54 B/clk (measured: 448 .x4 loads
// per step took 4.0 us = 17.5 cycles each, 112 .x16 loads 46 cycles each), so the narrow form that fetched
// one float4 per instruction spent a third of the phase on per-instruction cost.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float4 (&ha)[4], float4 (&hb)[4])
{
    uint32_t r[32];
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
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ha[e] = make_float4(__uint_as_float(r[8 * e]), __uint_as_float(r[8 * e + 1]), __uint_as_float(r[8 * e + 2]), __uint_as_float(r[8 * e + 3]));
        hb[e] = make_float4(__uint_as_float(r[8 * e + 4]), __uint_as_float(r[8 * e + 5]), __uint_as_float(r[8 * e + 6]), __uint_as_float(r[8 * e + 7]));
    }
}

struct PersistArgs {
    const float *hiC;  // CTA-tiled (float)log A of the columns [col_begin, col_begin+ncol), tile_geom.h
    int col_begin, ncol;  // destination columns this GPU owns (all of them unless the pass is state-sharded)
    const double *LAc;  // chain-major double table of models up to 4096 states (tables.cu), else null
    const double *LAcL;  // the same for wider models: 128 chains of `clp` elements (Kp/128 rounded up to 32), else null
    int clp;
    // the half-precision filter (k_flash_persist16): models up to 4096 states, unsharded passes
    const __half *hi16;   // (half)log A, CTA-tiled [cta][iteration of 256 states][column][256], -inf padding
    const double *LAc16;  // log A chain-major for 256 chains of 16: [(i * 256 + (k & 255)) * 16 + (k >> 8)]
    int Kp16;             // K rounded up to 256
    int tm_iters;         // sweep iterations (of 256 states) whose table operands live in tensor memory
    const float *LBmax;   // [M] the largest emission term of every symbol
    int exact_max;        // 1: find the maximum of every staged vector exactly (one more CTA barrier per step) instead of bounding it
    double lamax;         // the largest log A of the model
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
    int pinned;  // 1: the streamed part of every CTA's slice fits the ring whole — it is loaded once and never refilled
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
#pragma unroll 1
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

// Window scan of one column.  The sweep leaves every lane with the float-estimate maxima of four "chains":
// chain q = k & 127 holds the source states k = q + 128 u, u = 0 .. Kp/128 - 1 (lane w, accumulator c  <->
// q = 4 w + c).  The winner can only sit in a chain whose maximum lies inside the window (WINDOW_STEPS float
// steps below the column's best estimate), which is one chain in almost every column, sometimes two.
//
// For K <= 4096 a chain has at most 32 elements and the model carries the chain-major double table LAc
// (tables.cu): LAc[(i * 128 + q) * 32 + u] = log A[q + 128 u][i], -inf beyond K.  scan_fetch() has lane u
// load element u of each chain inside the window — ONE coalesced 256-byte request per chain and column —
// and scan_settle() derives both the float estimate ((float)la, the very number the sweep used) and the exact
// candidate from it.  (Before the chain table the estimates were re-read from the tiled float table, 31
// different lines per chain, and the doubles of the survivors fetched from LAd afterwards: two dependent
// round trips, the first of them 31 requests wide.  See profiles/README.md, round 2.)
//
// Columns with more chains inside the window than SCAN_SLOTS, and all columns of wider models, go through
// scan_slow(), which walks the chains synchronously.
constexpr int SCAN_SLOTS = 2;
constexpr int CHAIN_PAD = 32;   // elements per chain in LAc
struct Scan {
    double la0, la1;  // slot s holds element `lane` of the s-th chain inside the window (-inf: none)
    int q0, q1;       // the chains, warp-uniform; NO_CHAIN if the slot is empty
    float thr;        // window threshold, warp-uniform
    bool live;        // warp-uniform: the column exists and has a finite estimate
    bool overflow;    // warp-uniform: the column goes through scan_slow()
};


// Straight-line on purpose (no early exits, reductions by redux.sync): the two columns of a warp are scanned
// back to back and the compiler interleaves them, which halves this latency-bound stretch of the step.
__device__ __forceinline__ void scan_fetch(Scan &sc, const float (&cm)[4], const double *__restrict__ LAc, int i,
                                           bool have, int lane)
{
    const int top = __reduce_max_sync(FULL_MASK, ford(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]))));
    sc.live = have && top > ford(-FLT_MAX);
    // WINDOW_STEPS float steps below the best estimate, as a float again (clamped at -inf, where every chain is
    // inside the window and the column goes the slow way): comparing floats equals comparing their integer images
    sc.thr = unford(max(top - WINDOW_STEPS, ford(-INFINITY)));
    unsigned m4 = (cm[0] >= sc.thr ? 1u : 0u) | (cm[1] >= sc.thr ? 2u : 0u) | (cm[2] >= sc.thr ? 4u : 0u) |
                  (cm[3] >= sc.thr ? 8u : 0u);  // this lane's chains inside the window
    int lq = m4 ? 4 * lane + __ffs(m4) - 1 : NO_CHAIN;
    sc.q0 = __reduce_min_sync(FULL_MASK, lq);
    if (lq == sc.q0) m4 &= m4 - 1;
    lq = m4 ? 4 * lane + __ffs(m4) - 1 : NO_CHAIN;
    sc.q1 = __reduce_min_sync(FULL_MASK, lq);
    if (lq == sc.q1) m4 &= m4 - 1;
    sc.overflow = LAc == nullptr || __any_sync(FULL_MASK, m4 != 0);
    sc.la0 = sc.la1 = -INFINITY;
    if (sc.live && !sc.overflow) {
        // the table pads every chain to CHAIN_PAD elements with -inf, so all lanes load; q0 is a chain whenever
        // the column is live (the best estimate itself lies inside the window)
        const double *col = LAc + (size_t)i * (128 * CHAIN_PAD) + lane;
        sc.la0 = __ldg(col + sc.q0 * CHAIN_PAD);
        if (sc.q1 != NO_CHAIN) sc.la1 = __ldg(col + sc.q1 * CHAIN_PAD);
    }
}

// The rare column that does not fit the slots (and every column of a model wider than 4096 states): every
// chain inside the window, element by element, estimates from the tiled float table, each exact value loaded
// on the spot.  Kept out of line: inlined twice per step it only stretched the code the common path fetches.
__device__ __noinline__ Best scan_slow(float cm0, float cm1, float cm2, float cm3, int thr, float tmp,
                                       const float *__restrict__ round_base, int ncr, int rr, const float *sdelta,
                                       const double *__restrict__ LAd, int K, int Kp, int i, int lane)
{
    Best acc{-FLT_MAX, 0x7fffffff};
    const int chain_len = Kp >> 7;
    const float cm[4] = {cm0, cm1, cm2, cm3};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned hit = __ballot_sync(FULL_MASK, ford(cm[c]) >= thr);
        while (hit) {
            const int w = __ffs(hit) - 1;
            hit &= hit - 1;
            for (int u = lane; u < chain_len; u += 32) {
                const int k = 4 * (w + 32 * u) + c;
                if (k < K) {
                    const float pre = __fadd_rn(tmp, sdelta[k]);
                    const float est = __fadd_rn(pre, __ldg(round_base + tile_round_off(Kp, ncr, rr, k)));
                    if (ford(est) >= thr) {
                        const float x = exact_cand(pre, __ldg(LAd + (size_t)k * K + i));
                        if (x > -FLT_MAX) best_take(acc, x, k);
                    }
                }
            }
        }
    }
    return acc;
}

// Wider models (chains longer than 32): the chains inside the window straight from the chain-major double table
// LAcL[(i * 128 + q) * clp + u] — coalesced 256-byte requests, several in flight — instead of scan_slow()'s
// element-wise re-read of the tiled float table followed by a dependent load of the double.
__device__ __noinline__ Best scan_long(float cm0, float cm1, float cm2, float cm3, int thr, float tmp, const float *sdelta,
                                       const double *__restrict__ col, int clp, int K, int lane)
{
    Best acc{-FLT_MAX, 0x7fffffff};
    const float cm[4] = {cm0, cm1, cm2, cm3};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        unsigned hit = __ballot_sync(FULL_MASK, ford(cm[c]) >= thr);
        while (hit) {
            const int q = 4 * (__ffs(hit) - 1) + c;
            hit &= hit - 1;
            const double *ch = col + (size_t)q * clp;
#pragma unroll 4
            for (int u = lane; u < clp; u += 32) {
                const double la = __ldg(ch + u);
                const int k = q + 128 * u;
                if (k < K) {
                    const float pre = __fadd_rn(tmp, sdelta[k]);
                    if (ford(__fadd_rn(pre, __double2float_rn(la))) >= thr) {
                        const float x = exact_cand(pre, la);
                        if (x > -FLT_MAX) best_take(acc, x, k);
                    }
                }
            }
        }
    }
    return acc;
}

// The column's winner: exact candidates of everything inside the window, best of the warp.
__device__ __forceinline__ Best scan_settle(const Scan &sc, const float (&cm)[4], float tmp,
                                            const float *__restrict__ round_base, int ncr, int rr, const float *sdelta,
                                            const double *__restrict__ LAd, const double *__restrict__ LAcL, int clp, int K, int Kp,
                                            int i, int lane)
{
    Best acc{-FLT_MAX, 0x7fffffff};
    if (sc.live) {
        if (sc.overflow && LAcL != nullptr) {
            acc = scan_long(cm[0], cm[1], cm[2], cm[3], ford(sc.thr), tmp, sdelta, LAcL + (size_t)i * 128 * clp, clp, K, lane);
        } else if (sc.overflow) {
            acc = scan_slow(cm[0], cm[1], cm[2], cm[3], ford(sc.thr), tmp, round_base, ncr, rr, sdelta, LAd, K, Kp, i, lane);
        } else {
            const int k0 = sc.q0 + 128 * lane, k1 = sc.q1 + 128 * lane;  // element `lane` of either chain
            if (k0 < K) {
                const float pre = __fadd_rn(tmp, sdelta[k0]);
                if (__fadd_rn(pre, __double2float_rn(sc.la0)) >= sc.thr) {
                    const float x = exact_cand(pre, sc.la0);
                    if (x > -FLT_MAX) best_take(acc, x, k0);
                }
            }
            if (sc.q1 != NO_CHAIN && k1 < K) {
                const float pre = __fadd_rn(tmp, sdelta[k1]);
                if (__fadd_rn(pre, __double2float_rn(sc.la1)) >= sc.thr) {
                    const float x = exact_cand(pre, sc.la1);
                    if (x > -FLT_MAX) best_take(acc, x, k1);
                }
            }
        }
    }
    // best of the warp, "larger value, then smaller index", in two redux.sync instead of ten shuffles.  The value
    // comes back out of its integer image; that turns a -0 into +0, which cannot occur here (a sum is -0 only if
    // both terms are, and no logarithm is).
    const int ox = ford(acc.x);
    const int m = __reduce_max_sync(FULL_MASK, ox);
    Best b;
    b.k = __reduce_min_sync(FULL_MASK, ox == m ? acc.k : 0x7fffffff);
    b.x = __int_as_float(m >= 0 ? m : (int)((unsigned)(-m) | 0x80000000u));
    if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
    return b;
}

// TM: the first TM_RES_CHUNKS chunks of every owned column live in tensor memory (single-round CTAs only).
// PIN (needs TM): the rest of the CTA's slice fits shared memory and is loaded once (a.pinned).
// PEERS: state-sharded pass, results go to every GPU (a.npeer > 1).
// Compile-time switches rather than the fields of `a`: the step loop is latency-bound and branchy, and
// every variant it does not need is code its instruction fetch has to step over.
template <bool TM, bool PIN, bool PEERS>
__global__ void __launch_bounds__(NTHREADS, 1) k_flash_persist(const PersistArgs a)
{
    static_assert(TM || !PIN, "the pinned ring is a tensor-memory mode");
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
        if (PIN && lane == 0) {
            // the whole streamed part fits: one copy per chunk, packed back to back exactly as the table stores
            // them (the last chunk may be shorter), each with its own barrier; nothing is ever refilled
            const unsigned char *src = reinterpret_cast<const unsigned char *>(slab) + (size_t)nk_res * TILE_CH * ncols * sizeof(float);
            size_t off = 0;
            for (int u = nk_res; u < nk; ++u) {
                const uint32_t bytes = (uint32_t)(ncols * min(TILE_CH, a.Kp - u * TILE_CH)) * 4u;
                mbar_expect_tx(&full[u - nk_res], bytes);
                bulk_g2s(ring + off, src + off, bytes, &full[u - nk_res], 0, false);
                off += bytes;
            }
        }
        if (PIN) {
            // nothing to stream: the producer warp is done (prefetching this CTA's columns of the double table
            // into L2 here was tried — 125.8 MB against 126 MB of L2 — and changed nothing: 2.10 ms against 2.09)
        } else if (lane == 0) {
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
    // The emission terms of a step hang off a two-load chain (observation -> row of log B -> entry) that has
    // nothing to do with the delta hand-over, so it runs ahead of it: the observation one step early, the
    // entries of the first round before the poll.
    int ob_next = __ldg(a.ob + a.L + 1);
    const int pre_i0 = a.col_begin + c0 + min(warp * CPW, max(ncols - 1, 0));
    const int pre_i1 = a.col_begin + c0 + min(warp * CPW + 1, max(ncols - 1, 0));
    for (int s = 1; s <= a.nsteps; ++s) {
        unsigned long long *xout = a.xch + (size_t)(s & 1) * a.Kp;
        const bool last_step = s == a.nsteps;
        const int j = a.L + s;
        const float *tmp_row = a.LBf + (size_t)ob_next * a.Kp;  // F:167
        const float pre_tmp0 = __ldg(tmp_row + pre_i0), pre_tmp1 = __ldg(tmp_row + pre_i1);
        if (!last_step) ob_next = __ldg(a.ob + j + 1);
        const bool keep = j >= a.mid + 1;                        // F:242
        const bool tracing = a.trace != nullptr && s <= TRACE_STEPS && lane == 0 && (warp == 0 || warp == 1);  // one warp of either phase order, both with two columns
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
            // a missing column stands in as the CTA's last one in the preloaded pair, as its first one here: either way unused
            const float tmp0 = rho == 0 && have0 ? pre_tmp0 : __ldg(tmp_row + i0);
            const float tmp1 = rho == 0 && have1 ? pre_tmp1 : __ldg(tmp_row + i1);
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
#pragma unroll 1
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
                            tmem_ld32(tbase + 8u * (uint32_t)u, ha, hb);
#pragma unroll
                            for (int e = 0; e < 4; ++e) dd[e] = d4[(u + e) * 32];
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
                } else if (PIN) {
                    // pinned: the rest of the slice sits in shared memory, packed chunk after chunk.  Four 128-state
                    // iterations per trip, all twelve loads requested before the 96 FP instructions that use them —
                    // the same shape as the tensor-memory part above.  (One stage at a time, 6 loads then 48 FP
                    // instructions, left this phase at 2.3-2.9 us against 0.8-1.8 us for the tensor-memory part.)
                    if (s == 1)
                        for (int u = nk_res; u < nk; ++u) mbar_wait(&full[u - nk_res], 0);
                    if (have0) {
                        const float4 *base4 = reinterpret_cast<const float4 *>(ring) + lane;
                        const int n_it = (a.Kp - nk_res * TILE_CH) >> 7;  // 128-state iterations in shared memory
                        const int rs0 = have0 ? rr0 : 0, rs1 = have1 ? rr1 : rs0;
                        auto it_ptr = [&](int it, int row) -> const float4 * {
                            const int c = it >> 1;  // chunk (TILE_CH = 256 states = two iterations); the last one may be shorter
                            const int len4 = min(TILE_CH, a.Kp - (nk_res + c) * TILE_CH) >> 2;
                            return base4 + (size_t)c * ncr * (TILE_CH >> 2) + row * len4 + (it & 1) * 32;
                        };
                        int it = 0;
#pragma unroll 1
                        for (; it + 4 <= n_it; it += 4) {
                            float4 ha[4], hb[4], dd[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                ha[e] = *it_ptr(it + e, rs0);
                                hb[e] = *it_ptr(it + e, rs1);
                                dd[e] = dring[(it + e) * 32];
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) { FV_ACC2(dd[e], ha[e], hb[e]) }
                        }
                        for (; it < n_it; ++it) {
                            const float4 d = dring[it * 32];
                            const float4 h0 = *it_ptr(it, rs0);
                            const float4 h1 = *it_ptr(it, rs1);
                            FV_ACC2(d, h0, h1)
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
                if (tracing && ph == 0) tr[6] = clock64();  // between the two operand phases (tensor memory / shared memory, in the warp's order)
            }
            if (tracing) tr[2] = clock64();
            if (!have0) continue;  // warp-uniform
            const float *round_base = slab + (size_t)rho * TILE_RW * a.Kp;
            Scan s0, s1;
            scan_fetch(s0, cm0, a.LAc, i0, true, lane);
            scan_fetch(s1, cm1, a.LAc, i1, have1, lane);
            if (tracing) tr[3] = clock64();
            const Best r0 = scan_settle(s0, cm0, tmp0, round_base, ncr, rr0, sdelta, a.LAd, a.LAcL, a.clp, a.K, a.Kp, i0, lane);
            if (lane == 0) {
                if (PEERS) {
                    if (keep)
#pragma unroll 1
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
                const Best r1 = scan_settle(s1, cm1, tmp1, round_base, ncr, rr1, sdelta, a.LAd, a.LAcL, a.clp, a.K, a.Kp, i1, lane);
                if (lane == 0) {
                    if (PEERS) {
                        if (keep)
#pragma unroll 1
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
    if (PEERS) {
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


// =================================================================================================
// The same pass with a HALF-PRECISION filter (models up to 4096 states, unsharded).
//
// The float sweep above moves 4 bytes and issues 3 FP32 instructions per (source, destination) pair only to
// find out which few sources need the exact evaluation.  A much coarser estimate does that job as well, as
// long as its error is bounded rigorously and the window is widened to match:
//     a_k   = delta[k] - c            c = the largest delta of the step (every CTA computes it while staging)
//     est_k = fl16( fl16(max(a_k, -60000)) + fl16(log A[k][i]) )          one HADD2 per two pairs
// and the column keeps, per lane, the maxima of 8 chains (chain q = k & 255) with one HMNMX2 per two pairs:
// 1 instruction and 2 bytes per pair instead of 3 and 4, and no emission term at all (it is common to the
// column).  Error: with v_k = a_k + log A[k][i] (reals, both terms <= 0, no cancellation) the three half
// roundings give |est_k - v_k| <= 2.01 * 2^-11 |v_k| + 2^-22 while nothing is clamped, and the reference's
// candidate differs from tmp + c + v_k by at most 2 float spacings at that magnitude.  So with kt the argmax
// of est and k* the reference's winner (or any source tying with it):
//     est(k*) >= est(kt) - W,   W = 2.1 * 2^-10 |est(kt)| + 2^-20 + 2^-20 (|tmp| + |c| + |est(kt)|)
// (constants rounded up; tests/test_host_logic.py checks the bound on random and adversarial data).  Every
// chain whose maximum reaches est(kt) - W — 1.24 chains per column on the headline model, 8 or fewer in all
// but one column in a million — is then evaluated EXACTLY, all 16 elements, from the chain-major double
// table; the winner is the exact maximum with the lowest index, as the reference's strict '>' finds it.
// The clamp keeps est finite unless log A is -inf (so a column whose best estimate is -inf is dead for
// certain); it voids the bound only when est(kt) < -30000, and such a column (never seen) is evaluated in full.
// Stage delta_{s-1}: poll as delta_wait_load() does, keep the floats for the exact evaluation, and stage
// fl16(max(delta - c, H_CLAMP)) for the sweep.  c must be >= every entry and close to the largest one.  In the
// first step it IS the largest one (one redux.sync per warp, 14 words through shared memory, one more CTA
// barrier); afterwards the caller passes a bound it derived from the previous vector's maximum BEFORE the poll
// (c_bound), so no barrier separates the poll from the conversion, and this call only leaves the per-warp maxima
// of the vector it stages in wmax[] for the next step's bound.  Returns c; <= -FLT_MAX means no state is alive.
__device__ __forceinline__ float delta_stage16(const PersistArgs &a, int s, float *sdelta, __half *sdelta16, int *wmax, float c_bound,
                                               int ctid)
{
    constexpr int NB = 5;  // pairs per thread: 4096 / 2 / NCONS rounded up
    static_assert(NB * NCONS * 2 >= 4096, "one trip must cover the vector");
    const int Kp2 = a.Kp >> 1, Kh2 = a.Kp16 >> 1;
    float2 v[NB];
    if (s == 1) {
        const float2 *in2 = reinterpret_cast<const float2 *>(a.d_init);
#pragma unroll
        for (int e = 0; e < NB; ++e) {
            const int t = ctid + e * NCONS;
            v[e] = t < Kp2 ? __ldcg(in2 + t) : make_float2(0.f, 0.f);
        }
    } else {
        const unsigned long long *x = a.xch + (size_t)((s - 1) & 1) * a.Kp;
        const unsigned want = step_tag(a, s - 1);
        unsigned long long w0[NB], w1[NB];
        unsigned pending = 0;
#pragma unroll
        for (int e = 0; e < NB; ++e) {
            const int t = ctid + e * NCONS;
            w0[e] = w1[e] = 0;
            if (t < Kp2 && 2 * t < a.K) pending |= 1u << e;
        }
        unsigned long long tw = 0;
        for (uint32_t spins = 0; pending; ++spins) {
#pragma unroll
            for (int e = 0; e < NB; ++e)
                if (pending >> e & 1u) ld_volatile_2x64(x + 2 * (ctid + e * NCONS), w0[e], w1[e]);
#pragma unroll
            for (int e = 0; e < NB; ++e)
                if (pending >> e & 1u) {
                    const int k = 2 * (ctid + e * NCONS);
                    if ((unsigned)(w0[e] >> 32) == want && (k + 1 >= a.K || (unsigned)(w1[e] >> 32) == want)) pending &= ~(1u << e);
                }
            if (pending) watchdog(spins * 64u + 63u, tw, a.watchdog_ns);
        }
#pragma unroll
        for (int e = 0; e < NB; ++e) v[e] = make_float2(__uint_as_float((unsigned)w0[e]), __uint_as_float((unsigned)w1[e]));
    }
    float mx = -INFINITY;
    float2 *sd2 = reinterpret_cast<float2 *>(sdelta);
#pragma unroll
    for (int e = 0; e < NB; ++e) {
        const int t = ctid + e * NCONS, k = 2 * t;
        if (k >= a.K) v[e].x = 0.f;  // padding stays finite (the tables pad with -inf)
        else mx = fmaxf(mx, v[e].x);
        if (k + 1 >= a.K) v[e].y = 0.f;
        else mx = fmaxf(mx, v[e].y);
        if (t < Kp2) sd2[t] = v[e];
    }
    const int wm = __reduce_max_sync(FULL_MASK, ford(mx));
    if ((ctid & 31) == 0) wmax[ctid >> 5] = wm;
    float c = c_bound;
    if (s == 1 || a.exact_max) {
        named_bar_sync(1, NCONS);
        const int lane = ctid & 31;
        c = unford(__reduce_max_sync(FULL_MASK, lane < NCW ? wmax[lane] : ford(-INFINITY)));
    }
    if (c > -FLT_MAX) {
        __half2 *sh2 = reinterpret_cast<__half2 *>(sdelta16);
#pragma unroll
        for (int e = 0; e < NB; ++e) {
            const int t = ctid + e * NCONS, k = 2 * t;
            if (t < Kh2) {
                const float lo = k < a.K ? fmaxf(__fsub_rn(v[e].x, c), H_CLAMP) : 0.f;
                const float hi = k + 1 < a.K ? fmaxf(__fsub_rn(v[e].y, c), H_CLAMP) : 0.f;
                sh2[t] = __floats2half2_rn(lo, hi);
            }
        }
    }
    named_bar_sync(1, NCONS);
    return c;
}

__global__ void __launch_bounds__(NTHREADS, 1) k_flash_persist16(const PersistArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    int *wmax_base = reinterpret_cast<int *>(full + MAX_STAGES);  // [2][16] per-warp maxima of the staged vectors
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(full + 2 * MAX_STAGES);
    // (Two copies of the staged vector and no CTA barrier at the end of a step were tried: the warps that finish
    // early then spin in the next poll and take issue slots from the stragglers, whose scan -> publish chain IS the
    // critical path of the whole grid — 1.10 -> 1.53 ms.  Parked at a barrier they cost nothing.)
    const size_t stage_bytes = (size_t)a.Kp * 4 + (size_t)a.Kp16 * 2;
    unsigned char *stage_base = smem_raw + CTRL_BYTES;
    const uint4 *ring = reinterpret_cast<const uint4 *>(stage_base + stage_bytes);  // the table rows that are not in tensor memory

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int c0 = tile_c0(a.ncol, G, b), ncols = tile_c0(a.ncol, G, b + 1) - c0;  // at most TILE_RW
    const int n_it = a.Kp16 >> 8;               // sweep iterations of 256 source states
    const int it_tm = min(a.tm_iters, n_it) & ~3;  // those served from tensor memory, four per tcgen05.ld
    const __half *slab = a.hi16 + (size_t)c0 * a.Kp16;  // [iteration][column][256]

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *tmem_slot + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2);

    if (warp == NCW) {
        // the rows that stay in shared memory: loaded once, one bulk copy per iteration
        if (lane == 0 && it_tm < n_it && ncols > 0) {
            const uint32_t per_it = (uint32_t)ncols * 512u;
            mbar_expect_tx(&full[0], per_it * (uint32_t)(n_it - it_tm));
            for (int it = it_tm; it < n_it; ++it)
                bulk_g2s(reinterpret_cast<unsigned char *>(const_cast<uint4 *>(ring)) + (size_t)(it - it_tm) * per_it,
                         slab + (size_t)it * ncols * 256, per_it, &full[0], 0, false);
        }
        return;
    }

    const int rr0 = warp * CPW, rr1 = rr0 + 1;
    const bool have0 = rr0 < ncols, have1 = rr1 < ncols;
    const int rs0 = have0 ? rr0 : 0, rs1 = have1 ? rr1 : rs0;  // a duplicate row stands in for a missing one
    const int i0 = a.col_begin + c0 + rs0, i1 = a.col_begin + c0 + rs1;
    if (have0) tmem_fill16(tbase, slab, ncols, rs0, rs1, it_tm, lane);
    tmem_wait_st();
    tmem_fence_before();
    named_bar_sync(1, NCONS);
    tmem_fence_after();

    int ob_next = __ldg(a.ob + a.L + 1), ob_prev = 0;
    for (int s = 1; s <= a.nsteps; ++s) {
        float *sdelta = reinterpret_cast<float *>(stage_base);
        __half *sdelta16 = reinterpret_cast<__half *>(sdelta + a.Kp);
        int *wmax = wmax_base + 16 * (s & 1);
        unsigned long long *xout = a.xch + (size_t)(s & 1) * a.Kp;
        const bool last_step = s == a.nsteps;
        const int j = a.L + s;
        const float *tmp_row = a.LBf + (size_t)ob_next * a.Kp;  // F:167
        const float tmp0 = __ldg(tmp_row + i0), tmp1 = __ldg(tmp_row + i1);
        // A bound on the vector about to arrive, from the maximum of the one before it (left in wmax[] by the
        // previous step's staging): delta_{s-1}[i] = fl32(fl64(fl32(tmp_i + delta_{s-2}[k]) + log A[k][i]))
        //   <= fl32(fl64(fl32(tmpmax + max delta_{s-2}) + max log A))  — every rounding is monotone.
        float c_bound = 0.f;
        if (s > 1) {
            const float prev = unford(__reduce_max_sync(FULL_MASK, lane < NCW ? wmax_base[16 * ((s - 1) & 1) + lane] : ford(-INFINITY)));
            c_bound = prev > -FLT_MAX ? exact_cand(__fadd_rn(__ldg(a.LBmax + ob_prev), prev), a.lamax) : -FLT_MAX;
            if (!(c_bound > -FLT_MAX)) c_bound = -FLT_MAX;
        }
        ob_prev = ob_next;
        if (!last_step) ob_next = __ldg(a.ob + j + 1);
        const bool keep = j >= a.mid + 1;  // F:242
        const bool tracing = a.trace != nullptr && s <= TRACE_STEPS && lane == 0 && (warp == 0 || warp == 1);
        long long *tr = tracing ? a.trace + ((((size_t)(s - 1) * G + b) * 2 + warp) * TRACE_PTS) : nullptr;
        if (tracing) tr[0] = clock64();
        const float c = delta_stage16(a, s, sdelta, sdelta16, wmax, c_bound, tid);
        if (tracing) tr[1] = clock64();
        if (s == 1 && it_tm < n_it) mbar_wait(&full[0], 0);

        Best r0{-FLT_MAX, -1}, r1{-FLT_MAX, -1};
        if (have0 && c > -FLT_MAX) {  // else: no state alive, every column is dead
            const __half2 ninf = __float2half2_rn(-INFINITY);
            __half2 m0[4] = {ninf, ninf, ninf, ninf}, m1[4] = {ninf, ninf, ninf, ninf};
            const uint4 *d4 = reinterpret_cast<const uint4 *>(sdelta16) + lane;
#define FV_ACC16(D, H0, H1)                                      \
    m0[0] = __hmax2(m0[0], __hadd2(u2h((D).x), u2h((H0).x)));     \
    m0[1] = __hmax2(m0[1], __hadd2(u2h((D).y), u2h((H0).y)));     \
    m0[2] = __hmax2(m0[2], __hadd2(u2h((D).z), u2h((H0).z)));     \
    m0[3] = __hmax2(m0[3], __hadd2(u2h((D).w), u2h((H0).w)));     \
    m1[0] = __hmax2(m1[0], __hadd2(u2h((D).x), u2h((H1).x)));     \
    m1[1] = __hmax2(m1[1], __hadd2(u2h((D).y), u2h((H1).y)));     \
    m1[2] = __hmax2(m1[2], __hadd2(u2h((D).z), u2h((H1).z)));     \
    m1[3] = __hmax2(m1[3], __hadd2(u2h((D).w), u2h((H1).w)));
            // tensor memory and shared memory are separate read paths: odd warps take the shared-memory rows first
            const bool ring_first = warp & 1;
#pragma unroll 1
            for (int ph = 0; ph < 2; ++ph) {
                if ((ph == 0) != ring_first) {
#pragma unroll 1
                    for (int u = 0; u < it_tm; u += 4) {
                        uint32_t r[32];
                        uint4 dd[4];
                        tmem_ld32_raw(tbase + 8u * (uint32_t)u, r);
#pragma unroll
                        for (int e = 0; e < 4; ++e) dd[e] = d4[(u + e) * 32];
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const uint4 ha = make_uint4(r[8 * e], r[8 * e + 1], r[8 * e + 2], r[8 * e + 3]);
                            const uint4 hb = make_uint4(r[8 * e + 4], r[8 * e + 5], r[8 * e + 6], r[8 * e + 7]);
                            FV_ACC16(dd[e], ha, hb)
                        }
                    }
                } else {
                    const uint4 *p0 = ring + (size_t)rs0 * 32 + lane, *p1 = ring + (size_t)rs1 * 32 + lane;
                    const int step4 = ncols * 32;  // uint4 per iteration
                    int it = it_tm;
#pragma unroll 1
                    for (; it + 4 <= n_it; it += 4) {
                        uint4 ha[4], hb[4], dd[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            ha[e] = p0[(size_t)(it - it_tm + e) * step4];
                            hb[e] = p1[(size_t)(it - it_tm + e) * step4];
                            dd[e] = d4[(it + e) * 32];
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) { FV_ACC16(dd[e], ha[e], hb[e]) }
                    }
#pragma unroll 1
                    for (; it < n_it; ++it) {
                        const uint4 h0 = p0[(size_t)(it - it_tm) * step4], h1 = p1[(size_t)(it - it_tm) * step4], d = d4[it * 32];
                        FV_ACC16(d, h0, h1)
                    }
                }
                if (tracing && ph == 0) tr[6] = clock64();
            }
#undef FV_ACC16
            if (tracing) tr[2] = clock64();
            Scan16 s0, s1;
            scan16_fetch(s0, m0, tmp0, c, a.LAc16, i0, true, lane);
            scan16_fetch(s1, m1, tmp1, c, a.LAc16, i1, have1, lane);
            if (tracing) tr[3] = clock64();
            r0 = scan16_settle(s0, m0, tmp0, sdelta, a.LAc16, a.K, i0, lane);
            r1 = scan16_settle(s1, m1, tmp1, sdelta, a.LAc16, a.K, i1, lane);
        } else if (tracing) {
            tr[6] = tr[2] = tr[3] = clock64();
        }
        if (lane == 0 && have0) {
            publish_delta(xout, i0, r0.x, (int)step_tag(a, s));
            if (last_step) a.d_final[i0] = r0.x;
            if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i0, r0.k);
            if (have1) {
                publish_delta(xout, i1, r1.x, (int)step_tag(a, s));
                if (last_step) a.d_final[i1] = r1.x;
                if (keep) psi_store(a.psi, a.psi16, (size_t)(a.psi_row + (j - a.mid - 1)) * a.K + i1, r1.k);
            }
        }
        if (tracing) tr[4] = clock64();
        named_bar_sync(1, NCONS);  // the staged vector is overwritten by the next step's load; finished warps wait here, not in the poll
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
    // Pinned ring: with the first 2048 states of every column in tensor memory, what is left of a CTA's slice
    // (207 KB at K=3965 on 148 SMs) may fit shared memory whole.  Then it is loaded once: no refill traffic
    // (TMA writes cost the shared-memory pipe as much as the reads), no per-stage barrier hand-shakes, no L2 reads.
    a.pinned = 0;
    if (use_tmem && env_int("FLASHV_PIN", 1) != 0) {
        const int nk_res = std::min(TM_RES_CHUNKS, Kp / TILE_CH), nk = (Kp + TILE_CH - 1) / TILE_CH;
        const size_t rest = (size_t)(Kp - nk_res * TILE_CH) * cols_max * sizeof(float);
        if (nk - nk_res <= MAX_STAGES && fixed + rest <= (size_t)ctx->smem_optin) {
            a.pinned = 1, a.nstage = std::max(1, nk - nk_res);
            smem = fixed + std::max(rest, (size_t)16);
        }
    }
    const bool peers = a.npeer > 1;
    // Half-precision filter: whenever the model carries the tables (K <= 4096), the pass is not sharded and the
    // CTA's slice fits tensor memory + shared memory.  FLASHV_F16=0 keeps the float sweep.
    bool use_f16 = a.hi16 != nullptr && !peers && cols_max <= TILE_RW && env_int("FLASHV_F16", 1) != 0;
    if (use_f16) {
        const int n_it = a.Kp16 >> 8;
        a.exact_max = env_int("FLASHV_F16_EXACT_MAX", 0);
        a.tm_iters = env_int("FLASHV_F16_TM_ITERS", 16);  // all of them: measured 1.31 / 1.27 / 1.23 / 1.18 ms for 4 / 8 / 12 / 16 at K=3965
        const int it_tm = std::min(a.tm_iters, n_it) & ~3;
        const size_t need = CTRL_BYTES + (size_t)Kp * 4 + (size_t)a.Kp16 * 2 + (size_t)(n_it - it_tm) * cols_max * 512;
        if (it_tm > 16 || need > (size_t)ctx->smem_optin) use_f16 = false;
        else smem = std::max(need, (size_t)CTRL_BYTES + 64);
    }
    const void *fn = use_f16 ? (const void *)k_flash_persist16 : a.pinned  ? (peers ? (const void *)k_flash_persist<true, true, true> : (const void *)k_flash_persist<true, true, false>)
                     : use_tmem ? (peers ? (const void *)k_flash_persist<true, false, true> : (const void *)k_flash_persist<true, false, false>)
                                : (peers ? (const void *)k_flash_persist<false, false, true> : (const void *)k_flash_persist<false, false, false>);
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
    void *params[] = {(void *)&a};
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
    a.LAd = m->LAd, a.LAc = m->LAc, a.LAcL = m->LAcL, a.clp = m->clp, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
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
        a.hi16 = nullptr, a.LAc16 = nullptr, a.Kp16 = 0, a.LBmax = nullptr, a.lamax = 0.0;
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
        a.hi16 = m->hi16, a.LAc16 = m->LAc16, a.Kp16 = m->Kp16, a.LBmax = m->LBmax, a.lamax = m->lamax;
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
    a.hi16 = m->hi16, a.LAc16 = m->LAc16, a.Kp16 = m->Kp16, a.LBmax = m->LBmax, a.lamax = m->lamax;
    a.LAd = m->LAd, a.LAc = m->LAc, a.LAcL = m->LAcL, a.clp = m->clp, a.LBf = m->LBf, a.K = m->K, a.Kp = m->Kp;
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
