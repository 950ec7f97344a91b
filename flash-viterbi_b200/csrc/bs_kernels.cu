// bs_kernels.cu — FLASH-BS passes on the device.
//
// One thread-block cluster owns one trellis vector (a sequence's first pass or one task) for all of
// its steps, so a pass needs no grid-wide synchronisation.  Per step
//   1. the cluster scores every destination state against the B beam entries with the exact double
//      chain (S:437-446) — over the out-edge lists of the beam states when the model has them
//      (entries with A[s][i] == 0 give -inf and never win), else over the K x B table reads — and
//      all-gathers the scores into every CTA's shared memory;
//   2. every CTA builds the next beam with a radix select; the reference's min-heap insertions
//      (S:167-211) are replayed only where the heap's array layout is observable: ties at the beam's
//      minimum, the end scan of S:376-381 (slot 1 and slots B/2+2..B), and — at backtrack time — a
//      visited backpointer whose maximum was attained by two beam states (the next step scans the
//      slots in order with a strict '>').  See build_beam() for why the set alone suffices otherwise.
//
// The per-entry payload T3_State (S:55) is not carried in the heap: every step's predecessor
// state psi_j[i] goes to the backpointer store and the payload is recovered by walking it back
// from the end state, which is what the payload recursion of S:359-368 / S:448 computes forward.
//   S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
#include <stdio.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include <type_traits>

#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace cg = cooperative_groups;

namespace flashv {

// A heap entry as the kernels keep it: slot n (1-based, like the reference's array) lives at
// node[n] = {Value bits, State}; the two children of n are the 16 bytes at node[2n], so a sift level
// is one 128-bit shared-memory load.  node[B+1] is a sentinel with Value = +inf (never the smaller
// child), node[0] is unused.
struct __align__(8) HeapNode {
    float v;
    int s;
};

// Sift of S:96-123 / S:141-163 from `parent` with the held entry (v, st): follow the smaller child
// (the right one only if strictly smaller, S:146), stop at the first child that is >= v (S:152: ties
// stop the sift).
__device__ __forceinline__ void heap_sift_held(HeapNode *node, int total, int parent, float v, int st)
{
    int child = 2 * parent;
    while (child <= total) {
        const uint4 pair = *reinterpret_cast<const uint4 *>(node + child);
        float cv = __uint_as_float(pair.x);
        int cs = (int)pair.y;
        const float rv = __uint_as_float(pair.z);
        if (cv > rv) ++child, cv = rv, cs = (int)pair.w;
        if (v <= cv) break;
        node[parent] = HeapNode{cv, cs};
        parent = child;
        child *= 2;
    }
    node[parent] = HeapNode{v, st};
}

// The same sift from the root for the streaming part, where one lane runs it hundreds of times per
// step and its latency is the step's critical path.  The four grandchildren of the current node
// (32 contiguous bytes at node[4n]) are fetched while the children are being compared, so a level
// costs a short ALU chain instead of a shared-memory round trip.  Reads run ahead of the heap's
// end (up to node[2*total+3]): the buffer is that long, and entries past total+1 are never used.
__device__ __forceinline__ void heap_replace_root(HeapNode *node, int total, float v, int st)
{
    int n = 1;
    if (total >= 2) {
        uint4 kids = *reinterpret_cast<const uint4 *>(node + 2);
        uint4 ga = *reinterpret_cast<const uint4 *>(node + 4), gb = *reinterpret_cast<const uint4 *>(node + 6);
        while (true) {
            const bool right = __uint_as_float(kids.x) > __uint_as_float(kids.z);
            const float cv = __uint_as_float(right ? kids.z : kids.x);
            const int cs = (int)(right ? kids.w : kids.y);
            if (v <= cv) break;
            node[n] = HeapNode{cv, cs};
            n = 2 * n + (right ? 1 : 0);
            if (2 * n > total) break;
            kids = right ? gb : ga;
            ga = *reinterpret_cast<const uint4 *>(node + 4 * n);
            gb = *reinterpret_cast<const uint4 *>(node + 4 * n + 2);
        }
    }
    node[n] = HeapNode{v, st};
}

// Replay of generate_state_heap() (S:167-211) over score[0..K-1] by one warp.
//   i < B      : slot i+1 <- (score_i, i)                       S:172-179
//   i == B-1   : Floyd heapify                                   S:180-190
//   i >= B     : if score_i > H[1].Value replace root + sift     S:193-203
// Nodes of one depth have disjoint subtrees, so Floyd's node = total/2..1 order is reproduced by
// doing depths deepest-first with the nodes of a depth spread over lanes.  The streaming part
// ballots 32 scores at a time against the current minimum (which only grows); lane 0 then sifts the
// survivors one after the other — a sift is a chain of dependent 16-byte shared-memory loads, about
// 50 cycles per level, and nothing in it can be shared between lanes.
// With WAIT the scores arrive while the replay runs: chunk c (states 32c..32c+31) is complete once
// tags[c] == tag (written by the scoring warps of the same CTA).
template <bool WAIT>
__device__ void heap_replay_warp(const float *score, int K, int B, HeapNode *node, int lane, const volatile int *tags,
                                 int tag)
{
    auto wait_upto = [&](int last_state) {  // scores [0, last_state] are final
        if (!WAIT) return;
        const int c1 = min(last_state, K - 1) >> 5;
        while (tags[c1] != tag) {
        }
        // chunks complete in any order: the earlier ones were awaited by earlier calls, except
        // for the first call (the initial B states)
        __threadfence_block();
    };
    if (WAIT)
        for (int c = 0; c <= ((B - 1) >> 5); ++c) wait_upto(32 * c);
    for (int s = lane; s < B; s += 32) node[s + 1] = HeapNode{score[s], s};
    if (lane == 0) node[B + 1] = HeapNode{INFINITY, -1};
    __syncwarp();
    const int last_parent = B / 2;
    if (last_parent >= 1) {
        int depth = 31 - __clz(last_parent);
        for (; depth >= 0; --depth) {
            const int lo = 1 << depth;
            const int hi = min((2 << depth) - 1, last_parent);
            for (int n = lo + lane; n <= hi; n += 32) {
                const HeapNode held = node[n];
                heap_sift_held(node, B, n, held.v, held.s);
            }
            __syncwarp();
        }
    }
    float mn = node[1].v;
    for (int base = B; base < K; base += 32) {
        wait_upto(base + 31);
        if (WAIT && ((base + 31) >> 5) != (base >> 5)) wait_upto(base);  // B not a multiple of 32: two chunks
        const int i = base + lane;
        const float s = i < K ? score[i] : -INFINITY;
        unsigned enter = __ballot_sync(FULL_MASK, s > mn);
        if (enter) {
            if (lane == 0) {
                while (enter) {
                    const int l = __ffs(enter) - 1;
                    enter &= enter - 1;
                    const float v = score[base + l];
                    if (v > mn) {  // S:193, against the minimum as it is now
                        heap_replace_root(node, B, v, base + l);
                        mn = node[1].v;
                    }
                }
            }
            mn = __shfl_sync(FULL_MASK, mn, 0);
        }
    }
    __syncwarp();
}

// 1024 threads x 8 row reads in flight each.  (512 threads x 32 reads — the same bytes in flight with
// fewer, fatter threads — measured 35 % slower in the scoring phase: two rounds of states per CTA and
// register spills.)
constexpr int BS_THREADS = 1024;

struct BsArgs {
    const double *LAd, *LBd, *LPi;
    const int *csr_cut;       // out-edge lists by source (flash_sparse.cu: sparse_build), or null: dense scoring
    const uint16_t *csr_i;
    const double *csr_la;
    const float *LBf;
    int K, Kp, B, T;
    const VecDesc *vecs;
    int nvec;
    const int32_t *ob;
    int32_t *ans;
    float *score;
    void *psi;
    int psi16;
    int flagbit;        // backpointer-entry bit "several beam states attain this maximum"
    float *rows;        // [psi rows][K] score vectors of the steps mid .. R-1 (row of step j-1 = backpointer row of step j)
    int always_replay;  // FLASHV_BS_REPLAY=1: rebuild the heap by replay at every step (the slow, literal path)
    int score_buf1_off;  // floats from the first score vector to the second one (end of the CTA's other buffers)
    const uint8_t *ismid;
    long long *trace;  // optional: per CTA {score cycles, beam cycles} (FLASHV_BS_TRACE), else null
};

// ---- the beam of one step ---------------------------------------------------------------------
// What the NEXT step needs from the reference's heap is, almost always, only the SET of its B
// entries: a destination's value max_c x_c does not depend on the slot order, and its argmax does
// only when two beam states attain the maximum exactly (S:440-446 then keeps the first slot).
// The set is "the B largest scores": an entry above the final minimum tau can never have been
// evicted or refused (the root only grows), so the heap holds every score > tau plus some == tau;
// if the scores == tau are exactly as many as the free slots, the set is decided without knowing
// the heap's layout.  So per step the CTA runs a radix select (4 passes of 8 bits over a monotone
// integer key) instead of replaying K sequential heap insertions, and falls back to the replay
//   * when several scores tie at tau (the heap's layout decides which of them stay),
//   * at the last step of a full-range pass (the end scan S:376-381 reads slot positions),
//   * at backtrack time for a visited backpointer whose maximum was attained twice: the score
//     vector of the step before is kept in HBM, its heap is replayed and the destination is
//     re-evaluated in true slot order (bs_fix_entry).
// Nothing is approximated: the fast path is taken only where it provably equals the replay.
struct BeamScratch {
    int hist[256];
    int wsum[8];
    unsigned prefix;
    int need, count, ambiguous;
};

__device__ __forceinline__ unsigned score_key(float x)  // monotone: larger score <=> larger key
{
    return (unsigned)ford(x) ^ 0x80000000u;
}

// Whole CTA.  On return beam[0..B-1] holds the heap's entries (any order, or slots 1..B after a
// replay) and the return value says whether the replay ran (node[1..B] is then the true heap).
__device__ bool build_beam(const float *sscore, int K, int B, HeapNode *beam, HeapNode *node, BeamScratch *bs, bool force_replay)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    bool replay = force_replay;
    if (!replay) {
        if (tid == 0) bs->prefix = 0, bs->need = B, bs->count = 0, bs->ambiguous = 0;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (int b = tid; b < 256; b += nthr) bs->hist[b] = 0;
            __syncthreads();
            const unsigned prefix = bs->prefix;
            const unsigned himask = pass == 0 ? 0u : 0xffffffffu << (shift + 8);
            for (int i = tid; i < K; i += nthr) {
                const unsigned u = score_key(sscore[i]);
                if ((u & himask) == prefix) atomicAdd(&bs->hist[(u >> shift) & 255], 1);
            }
            __syncthreads();
            // bins from the top: thread t looks at bin 255-t; an inclusive scan over the threads gives
            // every bin the number of keys in bins above it, and exactly one bin straddles `need`
            {
                const int need = bs->need;
                int c = 0, incl = 0;
                if (tid < 256) {
                    c = bs->hist[255 - tid];
                    incl = c;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int t = __shfl_up_sync(FULL_MASK, incl, off);
                        if ((tid & 31) >= off) incl += t;
                    }
                    if ((tid & 31) == 31) bs->wsum[tid >> 5] = incl;
                }
                __syncthreads();
                if (tid < 256) {
                    int above = 0;  // keys in the bins of earlier warps
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (w < (tid >> 5)) above += bs->wsum[w];
                    const int before = above + incl - c;  // keys in bins strictly above this one
                    if (before < need && before + c >= need) {
                        bs->prefix = prefix | (unsigned)(255 - tid) << shift;
                        bs->need = need - before;
                        if (pass == 3) bs->ambiguous = c != need - before;  // several scores == tau and not all fit
                    }
                }
            }
            __syncthreads();
        }
        replay = bs->ambiguous != 0;
        if (!replay) {
            const unsigned tau = bs->prefix;
            for (int i = tid; i < K; i += nthr) {
                const float x = sscore[i];
                if (score_key(x) >= tau) beam[atomicAdd(&bs->count, 1)] = HeapNode{x, i};
            }
        }
    }
    if (replay) {
        if (tid < 32) heap_replay_warp<false>(sscore, K, B, node, tid, nullptr, 0);
        __syncthreads();
        for (int c = tid; c < B; c += nthr) beam[c] = node[c + 1];
    }
    __syncthreads();
    return replay;
}

// Exact backpointer of destination state i at a step whose previous score vector is `row`
// (executed by warp 0): replay the heap, then S:440-446 in slot order.
__device__ int bs_fix_entry(const BsArgs &a, const float *row, int i, int o, HeapNode *node, int lane)
{
    heap_replay_warp<false>(row, a.K, a.B, node, lane, nullptr, 0);
    const float tmp = __ldg(a.LBf + (size_t)o * a.Kp + i);
    Best b{-FLT_MAX, 0x7fffffff};
    for (int c = 1 + lane; c <= a.B; c += 32) {
        const HeapNode h = node[c];
        const float x = exact_cand(__fadd_rn(tmp, h.v), __ldg(a.LAd + (size_t)h.s * a.K + i));
        if (x > -FLT_MAX) best_take(b, x, c);  // larger value, then the earlier slot
    }
    b = warp_best(b);
    return b.k == 0x7fffffff ? -1 : node[b.k].s;
}

// dynamic shared memory: float sscore[Kp]; HeapNode beam[B] ; HeapNode node[2B+4]; BeamScratch
//
// One thread-block CLUSTER owns one trellis vector for all of its steps.  Scoring a step is K x B
// dependent-latency double reads (4 MB at K=3965, B=128) — more than one SM can keep in flight — so
// the destination states are split over the cluster's CTAs, two threads per state (half the beam
// each).  Every CTA writes its scores into the score vector of ALL CTAs through distributed shared
// memory; after a cluster barrier each CTA builds the beam for itself (identical inputs, identical
// result), so nothing but the scores ever crosses SMs.  CTA 0 does the end state and the walk back.
__global__ void __launch_bounds__(BS_THREADS) k_bs_pass(const BsArgs a)
{
    extern __shared__ float4 smem_f4[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    float *smem_f = reinterpret_cast<float *>(smem_f4);
    const int K = a.K, B = a.B, T = a.T;
    // Two score vectors: step j writes vector j&1 of every CTA while the slower CTAs may still be building
    // their beam from vector (j-1)&1 — one cluster barrier per step instead of two ("all scores arrived"; the
    // "everybody is done reading" barrier is implied: vector j&1 was last read two steps ago, before the
    // reader arrived at the barrier of step j-1).
    float *sscore0 = smem_f, *sscore1 = smem_f + a.score_buf1_off;
    HeapNode *beam = reinterpret_cast<HeapNode *>(smem_f + a.Kp);
    HeapNode *node = beam + ((B + 1) & ~1);
    BeamScratch *bs = reinterpret_cast<BeamScratch *>(node + 2 * B + 4);
    __shared__ int s_state;
    const int v = blockIdx.x / CS;
    if (v >= a.nvec) return;  // whole clusters leave together
    const VecDesc vd = a.vecs[v];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    const int32_t *ob = a.ob + (size_t)vd.seq * T;
    int32_t *ans = a.ans + (size_t)vd.seq * T;
    const bool full = (vd.flags & VEC_FULL_RANGE) != 0;
    // states per CTA: 8/CS consecutive eighths of the state range (the eighths are where the out-edge lists are cut)
    const int per8 = (K + 7) / 8, q_lo = rank * (8 / CS), q_hi = q_lo + 8 / CS;
    const int s_lo = min(K, q_lo * per8), s_hi = min(K, q_hi * per8);
    float *peer_score0[8];  // every CTA's vector 0 (vector 1 sits score_buf1_off floats further in each of them)
#pragma unroll
    for (int r = 0; r < 8; ++r) peer_score0[r] = r < CS ? cluster.map_shared_rank(sscore0, r) : sscore0;
    auto keep_row = [&](int j, const float *sscore) {  // the scores of step j are the "previous scores" of backpointer row j+1
        if (j < vd.mid || j > vd.R - 1) return;
        float *dst = a.rows + (size_t)(vd.psi_row + (j - vd.mid)) * K;
        for (int i = s_lo + tid; i < s_hi; i += nthr) dst[i] = sscore[i];
    };

    // start vector, S:411-426 (S:314-320 for the first pass).  prev < 0 is the reference's
    // vit->A[-1][i], which aliases Pi[i].  Every CTA computes all of it.
    {
        const int prev = vd.L == 0 ? -1 : ans[vd.L - 1];
        const int o = ob[vd.L];
        float *sstart = (vd.L & 1) ? sscore1 : sscore0;
        for (int i = tid; i < K; i += nthr) {
            const double head = prev < 0 ? a.LPi[i] : a.LAd[(size_t)prev * K + i];
            sstart[i] = __double2float_rn(__dadd_rn(head, a.LBd[(size_t)o * K + i]));
        }
    }
    __syncthreads();
    keep_row(vd.L, (vd.L & 1) ? sscore1 : sscore0);
    bool have_heap = build_beam((vd.L & 1) ? sscore1 : sscore0, K, B, beam, node, bs, a.always_replay || (full && vd.R == vd.L));
    cluster.sync();  // every CTA of the cluster is running (its shared memory may be written) before anyone's step-1 scores arrive

    long long t_score = 0, t_beam = 0, t_beam_max = 0;
    int n_replay = 0;
    const int Bh = (B + 1) >> 1;
    for (int j = vd.L + 1; j <= vd.R; ++j) {
        const long long c0 = clock64();
        const int o = ob[j];
        const bool keep = j >= vd.mid + 1;  // S:448: payload latches at j == mid+1
        const int boff = (j & 1) ? a.score_buf1_off : 0;  // this step's score vector
        float *sscore = (j & 1) ? sscore1 : sscore0;
        if (a.csr_cut) {
            // ---- scoring over the out-edges of the beam states (S:437-446 restricted to the candidates that
            // can win: an entry with A[s][i] == 0 gives -inf and never passes the strict '>').  Warp per
            // beam entry, lanes over the part of its out-edge list that falls into this CTA's states.
            // Pass 1: maximum per destination (shared-memory atomicMax on a monotone key).  Pass 2: the
            // same edges again — the ones attaining the maximum leave the smallest slot index and their
            // number (two or more = the reference's choice depends on the heap's slot order).
            unsigned *skey = reinterpret_cast<unsigned *>(bs + 1);  // [per8 * 8 / CS] each
            int *sslot = reinterpret_cast<int *>(skey + (s_hi - s_lo));
            int *scnt = sslot + (s_hi - s_lo);
            const int nloc = s_hi - s_lo;
            const unsigned KEY_DEAD = score_key(-FLT_MAX);
            for (int t = tid; t < nloc; t += nthr) skey[t] = KEY_DEAD, sslot[t] = 0x7fffffff, scnt[t] = 0;
            __syncthreads();
            const float *tmp_row = a.LBf + (size_t)o * a.Kp;  // S:439
            const int warp = tid >> 5, nwarp = nthr >> 5;
            // A warp takes UE beam entries at a time so that their list bounds, then their edges, are
            // requested together: the walk is a chain of dependent L2 reads (bounds -> edge -> emission
            // term), paid once per UE entries instead of once per entry.  Pass 2 re-reads the same
            // ~70 KB per CTA from L1.
            auto walk = [&](auto ue) {
            constexpr int UE = decltype(ue)::value;
#pragma unroll 1
            for (int pass2 = 0; pass2 < 2; ++pass2) {
                for (int c0 = warp; c0 < B; c0 += nwarp * UE) {  // entries c0, c0 + nwarp, ...: small beams still use every warp
                    HeapNode h[UE];
                    int e0[UE], e1[UE], longest = 0;
#pragma unroll
                    for (int u = 0; u < UE; ++u) {
                        h[u] = beam[min(c0 + u * nwarp, B - 1)];
                        e0[u] = __ldg(a.csr_cut + (size_t)h[u].s * 9 + q_lo);
                        e1[u] = c0 + u * nwarp < B ? __ldg(a.csr_cut + (size_t)h[u].s * 9 + q_hi) : 0;  // a missing entry has no edges
                    }
#pragma unroll
                    for (int u = 0; u < UE; ++u) longest = max(longest, e1[u] - e0[u]);
                    for (int off = lane; off < longest; off += 32) {
                        int i[UE];
                        double la[UE];
                        float tmpv[UE];
                        bool on[UE];
#pragma unroll
                        for (int u = 0; u < UE; ++u) {
                            on[u] = e0[u] + off < e1[u];
                            i[u] = on[u] ? (int)__ldg(a.csr_i + e0[u] + off) : s_lo;
                            la[u] = on[u] ? __ldg(a.csr_la + e0[u] + off) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < UE; ++u) tmpv[u] = __ldg(tmp_row + i[u]);
#pragma unroll
                        for (int u = 0; u < UE; ++u) {
                            const float x = exact_cand(__fadd_rn(tmpv[u], h[u].v), la[u]);
                            if (!on[u] || !(x > -FLT_MAX)) continue;
                            const unsigned key = score_key(x);
                            if (pass2 == 0) {
                                atomicMax(&skey[i[u] - s_lo], key);
                            } else if (key == skey[i[u] - s_lo]) {
                                atomicMin(&sslot[i[u] - s_lo], c0 + u * nwarp);
                                atomicAdd(&scnt[i[u] - s_lo], 1);
                            }
                        }
                    }
                }
                __syncthreads();
            }
            };
            if (B > nwarp)
                walk(std::integral_constant<int, 4>());
            else
                walk(std::integral_constant<int, 1>());  // small beams: one entry per warp, nothing to batch
            for (int t = tid; t < nloc; t += nthr) {
                const int i = s_lo + t;
                float best = -FLT_MAX;
                int arg = -1;
                if (skey[t] != KEY_DEAD) {
                    const int ob32 = (int)(skey[t] ^ 0x80000000u);  // inverse of score_key: ordinal -> float bits
                    best = __int_as_float(ob32 >= 0 ? ob32 : (int)((unsigned)(-ob32) | 0x80000000u));
                    arg = beam[sslot[t]].s;
                }
                const bool tie = scnt[t] > 1;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < CS) peer_score0[r][boff + i] = best;
                // with the true heap in beam[] (slot order) the smallest slot IS the reference's choice
                if (keep) psi_store(a.psi, a.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + i,
                                    arg >= 0 && tie && !have_heap ? arg | a.flagbit : arg);
            }
        } else {
        // (the loop bound is rounded up to whole pairs: both threads of a pair reach the shuffles)
        for (int li = tid >> 1; li < ((s_hi - s_lo + 15) & ~15); li += nthr >> 1) {
            const int i = s_lo + li;
            const bool live = i < s_hi;
            const int half = tid & 1;
            float best = -FLT_MAX;
            int arg = -1;
            bool tie = false;
            float tmp = 0.f;
            if (live) {
                tmp = __ldg(a.LBf + (size_t)o * a.Kp + i);  // S:439
                // S:440-446 over this thread's half of the beam.  The row reads are independent of the
                // running maximum: fetch a batch of them before the compare chain consumes any.
                int e0 = half ? Bh : 0;
                const int e1 = half ? B : Bh;
                auto batch = [&](auto ub) {  // ub() row reads in flight, then the compare chain
                    constexpr int UB = decltype(ub)::value;
                    for (; e0 + UB <= e1; e0 += UB) {
                        double la[UB];
#pragma unroll
                        for (int e = 0; e < UB; ++e) la[e] = __ldg(a.LAd + (size_t)beam[e0 + e].s * K + i);
#pragma unroll
                        for (int e = 0; e < UB; ++e) {
                            const HeapNode h = beam[e0 + e];
                            const float x = exact_cand(__fadd_rn(tmp, h.v), la[e]);
                            tie = x > best ? false : (tie || (x == best && arg >= 0));
                            if (x > best) best = x, arg = h.s;
                        }
                    }
                };
                batch(std::integral_constant<int, 8>());
                batch(std::integral_constant<int, 1>());
            }
            // the second half's result joins the first's: slot order = first half, then second half
            const float obest = __shfl_xor_sync(FULL_MASK, best, 1);
            const int oarg = __shfl_xor_sync(FULL_MASK, arg, 1);
            const bool otie = __shfl_xor_sync(FULL_MASK, (int)tie, 1) != 0;
            if (live && half == 0) {
                if (obest > best)
                    best = obest, arg = oarg, tie = otie;
                else if (obest == best && oarg >= 0 && arg >= 0)
                    tie = true;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < CS) peer_score0[r][boff + i] = best;
                // with the true heap in beam[] (slot order) the first maximum IS the reference's choice
                if (keep) psi_store(a.psi, a.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + i,
                                    arg >= 0 && tie && !have_heap ? arg | a.flagbit : arg);
            }
        }
        }
        cluster.sync();  // all scores of step j are in every CTA's vector j&1
        const long long c1 = clock64();
        keep_row(j, sscore);
        have_heap = build_beam(sscore, K, B, beam, node, bs, a.always_replay || (full && j == vd.R));
        __syncthreads();  // the beam is complete before this CTA's warps score step j+1 from it
        {
            const long long tb = clock64() - c1;
            t_score += c1 - c0, t_beam += tb, t_beam_max = tb > t_beam_max ? tb : t_beam_max;
            n_replay += have_heap ? 1 : 0;
        }
    }
    if (rank != 0) return;
    if (a.trace && tid == 0 && v == 0) a.trace[0] = t_score, a.trace[1] = t_beam, a.trace[2] = t_beam_max, a.trace[3] = n_replay;

    // ---- end state: S:374-383 / S:454-463 for a full-range pass, Find_T3_State (S:73-86) otherwise ----
    if (tid == 0) {
        int state = -1;
        if (full) {  // have_heap: slot 1, then slots B/2+2 .. B of the true heap
            float sc = node[1].v;
            int arg = 1;
            for (int c = B / 2 + 2; c <= B; ++c)
                if (node[c].v > sc) arg = c, sc = node[c].v;
            state = node[arg].s;
            ans[vd.R] = state;
            a.score[vd.seq] = sc;
        } else {  // -1 when Ans[R] fell out of the beam
            const int want = ans[vd.R];
            for (int c = 0; c < B; ++c)
                if (beam[c].s == want) {
                    state = want;
                    break;
                }
        }
        s_state = state;
    }
    __syncthreads();
    // ---- walk back (warp 0, every lane holds the same state) ------------------------------------------
    if (tid < 32) {
        int state = s_state;
        for (int j = vd.R; j >= vd.mid + 1; --j) {
            if (state >= 0) {
                const size_t ridx = (size_t)(vd.psi_row + (j - vd.mid - 1));
                const int raw = psi_load(a.psi, a.psi16, ridx * K + state);
                if (raw >= 0 && (raw & a.flagbit))
                    state = bs_fix_entry(a, a.rows + ridx * K, state, ob[j], node, lane);
                else
                    state = raw;
            }
            if (lane == 0 && (vd.flags & VEC_FIRST_PASS) && a.ismid[j - 1]) ans[j - 1] = state;
        }
        if (lane == 0 && !(vd.flags & VEC_FIRST_PASS)) ans[vd.mid] = state;
    }
}

static size_t bs_smem_base(int Kp, int B)
{
    // + the sparse scoring's per-destination key / slot / count arrays (at most all K states in one CTA)
    size_t n = (size_t)Kp * 4 + (size_t)(((B + 1) & ~1) + 2 * B + 4) * sizeof(HeapNode) + sizeof(BeamScratch) + (size_t)3 * Kp * 4 + 16;
    return (n + 15) & ~(size_t)15;
}
static size_t bs_smem_bytes(int Kp, int B) { return bs_smem_base(Kp, B) + (size_t)Kp * 4; }  // + the second score vector

int bs_run_pass(flashv_plan *p, const Pass &pass)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    BsArgs a;
    a.LAd = m->LAd, a.LBd = m->LBd, a.LPi = m->LPi, a.LBf = m->LBf;
    const bool dense_scoring = getenv("FLASHV_BS_DENSE") && atoi(getenv("FLASHV_BS_DENSE")) != 0;  // tests: force the K x B table reads
    a.csr_cut = dense_scoring ? nullptr : m->csr_cut, a.csr_i = m->csr_i, a.csr_la = m->csr_la;
    a.K = m->K, a.Kp = m->Kp, a.B = p->B, a.T = p->T;
    a.vecs = p->d_vecs + pass.vec_offset, a.nvec = pass.nvec;
    a.ob = p->d_ob, a.ans = p->d_ans, a.score = p->d_score;
    a.psi = p->d_psi, a.psi16 = p->psi16, a.ismid = p->d_ismid;
    a.flagbit = p->psi16 ? 0x8000 : 0x40000000, a.rows = p->d_bs_score;
    a.always_replay = getenv("FLASHV_BS_REPLAY") ? atoi(getenv("FLASHV_BS_REPLAY")) : 0;  // read per call: tests toggle it
    a.score_buf1_off = (int)(bs_smem_base(m->Kp, p->B) / sizeof(float));
    a.trace = nullptr;
    static long long *d_trace = nullptr;
    const bool tracing = getenv("FLASHV_BS_TRACE") != nullptr;
    if (tracing) {  // developer aid: cycles spent scoring vs rebuilding the heap, per vector
        if (!d_trace) FV_CUDA(cudaMalloc(&d_trace, sizeof(long long) * 2 * 65536));
        a.trace = pass.nvec <= 65536 ? d_trace : nullptr;
    }
    const size_t smem = bs_smem_bytes(m->Kp, p->B);
    if (smem > (size_t)ctx->smem_optin) {
        set_error("FLASH-BS: K=%d, B=%d needs %zu bytes of shared memory per CTA (limit %d)", m->K, p->B, smem,
                  ctx->smem_optin);
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaFuncSetAttribute(k_bs_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // cluster size: enough CTAs that a step's K x B row reads are spread over several SMs, but no more
    // clusters x CTAs than the GPU holds at once
    int cs = 8;
    while (cs > 1 && ((size_t)m->K * p->B < (size_t)cs * 32768 || pass.nvec * cs > 2 * ctx->sm_count)) cs >>= 1;
    if (const char *force = getenv("FLASHV_BS_CLUSTER")) {  // tests: 1, 2, 4 or 8 regardless of the sizes
        const int f = atoi(force);
        if (f == 1 || f == 2 || f == 4 || f == 8) cs = f;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)pass.nvec * cs), cfg.blockDim = dim3(BS_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    FV_CUDA(cudaLaunchKernelEx(&cfg, k_bs_pass, a));
    FV_CUDA(cudaGetLastError());
    ++p->launches;
    if (a.trace) {
        long long h[4];
        FV_CUDA(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[flashv bs trace] nvec=%d steps=%d vector 0: score %lld cycles, beam %lld cycles (slowest step %lld), %lld steps replayed the heap\n",
                pass.nvec, pass.max_steps, h[0], h[1], h[2], h[3]);
    }
    return FLASHV_OK;
}

// ---- single-step hooks for the parity tests ---------------------------------------------------
__global__ void __launch_bounds__(256) k_bs_score_once(const double *__restrict__ LAd, const float *__restrict__ LBf,
                                                       int K, int Kp, const float *__restrict__ hv,
                                                       const int32_t *__restrict__ hs, int B, int o,
                                                       float *__restrict__ score, int32_t *__restrict__ arg_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float tmp = LBf[(size_t)o * Kp + i];
    float best = -FLT_MAX;
    int arg = -1;
    for (int c = 0; c < B; ++c) {
        const float pre = __fadd_rn(tmp, hv[c]);
        const float x = exact_cand(pre, LAd[(size_t)hs[c] * K + i]);
        if (x > best) best = x, arg = c;
    }
    score[i] = best;
    arg_out[i] = arg;
}

__global__ void k_bs_replay_once(const float *__restrict__ score, int K, int B, float *hv_out, int32_t *hs_out, int debug)
{
    extern __shared__ float4 smem_f4[];
    HeapNode *node = reinterpret_cast<HeapNode *>(smem_f4);
    const long long c0 = clock64();
    heap_replay_warp<false>(score, K, B, node, threadIdx.x, nullptr, 0);
    if (debug && threadIdx.x == 0) printf("[flashv] heap replay K=%d B=%d: %lld cycles (scores in global memory)\n", K, B, clock64() - c0);
    for (int s = threadIdx.x; s < B; s += 32) hv_out[s] = node[s + 1].v, hs_out[s] = node[s + 1].s;
}

int bs_single_score(flashv_model *m, const float *hv_dev, const int32_t *hs_dev, int B, int o, float *score_dev,
                    int32_t *arg_dev)
{
    k_bs_score_once<<<(m->K + 255) / 256, 256, 0, m->ctx->stream>>>(m->LAd, m->LBf, m->K, m->Kp, hv_dev, hs_dev, B, o,
                                                                  score_dev, arg_dev);
    FV_CUDA(cudaGetLastError());
    return FLASHV_OK;
}

int bs_single_replay(flashv_ctx *ctx, const float *score_dev, int K, int B, float *hv_dev, int32_t *hs_dev)
{
    const size_t smem = (size_t)(2 * B + 4) * sizeof(HeapNode);
    FV_CUDA(cudaFuncSetAttribute(k_bs_replay_once, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bs_replay_once<<<1, 32, smem, ctx->stream>>>(score_dev, K, B, hv_dev, hs_dev, getenv("FLASHV_BS_TRACE") != nullptr);
    FV_CUDA(cudaGetLastError());
    return FLASHV_OK;
}

}  // namespace flashv
