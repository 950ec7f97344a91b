// bs_kernels.cu — FLASH-BS passes on the device.
//
// One CTA owns one trellis vector (a sequence's first pass or one task) for all of its steps, so
// a pass needs no grid-wide synchronisation: per step the CTA's threads score every destination
// state against the B beam entries in heap-array order (S:437-446, exact double chain), then
// warp 0 rebuilds the beam by replaying the reference's min-heap insertions over the K scores
// (S:167-211) — the array layout of the heap is observable (next step scans slots in order with
// strict '>', the end scan of S:376-381 looks at slot 1 and slots B/2+2..B only), so the set of
// the top-B alone is not enough.
//
// The per-entry payload T3_State (S:55) is not carried in the heap: every step's predecessor
// state psi_j[i] goes to the backpointer store and the payload is recovered by walking it back
// from the end state, which is what the payload recursion of S:359-368 / S:448 computes forward.
//   S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
#include <stdio.h>
#include <stdlib.h>

#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace flashv {

// Floyd sift of S:96-123 for one node; hv/hs are 0-based images of heap slots 1..total.
__device__ __forceinline__ void heap_sift_held(float *hv, int *hs, int total, int parent, float v, int st)
{
    int child = 2 * parent;
    while (child <= total) {
        float cv = hv[child - 1];
        if (child + 1 <= total) {
            float rv = hv[child];
            if (cv > rv) ++child, cv = rv;
        }
        if (v <= cv) break;  // S:114 / S:152: ties stop the sift
        hv[parent - 1] = cv;
        hs[parent - 1] = hs[child - 1];
        parent = child;
        child *= 2;
    }
    hv[parent - 1] = v;
    hs[parent - 1] = st;
}

// Replay of generate_state_heap() (S:167-211) over score[0..K-1] by one warp.
//   i < B      : slot i+1 <- (score_i, i)                       S:172-179
//   i == B-1   : Floyd heapify                                   S:180-190
//   i >= B     : if score_i > H[1].Value replace root + sift     S:193-203
// Nodes of one depth have disjoint subtrees, so Floyd's node = total/2..1 order is reproduced by
// doing depths deepest-first with the nodes of a depth spread over lanes.  The streaming part
// ballots 32 scores at a time against the current minimum (which only grows), so lanes only
// serialise on entries that can still enter.
__device__ void heap_replay_warp(const float *score, int K, int B, float *hv, int *hs, int lane)
{
    for (int s = lane; s < B; s += 32) hv[s] = score[s], hs[s] = s;
    __syncwarp();
    const int last_parent = B / 2;
    if (last_parent >= 1) {
        int depth = 31 - __clz(last_parent);
        for (; depth >= 0; --depth) {
            const int lo = 1 << depth;
            const int hi = min((2 << depth) - 1, last_parent);
            for (int node = lo + lane; node <= hi; node += 32) heap_sift_held(hv, hs, B, node, hv[node - 1], hs[node - 1]);
            __syncwarp();
        }
    }
    float mn = hv[0];
    if (B >= 2 && B <= 128) {
        // Streaming part for beams of up to 128 entries.  Which child a sift follows depends only on
        // the heap, not on the new value, so every lane keeps, per depth d, a bit mask "the right
        // child of the j-th node of depth d is strictly smaller" (S:146).  The root-to-leaf min-child
        // path is then a handful of ALU operations that every lane computes redundantly, so no
        // shuffle is needed to agree on it: lane d fetches the entry at depth d, one ballot finds
        // where the new value stops (S:152: first depth with v <= entry), lanes shift the entries
        // above it up by one depth, and a second ballot refreshes the bits of the nodes whose
        // children changed.  Three shared-memory round trips per replacement, none per depth.
        constexpr int MAXD = 7;  // depth of node 128
        unsigned long long dm[MAXD];  // dm[d] bit j: node 2^d + j prefers its right child (depths 0..6)
#pragma unroll
        for (int d = 0; d < MAXD; ++d) {
            dm[d] = 0;
            const int first = 1 << d, count = 1 << d;
            for (int base = 0; base < count; base += 32) {
                const int n = first + base + lane;
                const bool r = base + lane < count && 2 * n + 1 <= B && hv[2 * n - 1] > hv[2 * n];
                dm[d] |= (unsigned long long)__ballot_sync(FULL_MASK, r) << base;
            }
        }
        for (int base = B; base < K; base += 32) {
            const int i = base + lane;
            const float s = i < K ? score[i] : -INFINITY;
            unsigned enter = __ballot_sync(FULL_MASK, s > mn);
            while (enter) {
                const int l0 = __ffs(enter) - 1;
                enter &= enter - 1;
                const float v = score[base + l0];
                if (!(v > mn)) continue;  // S:193, against the minimum as it is now
                // min-child path: jd[d] = index within depth d of the path node, depth = last depth
                int jd[MAXD + 1];
                int depth = 0;
                jd[0] = 0;
#pragma unroll
                for (int d = 0; d < MAXD; ++d) {
                    const int node = (1 << d) + jd[d];
                    const bool more = depth == d && 2 * node <= B;
                    jd[d + 1] = 2 * jd[d] + (int)((dm[d] >> jd[d]) & 1ull);
                    if (more) depth = d + 1;
                }
                int mine = 1, parent = 1;  // path node at this lane's depth and the one above it
#pragma unroll
                for (int d = 1; d <= MAXD; ++d)
                    if (lane == d) mine = (1 << d) + jd[d], parent = (1 << (d - 1)) + jd[d - 1];
                const bool on_path = lane >= 1 && lane <= depth;
                const float cv = on_path ? hv[mine - 1] : 0.f;
                const int cs = on_path ? hs[mine - 1] : 0;
                const unsigned stopm = __ballot_sync(FULL_MASK, on_path && v <= cv);
                const int stop = stopm ? __ffs(stopm) - 1 : depth + 1;  // v ends at depth stop-1
                if (on_path && lane < stop) hv[parent - 1] = cv, hs[parent - 1] = cs;
                if (lane == stop - 1) hv[mine - 1] = v, hs[mine - 1] = base + l0;  // lane 0 has mine == 1
                __syncwarp();
                // nodes at depths 0 .. stop-2 had a child replaced: refresh their bits
                const bool upd = lane + 1 < stop && 2 * mine + 1 <= B;
                const bool nb = upd && hv[2 * mine - 1] > hv[2 * mine];
                const unsigned updm = __ballot_sync(FULL_MASK, upd), setm = __ballot_sync(FULL_MASK, nb);
#pragma unroll
                for (int d = 0; d < MAXD; ++d)
                    if (updm >> d & 1u) dm[d] = (dm[d] & ~(1ull << jd[d])) | ((unsigned long long)(setm >> d & 1u) << jd[d]);
                // the new minimum: the value itself if it stayed at the root, else the old depth-1 entry
                mn = stop == 1 ? v : __shfl_sync(FULL_MASK, cv, 1);
            }
        }
        __syncwarp();
        return;
    }
    for (int base = B; base < K; base += 32) {
        const int i = base + lane;
        const float s = i < K ? score[i] : -INFINITY;
        unsigned enter = __ballot_sync(FULL_MASK, s > mn);
        while (enter) {
            const int l = __ffs(enter) - 1;
            enter &= enter - 1;
            const float sv = __shfl_sync(FULL_MASK, s, l);
            if (sv > mn) {  // S:193, against the minimum as it is now
                if (lane == 0) heap_sift_held(hv, hs, B, 1, sv, base + l);
                __syncwarp();
                mn = hv[0];
            }
        }
    }
    __syncwarp();
}

struct BsArgs {
    const double *LAd, *LBd, *LPi;
    const float *LBf;
    int K, Kp, B, T;
    const VecDesc *vecs;
    int nvec;
    const int32_t *ob;
    int32_t *ans;
    float *score;
    void *psi;
    int psi16;
    const uint8_t *ismid;
    long long *trace;  // optional: per CTA {score cycles, heap cycles} (FLASHV_BS_TRACE), else null
};

// dynamic shared memory: float sscore[Kp]; float hv[2][B]; int hs[2][B]
__global__ void __launch_bounds__(1024) k_bs_pass(const BsArgs a)
{
    extern __shared__ float smem_f[];
    const int K = a.K, B = a.B, T = a.T;
    float *sscore = smem_f;
    float *hv0 = smem_f + a.Kp;
    int *hs0 = reinterpret_cast<int *>(hv0 + 2 * B);
    const int v = blockIdx.x;
    if (v >= a.nvec) return;
    const VecDesc vd = a.vecs[v];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int32_t *ob = a.ob + (size_t)vd.seq * T;
    int32_t *ans = a.ans + (size_t)vd.seq * T;

    // start vector, S:411-426 (S:314-320 for the first pass).  prev < 0 is the reference's
    // vit->A[-1][i], which aliases Pi[i].
    {
        const int prev = vd.L == 0 ? -1 : ans[vd.L - 1];
        const int o = ob[vd.L];
        for (int i = tid; i < K; i += nthr) {
            const double head = prev < 0 ? a.LPi[i] : a.LAd[(size_t)prev * K + i];
            sscore[i] = __double2float_rn(__dadd_rn(head, a.LBd[(size_t)o * K + i]));
        }
    }
    __syncthreads();
    int cur = 0;
    if (warp == 0) heap_replay_warp(sscore, K, B, hv0, hs0, lane);
    __syncthreads();

    long long t_score = 0, t_heap = 0;
    for (int j = vd.L + 1; j <= vd.R; ++j) {
        const long long c0 = clock64();
        const float *hv = hv0 + cur * B;
        const int *hs = hs0 + cur * B;
        const int o = ob[j];
        const bool keep = j >= vd.mid + 1;  // S:448: payload latches at j == mid+1
        for (int i = tid; i < K; i += nthr) {
            const float tmp = __ldg(a.LBf + (size_t)o * a.Kp + i);  // S:439
            float best = -FLT_MAX;
            int arg = -1;
            // S:440-446, slots in array order, strict '>'.  The row reads are independent of the
            // running maximum: fetch a batch of them before the compare chain consumes any.
            constexpr int UB = 8;
            int c = 0;
            for (; c + UB <= B; c += UB) {
                double la[UB];
#pragma unroll
                for (int e = 0; e < UB; ++e) la[e] = __ldg(a.LAd + (size_t)hs[c + e] * K + i);
#pragma unroll
                for (int e = 0; e < UB; ++e) {
                    const float x = exact_cand(__fadd_rn(tmp, hv[c + e]), la[e]);
                    if (x > best) best = x, arg = c + e;
                }
            }
            for (; c < B; ++c) {
                const float x = exact_cand(__fadd_rn(tmp, hv[c]), __ldg(a.LAd + (size_t)hs[c] * K + i));
                if (x > best) best = x, arg = c;
            }
            sscore[i] = best;
            if (keep) psi_store(a.psi, a.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + i, arg < 0 ? -1 : hs[arg]);
        }
        __syncthreads();
        const long long c1 = clock64();
        if (warp == 0) heap_replay_warp(sscore, K, B, hv0 + (cur ^ 1) * B, hs0 + (cur ^ 1) * B, lane);
        __syncthreads();
        t_score += c1 - c0, t_heap += clock64() - c1;
        cur ^= 1;
    }
    if (a.trace && tid == 0) a.trace[2 * v] = t_score, a.trace[2 * v + 1] = t_heap;

    if (tid == 0) {
        const float *hv = hv0 + cur * B;
        const int *hs = hs0 + cur * B;
        int state;
        if (vd.flags & VEC_FULL_RANGE) {  // S:374-383 / S:454-463
            float sc = hv[0];
            int arg = 0;
            for (int c = B / 2 + 1; c < B; ++c)
                if (hv[c] > sc) arg = c, sc = hv[c];
            state = hs[arg];
            ans[vd.R] = state;
            a.score[vd.seq] = sc;
        } else {  // Find_T3_State, S:73-86: -1 when Ans[R] fell out of the beam
            const int want = ans[vd.R];
            state = -1;
            for (int c = 0; c < B; ++c)
                if (hs[c] == want) {
                    state = want;
                    break;
                }
        }
        for (int j = vd.R; j >= vd.mid + 1; --j) {
            if (state >= 0) state = psi_load(a.psi, a.psi16, (size_t)(vd.psi_row + (j - vd.mid - 1)) * K + state);
            if ((vd.flags & VEC_FIRST_PASS) && a.ismid[j - 1]) ans[j - 1] = state;
        }
        if (!(vd.flags & VEC_FIRST_PASS)) ans[vd.mid] = state;
    }
}

static size_t bs_smem_bytes(int Kp, int B) { return (size_t)Kp * 4 + (size_t)B * 16; }

int bs_run_pass(flashv_plan *p, const Pass &pass)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    BsArgs a;
    a.LAd = m->LAd, a.LBd = m->LBd, a.LPi = m->LPi, a.LBf = m->LBf;
    a.K = m->K, a.Kp = m->Kp, a.B = p->B, a.T = p->T;
    a.vecs = p->d_vecs + pass.vec_offset, a.nvec = pass.nvec;
    a.ob = p->d_ob, a.ans = p->d_ans, a.score = p->d_score;
    a.psi = p->d_psi, a.psi16 = p->psi16, a.ismid = p->d_ismid;
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
    k_bs_pass<<<pass.nvec, 1024, smem, ctx->stream>>>(a);
    FV_CUDA(cudaGetLastError());
    ++p->launches;
    if (a.trace) {
        long long h[2];
        FV_CUDA(cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[flashv bs trace] nvec=%d steps=%d vector 0: score %lld cycles, heap %lld cycles\n", pass.nvec,
                pass.max_steps, h[0], h[1]);
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
    extern __shared__ float smem_f[];
    float *hv = smem_f;
    int *hs = reinterpret_cast<int *>(smem_f + B);
    const long long c0 = clock64();
    heap_replay_warp(score, K, B, hv, hs, threadIdx.x);
    if (debug && threadIdx.x == 0) printf("[flashv] heap replay K=%d B=%d: %lld cycles (scores in global memory)\n", K, B, clock64() - c0);
    for (int s = threadIdx.x; s < B; s += 32) hv_out[s] = hv[s], hs_out[s] = hs[s];
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
    const size_t smem = (size_t)B * 8;
    FV_CUDA(cudaFuncSetAttribute(k_bs_replay_once, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bs_replay_once<<<1, 32, smem, ctx->stream>>>(score_dev, K, B, hv_dev, hs_dev, getenv("FLASHV_BS_TRACE") != nullptr);
    FV_CUDA(cudaGetLastError());
    return FLASHV_OK;
}

}  // namespace flashv
