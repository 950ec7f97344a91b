// flash_group.cu — batches of sequences over a small table (BASELINE config 4: 8192 sequences,
// K=512, T=1024): one CTA walks a group of 8 trellis vectors through ALL their steps.
//
// The batched path is bound by instruction issue, not by memory (the table is L2-resident and
// every value read meets 8 vectors), so the kernel is organised around instructions per update:
//
//   * one THREAD owns two destination columns i and all 8 vectors of the group: a 2x8 register
//     tile of running maxima, no cross-lane reduction at all.  Per source state k a thread issues
//     one 8-byte load of hiS[k][i..i+1] (source-major float table, 256 contiguous bytes per warp),
//     two broadcast 16-byte shared-memory loads of delta[k][0..7], and 16 x (FADD, FMNMX).
//   * 2 instructions per update instead of 3: the per-step emission term tmp = log B[i][o]
//     (F:167) does not depend on k, so the running maximum is kept of  m2(k) = delta[k] (+) hi[k][i]
//     and tmp is only applied to the winner.  m2(k) + tmp is within 2.5 float steps of the reference's
//     candidate value (derivation at filter_threshold()), so the exact first-argmax is among
//     the k with m2(k) >= max m2 - 12 G, G = the float spacing at |max m2 + tmp|.  Those few (almost
//     always one) are re-evaluated with the reference's exact rounding chain (F:170) from the
//     double table, lowest k first — the result is bit-identical to the reference's.
//   * which k?  The k range is swept in blocks of 32; at each block end a thread folds the block
//     maxima into (top, runner-up, block of top) — 5 instructions per pair per block.  If the
//     runner-up block is outside the window (97 % of the pairs at |delta| ~ 10^3) only the top
//     block has to be searched: the warp does that together, 32 lanes over the 32 states of a
//     lane's block (one coalesced 128-byte read of hiT[i][kb..kb+31]), one lane after the other.
//     Pairs with two blocks inside the window go to a list and a warp scans their whole column.
//   * delta ping-pongs in shared memory for the whole pass in two layouts ([k][vector] for the
//     sweep, [vector][k] for the searches); only backpointer rows and the final delta of
//     full-range vectors go to HBM.  Groups are independent: no inter-CTA dependency.
//   * a task's last step needs one column only (Ans[mid] = T2[cur][Ans[R]], F:248/F:261).
//
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <stdlib.h>

#include <algorithm>

#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace flashv {

constexpr int GQ = 8;            // vectors per group
constexpr int GNT = 256;         // threads per CTA: one column pair each (K = 512), two CTAs per SM
constexpr int GBLK = 32;         // source states per search block
constexpr int GLIST_CAP = 1024;  // two-block pairs per step handled by the list phase (more: resolved in place)

struct GroupArgs {
    const float *hiS;  // [Kp][Kp] (float)log A, source-major: hiS[k][i]; -inf padding
    const float *hiT;  // [K][Kp]  the same, destination-major
    const double *LAd, *LBd, *LPi;
    const float *LBf;
    int K, Kp;
    const VecDesc *vecs;
    int nvec;
    const int32_t *ob;
    const int32_t *ans;
    int T;
    float *dfinal;  // [nvec][Kp] delta after each vector's last step (full-range vectors read it)
    void *psi;
    int psi16;
};

// What a thread leaves in shared memory for the search of one of its 16 pairs.
struct PairRec {
    float thr;  // window threshold; +inf when the pair is not searched (dead, or listed)
    int kb;     // first source state of the block holding the largest estimate
};

size_t group_smem_bytes(int Kp)
{
    // delta [2][Kp][GQ] + delta [2][GQ][Kp] + PairRec [16][GNT] + VecDesc[GQ] + per-step vector info [6][GQ] + list
    return (size_t)4 * GQ * Kp * sizeof(float) + (size_t)2 * GQ * GNT * sizeof(PairRec) + GQ * sizeof(VecDesc) +
           (6 * GQ + GLIST_CAP + 4) * sizeof(int);
}

// 8 source states of the sweep: bm[r][q] = max(bm[r][q], delta[k][q] + h[k][r]).  Two states share
// one three-input maximum (FMNMX3: half rate on its pipe but one issue slot instead of two — the
// sweep is bound by issue slots).
__device__ __forceinline__ void sweep8(float (&bm)[2][GQ], const float2 (&h)[8], const float *dk)
{
#pragma unroll
    for (int u = 0; u < 8; u += 2) {
        const float4 da0 = reinterpret_cast<const float4 *>(dk + u * GQ)[0];
        const float4 db0 = reinterpret_cast<const float4 *>(dk + u * GQ)[1];
        const float4 da1 = reinterpret_cast<const float4 *>(dk + u * GQ)[2];
        const float4 db1 = reinterpret_cast<const float4 *>(dk + u * GQ)[3];
        const float d0[GQ] = {da0.x, da0.y, da0.z, da0.w, db0.x, db0.y, db0.z, db0.w};
        const float d1[GQ] = {da1.x, da1.y, da1.z, da1.w, db1.x, db1.y, db1.z, db1.w};
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
            bm[0][q] = fmaxf(fmaxf(bm[0][q], __fadd_rn(d0[q], h[u].x)), __fadd_rn(d1[q], h[u + 1].x));
            bm[1][q] = fmaxf(fmaxf(bm[1][q], __fadd_rn(d0[q], h[u].y)), __fadd_rn(d1[q], h[u + 1].y));
        }
    }
}

__global__ void __launch_bounds__(GNT, 2) k_flash_group_cols(const GroupArgs a)
{
    extern __shared__ float4 sgroup4[];
    const int K = a.K, Kp = a.Kp;
    float *skq = reinterpret_cast<float *>(sgroup4);      // [2][Kp][GQ]
    float *sqk = skq + (size_t)2 * Kp * GQ;               // [2][GQ][Kp]
    PairRec *srec = reinterpret_cast<PairRec *>(sqk + (size_t)2 * GQ * Kp);  // [2*GQ][GNT]
    VecDesc *svd = reinterpret_cast<VecDesc *>(srec + 2 * GQ * GNT);
    int *sobs = reinterpret_cast<int *>(svd + GQ);        // [2][GQ] observation of each vector at this / the next step
    int *spsi = sobs + 2 * GQ;                            // [2][GQ] backpointer row this step writes (-1: none), same parity scheme
    int *sfin = spsi + 2 * GQ;                            // [2][GQ] 1: this step is the vector's last (final delta goes out)
    int *scount = sfin + 2 * GQ;
    int *slist = scount + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ngroups = (a.nvec + GQ - 1) / GQ;
    const int ncp = Kp >> 1;  // column pairs (Kp is a multiple of 128: whole warps are in or out of a round)

    // what step s of vector q needs besides delta: its observation (F:167), the backpointer row it
    // appends (F:242: only steps j >= mid+1 are ever read back) and whether it is the last step
    auto step_info = [&](const VecDesc &d, int s, int q) {
        const int par = (s & 1) * GQ + q, j = d.L + s;
        sobs[par] = a.ob[(size_t)d.seq * a.T + min(j, d.R)];
        spsi[par] = (j <= d.R && j >= d.mid + 1) ? d.psi_row + (j - d.mid - 1) : -1;
        sfin[par] = j == d.R;
    };
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int v0 = g * GQ;
        __syncthreads();  // previous group done with the buffers
        if (tid < GQ) {
            VecDesc d = a.vecs[min(v0 + tid, a.nvec - 1)];
            if (v0 + tid >= a.nvec) d.R = d.L;  // padding vector: no steps
            svd[tid] = d;
            step_info(d, 1, tid);
        }
        __syncthreads();
        int gsteps = 0;
#pragma unroll
        for (int q = 0; q < GQ; ++q) gsteps = max(gsteps, svd[q].R - svd[q].L);
        const bool one_column_end = !(svd[0].flags & VEC_FULL_RANGE);  // flags are per pass
        // start vectors, F:142 / F:220
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
            const VecDesc d = svd[q];
            const int prev = d.L == 0 ? -1 : a.ans[(size_t)d.seq * a.T + d.L - 1];
            const int o = a.ob[(size_t)d.seq * a.T + d.L];
            for (int i = tid; i < Kp; i += GNT) {
                float v = 0.f;  // padding states stay finite (the tables pad with -inf)
                if (i < K) {
                    const double head = prev < 0 ? a.LPi[i] : a.LAd[(size_t)prev * K + i];
                    v = __double2float_rn(__dadd_rn(head, a.LBd[(size_t)o * K + i]));
                }
                skq[(size_t)i * GQ + q] = v;
                sqk[(size_t)q * Kp + i] = v;
            }
        }
        __syncthreads();
        int cur = 0;
        for (int s = 1; s <= gsteps; ++s) {
            const float *in_kq = skq + (size_t)cur * Kp * GQ;
            const float *in_qk = sqk + (size_t)cur * GQ * Kp;
            float *out_kq = skq + (size_t)(cur ^ 1) * Kp * GQ;
            float *out_qk = sqk + (size_t)(cur ^ 1) * GQ * Kp;
            const int *obs = sobs + (s & 1) * GQ;  // F:167: tmp = LBf[obs[q]][i]
            const int *psirow = spsi + (s & 1) * GQ, *fin = sfin + (s & 1) * GQ;
            if (s == gsteps && one_column_end) {
                // ---- the group's last step: one column per vector, three-operation estimate -------
                for (int q = warp; q < GQ; q += GNT / 32) {
                    const VecDesc d = svd[q];
                    if (d.R - d.L != s) continue;  // finished earlier (or padding)
                    const int e = a.ans[(size_t)d.seq * a.T + d.R];
                    if (e < 0 || e >= K) continue;
                    const float tmp = __ldg(a.LBf + (size_t)obs[q] * Kp + e);
                    const float *col = a.hiT + (size_t)e * Kp;
                    const float *delta = in_qk + (size_t)q * Kp;
                    const float4 *col4 = reinterpret_cast<const float4 *>(col);
                    const float4 *d4 = reinterpret_cast<const float4 *>(delta);
                    float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 4
                    for (int t = lane; t < (Kp >> 2); t += 32) {
                        const float4 h = __ldg(col4 + t);
                        const float4 dv = d4[t];
                        cm[0] = fmaxf(cm[0], __fadd_rn(__fadd_rn(tmp, dv.x), h.x));
                        cm[1] = fmaxf(cm[1], __fadd_rn(__fadd_rn(tmp, dv.y), h.y));
                        cm[2] = fmaxf(cm[2], __fadd_rn(__fadd_rn(tmp, dv.z), h.z));
                        cm[3] = fmaxf(cm[3], __fadd_rn(__fadd_rn(tmp, dv.w), h.w));
                    }
                    const Best b = resolve_column(cm, tmp, col, delta, a.LAd, K, Kp, e, lane);
                    if (lane == 0) psi_store(a.psi, a.psi16, (size_t)(d.psi_row + (d.R - d.mid - 1)) * K + e, b.k);
                }
                break;  // the loop over groups starts with a barrier
            }
            if (tid == 0) *scount = 0;  // readers of the previous step are behind that step's last barrier
            if (tid < GQ) step_info(svd[tid], s + 1, tid);  // the other parity: this step's entries are still read

            // the result of one (column, vector) pair: delta' in both layouts, backpointer, final delta
            auto emit = [&](int q, int i, Best b) {
                if (!(b.x > -FLT_MAX)) b.x = -FLT_MAX, b.k = -1;
                out_kq[i * GQ + q] = b.x;
                out_qk[q * Kp + i] = b.x;
                if (i >= K) return;  // padding column
                const int row = psirow[q];
                if (row >= 0) psi_store(a.psi, a.psi16, (size_t)row * K + i, b.k);
                if (fin[q]) a.dfinal[(size_t)(v0 + q) * Kp + i] = b.x;
            };

            for (int cp = tid; cp < ncp; cp += GNT) {
                const int i0 = 2 * cp;
                {
                    // ---- sweep: running maxima of m2 = delta[k] + hi[k][i], folded per block of 32 --
                    float top[2][GQ], second[2][GQ];
                    int topblk[2][GQ];
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int q = 0; q < GQ; ++q) top[r][q] = -INFINITY, second[r][q] = -INFINITY, topblk[r][q] = 0;
                    const float2 *hrow = reinterpret_cast<const float2 *>(a.hiS + i0);  // row k: hrow + k*hstride
                    const size_t hstride = (size_t)Kp >> 1;
                    const float2 *hend = hrow + (size_t)Kp * hstride;
                    float2 ha[8], hb[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) ha[u] = __ldg(hrow + u * hstride);
                    hrow += 8 * hstride;
                    const float *dk = in_kq;
                    for (int kb = 0; kb < Kp; kb += GBLK) {
                        float bm[2][GQ];
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int q = 0; q < GQ; ++q) bm[r][q] = -INFINITY;
#pragma unroll
                        for (int half = 0; half < GBLK / 16; ++half) {
#pragma unroll
                            for (int u = 0; u < 8; ++u) hb[u] = __ldg(hrow + u * hstride);
                            hrow += 8 * hstride;
                            sweep8(bm, ha, dk);
                            dk += 8 * GQ;
                            if (hrow < hend) {
#pragma unroll
                                for (int u = 0; u < 8; ++u) ha[u] = __ldg(hrow + u * hstride);
                            }
                            hrow += 8 * hstride;
                            sweep8(bm, hb, dk);
                            dk += 8 * GQ;
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r)
#pragma unroll
                            for (int q = 0; q < GQ; ++q) {
                                second[r][q] = fmaxf(second[r][q], fminf(top[r][q], bm[r][q]));
                                topblk[r][q] = bm[r][q] > top[r][q] ? kb : topblk[r][q];
                                top[r][q] = fmaxf(top[r][q], bm[r][q]);
                            }
                    }
                    // ---- thresholds; pairs with two blocks inside the window go to the list -----------
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int q = 0; q < GQ; ++q) {
                            const float tmp = __ldg(a.LBf + (unsigned)(obs[q] * Kp + i0 + r));  // rows are padded to Kp
                            float thr = filter_threshold(top[r][q], tmp);
                            if (!(top[r][q] > -FLT_MAX)) {
                                emit(q, i0 + r, Best{-FLT_MAX, -1});
                                thr = INFINITY;
                            } else if (second[r][q] >= thr) {
                                const int slot = atomicAdd(scount, 1);
                                if (slot < GLIST_CAP) {
                                    slist[slot] = q << 24 | (i0 + r);
                                } else {  // list full: this thread scans the column itself
                                    Best b{-FLT_MAX, 0x7fffffff};
                                    for (int k = 0; k < K; ++k) {
                                        const float dv = in_qk[(size_t)q * Kp + k];
                                        if (__fadd_rn(dv, __ldg(a.hiS + (size_t)k * Kp + i0 + r)) >= thr) {
                                            const float x = exact_cand(__fadd_rn(tmp, dv), __ldg(a.LAd + (size_t)k * K + min(i0 + r, K - 1)));
                                            if (x > -FLT_MAX) best_take(b, x, k);
                                        }
                                    }
                                    emit(q, i0 + r, b);
                                }
                                thr = INFINITY;
                            }
                            PairRec rec;
                            rec.thr = thr, rec.kb = topblk[r][q];
                            srec[(r * GQ + q) * GNT + tid] = rec;
                        }
                }
                __syncwarp();  // a warp only reads the records of its own lanes
                // ---- search of the top blocks, the warp together: eight lanes (4 states each, one 128-bit
                // load) cover the 32 states of one owner lane's block, so one instruction serves four
                // owners.  A unit = one pair set (r,q) x 16 owners = 4 loads per lane; three units are in
                // flight.  Owners get their hits as 4 bytes (byte c, bit t <-> state kb + 4t + c) and leave
                // them in their record's threshold slot.
                const int cp_warp = cp - lane;  // the warp's first column pair
                const int sub = lane & 7, grp = lane >> 3;
                const unsigned sel_lo = (unsigned)grp | (unsigned)(4 + grp) << 4;  // bytes {x.grp, y.grp}
                auto issue = [&](int u, float4 (&hbuf)[4]) {
                    const int p = u >> 1, r = p / GQ, l0 = (u & 1) * 16 + grp;
                    const PairRec *recs = srec + p * GNT + (tid - lane) + l0;
                    const float4 *h4 = reinterpret_cast<const float4 *>(a.hiT) + sub;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = min(2 * (cp_warp + l0 + 4 * j) + r, K - 1);  // padding columns: the last real one
                        hbuf[j] = __ldg(h4 + ((unsigned)(row * Kp + recs[4 * j].kb) >> 2));
                    }
                };
                auto consume = [&](int u, const float4 (&hbuf)[4]) {
                    const int p = u >> 1, q = p % GQ, l0 = (u & 1) * 16 + grp;
                    PairRec *recs = srec + p * GNT + (tid - lane) + l0;
                    const float *dq = in_qk + q * Kp + 4 * sub;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const PairRec rec = recs[4 * j];
                        const float4 d = *reinterpret_cast<const float4 *>(dq + rec.kb);
                        const unsigned b0 = __ballot_sync(FULL_MASK, __fadd_rn(d.x, hbuf[j].x) >= rec.thr);
                        const unsigned b1 = __ballot_sync(FULL_MASK, __fadd_rn(d.y, hbuf[j].y) >= rec.thr);
                        const unsigned b2 = __ballot_sync(FULL_MASK, __fadd_rn(d.z, hbuf[j].z) >= rec.thr);
                        const unsigned b3 = __ballot_sync(FULL_MASK, __fadd_rn(d.w, hbuf[j].w) >= rec.thr);
                        const unsigned hit = __byte_perm(__byte_perm(b0, b1, sel_lo), __byte_perm(b2, b3, sel_lo), 0x5410);
                        // every lane consumed this record before its ballots completed: safe to overwrite
                        if (sub == 0) recs[4 * j].thr = __uint_as_float(hit);  // one writer per owner: its group's first lane
                    }
                };
                {
                    constexpr int NU = 4 * GQ;  // 16 pair sets x 2 halves
                    float4 hb0[4], hb1[4], hb2[4];
                    issue(0, hb0);
                    issue(1, hb1);
                    issue(2, hb2);
                    for (int u = 0; u < NU; u += 3) {
                        consume(u, hb0);
                        if (u + 3 < NU) issue(u + 3, hb0);
                        if (u + 1 < NU) consume(u + 1, hb1);
                        if (u + 4 < NU) issue(u + 4, hb1);
                        if (u + 2 < NU) consume(u + 2, hb2);
                        if (u + 5 < NU) issue(u + 5, hb2);
                    }
                }
                __syncwarp();
                // ---- the owners evaluate their candidates exactly (F:170), four pairs at a time so that
                // the loads of the doubles overlap
                for (int p0 = 0; p0 < 2 * GQ; p0 += 4) {
                    unsigned got[4];
                    int kb[4], kk[4];
                    double la[4];
                    float tmp[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int p = p0 + e, r = p / GQ, q = p % GQ;
                        const PairRec rec = srec[p * GNT + tid];
                        got[e] = __float_as_uint(rec.thr), kb[e] = rec.kb;
                        const int bit = __ffs(got[e]) - 1;
                        kk[e] = kb[e] + 4 * (bit & 7) + (bit >> 3);
                        la[e] = 0.0, tmp[e] = 0.f;
                        if (got[e]) {
                            la[e] = __ldg(a.LAd + (unsigned)(kk[e] * K + i0 + r));
                            tmp[e] = __ldg(a.LBf + (unsigned)(obs[q] * Kp + i0 + r));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int p = p0 + e, r = p / GQ, q = p % GQ;
                        if (!got[e]) continue;  // dead or listed (threshold +inf): nothing passed
                        Best b{-FLT_MAX, 0x7fffffff};
                        float x = exact_cand(__fadd_rn(tmp[e], in_qk[q * Kp + kk[e]]), la[e]);
                        if (x > -FLT_MAX) best_take(b, x, kk[e]);
                        unsigned rest = got[e] & (got[e] - 1);
                        while (rest) {  // several candidates inside the window (best_take keeps the lowest index)
                            const int bit = __ffs(rest) - 1;
                            const int k = kb[e] + 4 * (bit & 7) + (bit >> 3);
                            rest &= rest - 1;
                            x = exact_cand(__fadd_rn(tmp[e], in_qk[q * Kp + k]), __ldg(a.LAd + (unsigned)(k * K + i0 + r)));
                            if (x > -FLT_MAX) best_take(b, x, k);
                        }
                        emit(q, i0 + r, b);
                    }
                }
                __syncwarp();  // records are rewritten in the next round
            }
            __syncthreads();
            // ---- list phase: one warp per listed pair scans the whole column --------------------------
            const int nlisted = min(*scount, GLIST_CAP);
            for (int e = warp; e < nlisted; e += GNT / 32) {
                const int q = slist[e] >> 24, i = slist[e] & 0xffffff;
                const float tmpv = __ldg(a.LBf + (size_t)obs[q] * Kp + i);
                const float4 *col4 = reinterpret_cast<const float4 *>(a.hiT + (size_t)min(i, K - 1) * Kp);
                const float4 *d4 = reinterpret_cast<const float4 *>(in_qk + (size_t)q * Kp);
                float m = -INFINITY;
                for (int t = lane; t < (Kp >> 2); t += 32) {
                    const float4 h = __ldg(col4 + t);
                    const float4 dv = d4[t];
                    m = fmaxf(fmaxf(fmaxf(m, __fadd_rn(dv.x, h.x)), __fadd_rn(dv.y, h.y)),
                              fmaxf(__fadd_rn(dv.z, h.z), __fadd_rn(dv.w, h.w)));
                }
                const float thr = filter_threshold(warp_max(m), tmpv);
                Best b{-FLT_MAX, 0x7fffffff};
                for (int t = lane; t < (Kp >> 2); t += 32) {
                    const float4 h = __ldg(col4 + t);
                    const float4 dv = d4[t];
                    const float hh[4] = {h.x, h.y, h.z, h.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k = 4 * t + c;
                        if (k < K && i < K && __fadd_rn(dd[c], hh[c]) >= thr) {
                            const float x = exact_cand(__fadd_rn(tmpv, dd[c]), __ldg(a.LAd + (size_t)k * K + i));
                            if (x > -FLT_MAX) best_take(b, x, k);
                        }
                    }
                }
                b = warp_best(b);
                if (lane == 0) emit(q, i, b);
            }
            __syncthreads();
            cur ^= 1;
        }
    }
}

bool group_engine_fits(const flashv_model *m)
{
    return m->hiS != nullptr && group_smem_bytes(m->Kp) <= (size_t)200 * 1024;
}

int group_run_pass(flashv_plan *p, const Pass &pass, float *dfinal)
{
    flashv_model *m = p->model;
    flashv_ctx *ctx = m->ctx;
    GroupArgs g;
    g.hiS = m->hiS, g.hiT = m->hiT, g.LAd = m->LAd, g.LBd = m->LBd, g.LPi = m->LPi, g.LBf = m->LBf;
    g.K = m->K, g.Kp = m->Kp;
    g.vecs = p->d_vecs + pass.vec_offset, g.nvec = pass.nvec, g.ob = p->d_ob, g.ans = p->d_ans, g.T = p->T;
    g.dfinal = dfinal, g.psi = p->d_psi, g.psi16 = p->psi16;
    const size_t smem = group_smem_bytes(m->Kp);
    FV_CUDA(cudaFuncSetAttribute(k_flash_group_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ngroups = (pass.nvec + GQ - 1) / GQ;
    const int per_sm = 2 * (smem + 1024) <= (size_t)ctx->smem_optin ? 2 : 1;
    const int grid = std::min(per_sm * ctx->sm_count, ngroups);
    k_flash_group_cols<<<grid, GNT, smem, ctx->stream>>>(g);
    FV_CUDA(cudaGetLastError());
    ++p->launches;
    return FLASHV_OK;
}

}  // namespace flashv
