/*
 * flashv_oracle.c — CPU restatement of the reference FLASH / FLASH-BS decoders.
 *
 * TEST INFRASTRUCTURE ONLY (see flashv_oracle.h).  Runtime K/M/T/N/B instead of the
 * reference's compile-time #defines, single FIFO consumer instead of a pthread pool (the
 * result is schedule-independent: every task reads only Ans[] entries written by its
 * ancestors, F:298-304 after F:291), libm log() hoisted into tables (bit-identical — same
 * libm, same arguments).  Arithmetic, comparison order and tie rules follow the cited
 * lines literally.  Parity: pinned to the real reference by tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared (oracle/Makefile).  No -ffast-math:
 * the float -> double -> float rounding chain of F:170 is the whole point.
 *
 *   F: = /root/reference/src/FLASH_Viterbi_multithread.c
 *   S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
 */
#include "flashv_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct fvo_model {
    int K, M;
    /* lean form (fvo_model_create_lean, for shapes where two K x K double tables do not fit a host):
     * per destination column the ascending list of sources k with A[k][i] > 0 and their logarithms.
     * A source with A[k][i] == 0 has ktmp = -inf, which never passes "ktmp > score" from -FLT_MAX
     * (F:167, F:171), so walking only the lists executes exactly the comparisons that can succeed, in
     * the same order.  tests/test_oracle_golden.py checks lean == literal on every golden model. */
    int lean;
    const float *A; /* lean: borrowed, must outlive the model (start vectors read row Ans[L-1], F:220) */
    long *cptr;     /* lean: [K+1] */
    int *ck;        /* lean: [nnz] source state */
    double *cla;    /* lean: [nnz] log((double)A[k][i]) */
    double *LA;   /* [k][i]  log((double)A[k][i])            F:170 */
    double *LAt;  /* [i][k]  same numbers, transposed so the k loop is unit stride */
    double *LB;   /* [i][o]  log((double)B[i][o])            F:142 */
    float *LBf;   /* [i][o]  (float)LB — "tmp = log(...)"    F:167 */
    double *LPi;  /* [i]     log((double)Pi[i])              F:142 */
};

fvo_model *fvo_model_create(int K, int M, const float *A, const float *B, const float *Pi)
{
    fvo_model *m = (fvo_model *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    m->K = K;
    m->M = M;
    m->LA = (double *)malloc(sizeof(double) * (size_t)K * K);
    m->LAt = (double *)malloc(sizeof(double) * (size_t)K * K);
    m->LB = (double *)malloc(sizeof(double) * (size_t)K * M);
    m->LBf = (float *)malloc(sizeof(float) * (size_t)K * M);
    m->LPi = (double *)malloc(sizeof(double) * (size_t)K);
    if (!m->LA || !m->LAt || !m->LB || !m->LBf || !m->LPi) {
        fvo_model_free(m);
        return NULL;
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < K; ++i) {
            double v = log((double)A[(size_t)k * K + i]);
            m->LA[(size_t)k * K + i] = v;
            m->LAt[(size_t)i * K + k] = v;
        }
    for (int i = 0; i < K; ++i) {
        for (int o = 0; o < M; ++o) {
            double v = log((double)B[(size_t)i * M + o]);
            m->LB[(size_t)i * M + o] = v;
            m->LBf[(size_t)i * M + o] = (float)v;
        }
        m->LPi[i] = log((double)Pi[i]);
    }
    return m;
}

fvo_model *fvo_model_create_lean(int K, int M, const float *A, const float *B, const float *Pi)
{
    fvo_model *m = (fvo_model *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    m->K = K;
    m->M = M;
    m->lean = 1;
    m->A = A;
    m->cptr = (long *)calloc((size_t)K + 1, sizeof(long));
    m->LB = (double *)malloc(sizeof(double) * (size_t)K * M);
    m->LBf = (float *)malloc(sizeof(float) * (size_t)K * M);
    m->LPi = (double *)malloc(sizeof(double) * (size_t)K);
    if (!m->cptr || !m->LB || !m->LBf || !m->LPi) {
        fvo_model_free(m);
        return NULL;
    }
    /* every thread owns a range of columns: it reads that slice of every row, so its lists come out in
     * ascending k without any coordination */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < K; ++i) {
        long c = 0;
        for (int k = 0; k < K; ++k) c += A[(size_t)k * K + i] > 0.0f;
        m->cptr[i + 1] = c;
    }
    for (int i = 0; i < K; ++i) m->cptr[i + 1] += m->cptr[i];
    const long nnz = m->cptr[K];
    m->ck = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    m->cla = (double *)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    if (!m->ck || !m->cla) {
        fvo_model_free(m);
        return NULL;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < K; ++i) {
        long e = m->cptr[i];
        for (int k = 0; k < K; ++k) {
            const float v = A[(size_t)k * K + i];
            if (v > 0.0f) {
                m->ck[e] = k;
                m->cla[e] = log((double)v);
                ++e;
            }
        }
    }
    for (int i = 0; i < K; ++i) {
        for (int o = 0; o < M; ++o) {
            double v = log((double)B[(size_t)i * M + o]);
            m->LB[(size_t)i * M + o] = v;
            m->LBf[(size_t)i * M + o] = (float)v;
        }
        m->LPi[i] = log((double)Pi[i]);
    }
    return m;
}

void fvo_model_free(fvo_model *m)
{
    if (!m) return;
    free(m->cptr);
    free(m->ck);
    free(m->cla);
    free(m->LA);
    free(m->LAt);
    free(m->LB);
    free(m->LBf);
    free(m->LPi);
    free(m);
}

/* ---------------------------------------------------------------- dense FLASH pieces */

/* F:142 / F:212 (prev_state < 0: pi form) and F:150 / F:220 (restart from a fixed state).
 * Both are double + double, rounded once to float on the store into T1. */
void fvo_flash_init(const fvo_model *m, int prev_state, int o, float *d_out)
{
    const int K = m->K, M = m->M;
    for (int i = 0; i < K; ++i) {
        double head = prev_state < 0 ? m->LPi[i]
                      : m->lean    ? log((double)m->A[(size_t)prev_state * K + i])
                                   : m->LA[(size_t)prev_state * K + i];
        d_out[i] = (float)(head + m->LB[(size_t)i * M + o]);
    }
}

/* F:165-174 (== F:231-241).  ktmp = tmp + T1[k] + log(A[k][i]): float add, promoted to
 * double for the second add, rounded to float by the assignment; strict '>' from
 * (-FLT_MAX, -1) so the lowest k among equal maxima wins and dead columns give arg -1. */
void fvo_flash_step(const fvo_model *m, const float *d_in, int o, float *d_out, int *psi)
{
    const int K = m->K, M = m->M;
    if (m->lean) {
#pragma omp parallel for schedule(dynamic, 64)
        for (int i = 0; i < K; ++i) {
            const float tmp = m->LBf[(size_t)i * M + o];
            float best = -FLT_MAX;
            int arg = -1;
            for (long e = m->cptr[i]; e < m->cptr[i + 1]; ++e) {
                float pre = tmp + d_in[m->ck[e]];
                double wide = (double)pre + m->cla[e];
                float cand = (float)wide;
                if (cand > best) {
                    best = cand;
                    arg = m->ck[e];
                }
            }
            d_out[i] = best;
            psi[i] = arg;
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < K; ++i) {
        const double *col = m->LAt + (size_t)i * K;
        const float tmp = m->LBf[(size_t)i * M + o];
        float best = -FLT_MAX;
        int arg = -1;
        for (int k = 0; k < K; ++k) {
            float pre = tmp + d_in[k];
            double wide = (double)pre + col[k];
            float cand = (float)wide;
            if (cand > best) {
                best = cand;
                arg = k;
            }
        }
        d_out[i] = best;
        psi[i] = arg;
    }
}

/* F:129-136: segment boundaries of the N-way pass over (L,R). */
static void split_points(int L, int R, int N, int *mids)
{
    int gap = (R - L) / N, extra = (R - L) % N;
    int at = L;
    for (int t = 0; t + 1 < N; ++t) {
        at += gap;
        if (extra) {
            --extra;
            ++at;
        }
        mids[t] = at;
    }
}

static int first_pass_runs(int T, int N) { return N > 2 && T >= 2 * N; } /* F:342 */

int fvo_task_list(int T, int N, int *L, int *R, int *first_pass, int *mids)
{
    int head = 0, tail = 0; /* tail = next free slot, head = next to pop */
    int fp = first_pass_runs(T, N);
    if (first_pass) *first_pass = fp;
    if (fp) {
        int *mp = mids ? mids : (int *)malloc(sizeof(int) * (size_t)(N - 1));
        split_points(0, T - 1, N, mp);
        /* F:349-353 */
        L[tail] = 0, R[tail] = mp[0], ++tail;
        for (int t = 0; t + 2 < N; ++t) L[tail] = mp[t] + 1, R[tail] = mp[t + 1], ++tail;
        L[tail] = mp[N - 2] + 1, R[tail] = T - 1, ++tail;
        if (!mids) free(mp);
    } else {
        L[tail] = 0, R[tail] = T - 1, ++tail; /* F:357-359 */
    }
    /* F:284-304: pop, split at mid, push (L,mid) and (mid+1,R) when non-trivial.  The
     * reference stops after popping queue index T-2, i.e. after T-N (or T-1) tasks. */
    const int want = fp ? T - N : T - 1;
    while (head < tail && head < want) {
        int l = L[head], r = R[head];
        ++head;
        if (r <= l + 1) continue;
        int mid = (l + r) >> 1;
        if (tail < T) L[tail] = l, R[tail] = mid, ++tail;
        if (r > mid + 1 && tail < T) L[tail] = mid + 1, R[tail] = r, ++tail;
    }
    return head;
}

long fvo_executed_steps(int T, int N)
{
    int *L = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int *R = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int fp = 0;
    int n = fvo_task_list(T, N, L, R, &fp, NULL);
    long s = fp ? T - 1 : 0;
    for (int q = 0; q < n; ++q) s += R[q] - L[q];
    free(L);
    free(R);
    return s;
}

/* sizeof(ThreadPool) of F:36-43 / S:38-45 on LP64 glibc. */
static int pool_bytes(int N)
{
    size_t raw = sizeof(pthread_mutex_t) + sizeof(pthread_cond_t) + sizeof(pthread_t) * (size_t)N +
                 3 * sizeof(int);
    return (int)((raw + 7u) & ~(size_t)7u);
}

int fvo_flash_memory_bytes(int K, int T, int N)
{
    int mem = 0;
    if (first_pass_runs(T, N)) /* F:355 */
        mem = (int)(sizeof(int) * (size_t)(N - 1) + sizeof(float) * 2u * K +
                    sizeof(int) * 2u * (size_t)(N - 1) * K);
    int per = N * (int)(2 * K * sizeof(float) + 2 * K * sizeof(int)); /* F:364 */
    if (per > mem) mem = per;
    return mem + pool_bytes(N) + (int)sizeof(size_t); /* F:367 */
}

int fvo_bs_memory_bytes(int T, int N, int Bw)
{
    int mem = 0;
    if (first_pass_runs(T, N)) /* S:564 */
        mem = (int)(sizeof(int) * (size_t)(N - 1) + 12u * 2u * (size_t)(N - 1) * (Bw + 1));
    int per = N * (int)(2 * (Bw + 1) * 12); /* S:573 */
    if (per > mem) mem = per;
    return mem + pool_bytes(N) + (int)sizeof(size_t); /* S:576 */
}

static int first_argmax(const float *d, int K, float *score)
{
    /* F:188-193 */
    float s = d[0];
    int arg = 0;
    for (int i = 1; i < K; ++i)
        if (d[i] > s) s = d[i], arg = i;
    if (score) *score = s;
    return arg;
}

/* nvviterNdivide, F:126-202, for (L,R) = (0,T-1). */
static void flash_first_pass(const fvo_model *m, const int *ob, int T, int N, const int *mids,
                             int *ans, float *score)
{
    const int K = m->K;
    float *d[2];
    int *trk[2], *psi;
    d[0] = (float *)malloc(sizeof(float) * (size_t)K);
    d[1] = (float *)malloc(sizeof(float) * (size_t)K);
    trk[0] = (int *)calloc((size_t)(N - 1) * K, sizeof(int));
    trk[1] = (int *)calloc((size_t)(N - 1) * K, sizeof(int));
    psi = (int *)malloc(sizeof(int) * (size_t)K);
    fvo_flash_init(m, -1, ob[0], d[0]);
    int cur = 0, p = -1;
    for (int j = 1; j <= T - 1; ++j) {
        while (p + 2 < N && j > mids[p + 1] + 1) ++p; /* F:163 */
        fvo_flash_step(m, d[cur], ob[j], d[cur ^ 1], psi);
        for (int t = 0; t + 1 < N; ++t) { /* F:176-179 */
            int *dst = trk[cur ^ 1] + (size_t)t * K;
            const int *src = trk[cur] + (size_t)t * K;
            if (t <= p)
                for (int i = 0; i < K; ++i) dst[i] = psi[i] < 0 ? -1 : src[psi[i]];
            else
                memcpy(dst, psi, sizeof(int) * (size_t)K);
        }
        cur ^= 1;
    }
    int end = first_argmax(d[cur], K, score);
    ans[T - 1] = end;
    for (int t = 0; t + 1 < N; ++t) ans[mids[t]] = trk[cur][(size_t)t * K + end]; /* F:198-201 */
    free(d[0]);
    free(d[1]);
    free(trk[0]);
    free(trk[1]);
    free(psi);
}

/* nvviter, F:204-262. */
static void flash_task(const fvo_model *m, const int *ob, int T, int L, int R, int mid, int *ans,
                       float *score, float *d0, float *d1, int *t0, int *t1, int *psi)
{
    const int K = m->K;
    float *d[2] = {d0, d1};
    int *trk[2] = {t0, t1};
    const int prev = L == 0 ? -1 : ans[L - 1];
    fvo_flash_init(m, prev, ob[L], d[0]);
    for (int i = 0; i < K; ++i) trk[0][i] = prev; /* F:221 (unset when L == 0; never read) */
    int cur = 0;
    for (int j = L + 1; j <= R; ++j) {
        fvo_flash_step(m, d[cur], ob[j], d[cur ^ 1], psi);
        if (j > mid + 1) /* F:242 */
            for (int i = 0; i < K; ++i) trk[cur ^ 1][i] = psi[i] < 0 ? -1 : trk[cur][psi[i]];
        else
            memcpy(trk[cur ^ 1], psi, sizeof(int) * (size_t)K);
        cur ^= 1;
    }
    int end = ans[R];
    if (L == 0 && R == T - 1) { /* F:249-259 */
        end = first_argmax(d[cur], K, score);
        ans[R] = end;
    }
    ans[mid] = trk[cur][end]; /* F:261 */
}

int fvo_flash_decode(const fvo_model *m, const int *ob, int T, int N, int *path, float *score,
                     int *memory_bytes)
{
    const int K = m->K;
    if (T < 2 || N < 1) return -1;
    if (N > 2 && T == 2 * N) return -1; /* SURVEY §8a "edge": reference leaves Ans[] unset */
    int *L = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int *R = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int *mids = (int *)malloc(sizeof(int) * (size_t)(N > 1 ? N : 1));
    int fp = 0;
    int ntask = fvo_task_list(T, N, L, R, &fp, mids);
    float sc = -FLT_MAX;
    for (int j = 0; j < T; ++j) path[j] = -2; /* poison: every entry must be written */
    if (fp) flash_first_pass(m, ob, T, N, mids, path, &sc);
    float *d0 = (float *)malloc(sizeof(float) * (size_t)K);
    float *d1 = (float *)malloc(sizeof(float) * (size_t)K);
    int *t0 = (int *)malloc(sizeof(int) * (size_t)K);
    int *t1 = (int *)malloc(sizeof(int) * (size_t)K);
    int *psi = (int *)malloc(sizeof(int) * (size_t)K);
    for (int q = 0; q < ntask; ++q) /* F:284-291, FIFO order */
        flash_task(m, ob, T, L[q], R[q], (L[q] + R[q]) >> 1, path, &sc, d0, d1, t0, t1, psi);
    if (score) *score = sc;
    if (memory_bytes) *memory_bytes = fvo_flash_memory_bytes(K, T, N);
    free(L), free(R), free(mids), free(d0), free(d1), free(t0), free(t1), free(psi);
    return 0;
}

/* ---------------------------------------------------------------- vanilla Viterbi (sanity path) */

/* viterbi() of "Base_line/C implementations/vanilla Viterbi.c":124-171 (V: below).  NOT the FLASH arithmetic:
 * the candidate is  T1[k][j-1] + log(A[k][i]) + log(B[i][o_j])  (V:140) — the float promoted to double, two
 * double adds left to right, one rounding to float on the store; strict '>' from (-FLT_MAX, -1) (V:136-145);
 * last column: first maximum from (-FLT_MAX, -1) (V:152-161); backtrack through the full T2 table (V:167-170).
 * memory = sizeof(T1)+sizeof(T2) (V:172).  Returns -1 where the reference walks off its tables (a dead column
 * on the path: arg = -1, V:147/163 only print an error).  Needs the full (non-lean) model. */
int fvo_vanilla_decode(const fvo_model *m, const int *ob, int T, int *path, float *score, int *memory_bytes)
{
    const int K = m->K, M = m->M;
    if (T < 1 || m->lean) return -1;
    float *d0 = (float *)malloc(sizeof(float) * (size_t)K), *d1 = (float *)malloc(sizeof(float) * (size_t)K);
    int *T2 = (int *)malloc(sizeof(int) * (size_t)K * (size_t)T);
    for (int i = 0; i < K; ++i) d0[i] = (float)(m->LPi[i] + m->LB[(size_t)i * M + ob[0]]); /* V:120, V:128 */
    for (int j = 1; j < T; ++j) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < K; ++i) {
            const double *col = m->LAt + (size_t)i * K;
            const double lb = m->LB[(size_t)i * M + ob[j]];
            float best = -FLT_MAX;
            int arg = -1;
            for (int k = 0; k < K; ++k) {
                float cand = (float)(((double)d0[k] + col[k]) + lb); /* V:140 */
                if (cand > best) {
                    best = cand;
                    arg = k;
                }
            }
            d1[i] = best;
            T2[(size_t)i * T + j] = arg;
        }
        float *t = d0;
        d0 = d1, d1 = t;
    }
    float best = -FLT_MAX;
    int arg = -1;
    for (int i = 0; i < K; ++i)
        if (d0[i] > best) best = d0[i], arg = i; /* V:154-160 */
    int rc = arg < 0 ? -1 : 0;
    if (rc == 0) {
        path[T - 1] = arg;
        for (int j = T - 1; j > 0; --j) { /* V:167-170 */
            const int prev = T2[(size_t)path[j] * T + j];
            if (prev < 0) {
                rc = -1;
                break;
            }
            path[j - 1] = prev;
        }
    }
    if (score) *score = best;
    if (memory_bytes) *memory_bytes = (int)(sizeof(float) * (size_t)K * T + sizeof(int) * (size_t)K * T); /* V:172 */
    free(d0), free(d1), free(T2);
    return rc;
}

/* ---------------------------------------------------------------- FLASH-BS pieces */

typedef struct {
    float v; /* slot 0: element count, as in S:75 */
    int state;
    int pay;
} slot_t; /* S:51-56 */

static void heap_reset(slot_t *h) /* S:65-70 */
{
    h[0].v = 0;
    h[0].state = -1;
    h[0].pay = -1;
}

static void heap_floyd(slot_t *h) /* S:96-123 */
{
    int total = (int)h[0].v;
    for (int node = total / 2; node > 0; --node) {
        int parent = node, child = 2 * node;
        slot_t held = h[parent];
        for (; child <= total; child *= 2) {
            if (child + 1 <= total && h[child].v > h[child + 1].v) ++child;
            if (held.v <= h[child].v) break;
            h[parent] = h[child];
            parent = child;
        }
        h[parent] = held;
    }
}

static void heap_replace_root(slot_t *h, float v, int state, int pay) /* S:131-165 */
{
    h[1].v = v, h[1].state = state, h[1].pay = pay;
    int total = (int)h[0].v, parent = 1, child = 2;
    while (child <= total) {
        if (child + 1 <= total && h[child].v > h[child + 1].v) ++child;
        if (h[parent].v <= h[child].v) break;
        slot_t sw = h[parent];
        h[parent] = h[child];
        h[child] = sw;
        parent = child;
        child *= 2;
    }
}

static void heap_feed(slot_t *h, int Bw, float v, int i, int pay) /* S:167-211 */
{
    if (i < Bw) {
        h[i + 1].v = v, h[i + 1].state = i, h[i + 1].pay = pay;
        h[0].v += 1;
        if (i == Bw - 1) heap_floyd(h);
    } else if (v > h[1].v) {
        heap_replace_root(h, v, i, pay);
    }
}

void fvo_bs_heap_replay(int K, int Bw, const float *score, const int *payload, float *hval,
                        int *hstate, int *hpay)
{
    slot_t *h = (slot_t *)malloc(sizeof(slot_t) * (size_t)(Bw + 1));
    heap_reset(h);
    for (int i = 0; i < K; ++i) heap_feed(h, Bw, score[i], i, payload ? payload[i] : -1);
    for (int s = 0; s < Bw; ++s) hval[s] = h[s + 1].v, hstate[s] = h[s + 1].state, hpay[s] = h[s + 1].pay;
    free(h);
}

/* Row of log A for a beam predecessor.  state < 0 reproduces vit->A[-1][i] == vit->Pi[i]
 * (Pi[K] sits directly in front of A[K][K] in VIT, S:27-30). */
static inline double la_from(const fvo_model *m, int state, int i)
{
    return state < 0 ? m->LPi[i] : m->LA[(size_t)state * m->K + i];
}

/* S:437-446 for every destination i: max over beam slots in array order, strict '>'. */
static void bs_scores(const fvo_model *m, const slot_t *h, int Bw, int o, float *score, int *arg_slot)
{
    const int K = m->K, M = m->M;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < K; ++i) {
        const float tmp = m->LBf[(size_t)i * M + o];
        float best = -FLT_MAX;
        int arg = -1;
        for (int c = 0; c < Bw; ++c) {
            float pre = tmp + h[c + 1].v;
            double wide = (double)pre + la_from(m, h[c + 1].state, i);
            float cand = (float)wide;
            if (cand > best) best = cand, arg = c;
        }
        score[i] = best;
        arg_slot[i] = arg;
    }
}

void fvo_bs_score_step(const fvo_model *m, const float *hval, const int *hstate, int Bw, int o,
                       float *score, int *arg_slot)
{
    slot_t *h = (slot_t *)malloc(sizeof(slot_t) * (size_t)(Bw + 1));
    heap_reset(h);
    h[0].v = (float)Bw;
    for (int s = 0; s < Bw; ++s) h[s + 1].v = hval[s], h[s + 1].state = hstate[s], h[s + 1].pay = -1;
    bs_scores(m, h, Bw, o, score, arg_slot);
    free(h);
}

static int bs_end_scan(const slot_t *h, int Bw, float *score) /* S:376-381 / S:456-461 */
{
    float s = h[1].v;
    int arg = 0;
    for (int c = Bw / 2 + 1; c < Bw; ++c)
        if (h[c + 1].v > s) arg = c, s = h[c + 1].v;
    if (score) *score = s;
    return arg;
}

static int bs_find_payload(const slot_t *h, int state) /* S:73-86 */
{
    int total = (int)h[0].v;
    for (int c = 1; c <= total; ++c)
        if (h[c].state == state) return h[c].pay;
    return -1;
}

/* nvviterNdivide, S:295-399, for (L,R) = (0,T-1): N-1 heaps with identical (Value,State)
 * layout and per-tracker payloads; scores are read from heap #1 (S:352-353). */
static void bs_first_pass(const fvo_model *m, const int *ob, int T, int N, int Bw, const int *mids,
                          int *ans, float *score)
{
    const int K = m->K, M = m->M, H = N - 1;
    const size_t hs = (size_t)Bw + 1;
    slot_t *heap[2];
    heap[0] = (slot_t *)calloc(hs * H, sizeof(slot_t));
    heap[1] = (slot_t *)calloc(hs * H, sizeof(slot_t));
    float *sc = (float *)malloc(sizeof(float) * (size_t)K);
    int *as = (int *)malloc(sizeof(int) * (size_t)K);
    for (int t = 0; t < H; ++t) heap_reset(heap[0] + t * hs);
    for (int i = 0; i < K; ++i) { /* S:314-320 */
        float v = (float)(m->LPi[i] + m->LB[(size_t)i * M + ob[0]]);
        for (int t = 0; t < H; ++t) heap_feed(heap[0] + t * hs, Bw, v, i, -1);
    }
    int cur = 0, p = -1;
    for (int j = 1; j <= T - 1; ++j) {
        while (p + 2 < N && j > mids[p + 1] + 1) ++p; /* S:342 */
        for (int t = 0; t < H; ++t) heap_reset(heap[cur ^ 1] + t * hs);
        bs_scores(m, heap[cur] + 1 * hs, Bw, ob[j], sc, as);
        for (int i = 0; i < K; ++i)
            for (int t = 0; t < H; ++t) { /* S:359-368; slot 0 carries (-1,-1) for arg -1 */
                const slot_t *from = heap[cur] + t * hs + (as[i] + 1);
                heap_feed(heap[cur ^ 1] + t * hs, Bw, sc[i], i, t <= p ? from->pay : from->state);
            }
        cur ^= 1;
    }
    int arg = bs_end_scan(heap[cur] + 1 * hs, Bw, score); /* S:376-381 */
    ans[T - 1] = heap[cur][1 * hs + arg + 1].state;
    for (int t = 0; t < H; ++t) ans[mids[t]] = heap[cur][t * hs + arg + 1].pay;
    free(heap[0]), free(heap[1]), free(sc), free(as);
}

/* nvviter, S:401-473. */
static void bs_task(const fvo_model *m, const int *ob, int T, int Bw, int L, int R, int mid, int *ans,
                    float *score, slot_t *h0, slot_t *h1, float *sc, int *as)
{
    const int K = m->K, M = m->M;
    slot_t *heap[2] = {h0, h1};
    const int prev = L == 0 ? -1 : ans[L - 1];
    heap_reset(heap[0]);
    for (int i = 0; i < K; ++i) { /* S:411-426 */
        float v = (float)(la_from(m, prev, i) + m->LB[(size_t)i * M + ob[L]]);
        heap_feed(heap[0], Bw, v, i, -1);
    }
    int cur = 0;
    for (int j = L + 1; j <= R; ++j) {
        heap_reset(heap[cur ^ 1]);
        bs_scores(m, heap[cur], Bw, ob[j], sc, as);
        for (int i = 0; i < K; ++i) { /* S:447-448 */
            const slot_t *from = heap[cur] + (as[i] + 1);
            heap_feed(heap[cur ^ 1], Bw, sc[i], i, j > mid + 1 ? from->pay : from->state);
        }
        cur ^= 1;
    }
    if (L == 0 && R == T - 1) { /* S:454-465 */
        int arg = bs_end_scan(heap[cur], Bw, score);
        ans[R] = heap[cur][arg + 1].state;
        ans[mid] = heap[cur][arg + 1].pay;
    } else {
        ans[mid] = bs_find_payload(heap[cur], ans[R]); /* S:466-471 */
    }
}

int fvo_bs_decode(const fvo_model *m, const int *ob, int T, int N, int Bw, int *path, float *score,
                  int *memory_bytes)
{
    const int K = m->K;
    if (T < 2 || N < 1 || Bw < 1 || Bw > K) return -1;
    if (N > 2 && T == 2 * N) return -1;
    int *L = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int *R = (int *)malloc(sizeof(int) * (size_t)(T + 2));
    int *mids = (int *)malloc(sizeof(int) * (size_t)(N > 1 ? N : 1));
    int fp = 0;
    int ntask = fvo_task_list(T, N, L, R, &fp, mids);
    float scv = -FLT_MAX;
    for (int j = 0; j < T; ++j) path[j] = -2;
    if (fp) bs_first_pass(m, ob, T, N, Bw, mids, path, &scv);
    slot_t *h0 = (slot_t *)calloc((size_t)Bw + 1, sizeof(slot_t));
    slot_t *h1 = (slot_t *)calloc((size_t)Bw + 1, sizeof(slot_t));
    float *sc = (float *)malloc(sizeof(float) * (size_t)K);
    int *as = (int *)malloc(sizeof(int) * (size_t)K);
    for (int q = 0; q < ntask; ++q)
        bs_task(m, ob, T, Bw, L[q], R[q], (L[q] + R[q]) >> 1, path, &scv, h0, h1, sc, as);
    if (score) *score = scv;
    if (memory_bytes) *memory_bytes = fvo_bs_memory_bytes(T, N, Bw);
    free(L), free(R), free(mids), free(h0), free(h1), free(sc), free(as);
    return 0;
}

/* ---------------------------------------------------------------- text ingest */

/* F:82-93: the reference reads every probability with fscanf("%f") straight into a float
 * (strtof semantics, one rounding from the decimal text).  Returns the count read. */
long fvo_read_floats(const char *path, long n, float *out)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    long got = 0;
    while (got < n && fscanf(fp, "%f", &out[got]) == 1) ++got;
    fclose(fp);
    return got;
}

long fvo_read_ints(const char *path, long n, int *out) /* F:76-77 */
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    long got = 0;
    while (got < n && fscanf(fp, "%d", &out[got]) == 1) ++got;
    fclose(fp);
    return got;
}
