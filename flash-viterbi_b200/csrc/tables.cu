// tables.cu — log tables of one HMM: host libm -> device layouts.
//
// The reference calls libm log() inside its innermost loop (F:170; also F:142,150,167,212,220,
// 233,236).  glibc's log is not correctly rounded and CUDA's differs from it in the last bit on
// some inputs, so a bit-exact decoder must take its logarithms from the same host libm the
// reference links.  They depend only on the model, so they are computed ONCE here (all host
// cores), uploaded, and re-laid-out on the device:
//   LAd [k][i] double  source-major, as the reference stores A (F:27): exact re-evaluation and
//                      start vectors (F:220 reads row Ans[L-1])
//   hiT [i][k] float   destination-major, padded to Kp: the stream the trellis kernels read —
//                      the max over k for one destination i is one contiguous run
//   hiC        float   the same numbers CTA-tiled for the persistent engine (tile_geom.h)
//   LBf [o][i] float   the per-step "tmp" (F:167), symbol-major so one step reads one row
//   LBd [o][i] double  start vectors (F:142, F:220)
//   LPi [i]    double
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "flashv_internal.h"
#include "tile_geom.h"

namespace flashv {

// hiT[i][k] = (float)LAd[k][i] for k < K, -inf in the padding; 32x32 tiles through shared memory
// so both the double reads (along i) and the float writes (along k) are coalesced.
__global__ void k_transpose_to_f32(const double *__restrict__ LAd, float *__restrict__ hiT, int K, int Kp)
{
    __shared__ float tile[32][33];
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, i = i0 + threadIdx.x;
        float v = -INFINITY;
        if (k < K && i < K) v = __double2float_rn(LAd[(size_t)k * K + i]);
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, k = k0 + threadIdx.x;
        if (i < K && k < Kp) hiT[(size_t)i * Kp + k] = tile[threadIdx.x][r];
    }
}

// hiC: the same numbers CTA-tiled for the persistent engine (tile_geom.h).  One thread per (k, i),
// reads coalesced along i from the hiT-independent double table, scattered 4-byte writes.
// The slice form serves the state-sharded pass: only the columns [col_begin, col_begin+ncol),
// tiled for a grid of G CTAs over those ncol columns.
__global__ void k_build_tiled(const double *__restrict__ LAd, float *__restrict__ hiC, int K, int Kp, int col_begin,
                              int ncol, int G)
{
    const int il = blockIdx.x * blockDim.x + threadIdx.x;  // column index inside the slice
    const int k = blockIdx.y;
    if (il >= ncol) return;
    const float v = k < K ? __double2float_rn(LAd[(size_t)k * K + col_begin + il]) : -INFINITY;
    hiC[tile_off(ncol, Kp, G, il, k)] = v;
}

// hiS[k][i] = (float)LAd[k][i], Kp x Kp with -inf padding: the group engine's sweep reads it row by row.
__global__ void k_build_source_major(const double *__restrict__ LAd, float *__restrict__ hiS, int K, int Kp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i >= Kp) return;
    hiS[(size_t)k * Kp + i] = (k < K && i < K) ? __double2float_rn(LAd[(size_t)k * K + i]) : -INFINITY;
}

// LAc: the double table chain-major, for the persistent engine's window scan (flash_persistent.cu): chain
// q = k & 127 of column i holds its up to 32 source states k = q + 128 u next to each other,
// LAc[(i * 128 + q) * 32 + u], -inf beyond K — 32 KB per column, models up to 4096 states.
__global__ void k_build_chains(const double *__restrict__ LAd, double *__restrict__ LAc, int K)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // coalesced reads along i, scattered 8-byte writes
    const int k = blockIdx.y;                             // 0 .. 4095
    if (i >= K) return;
    LAc[((size_t)i * 128 + (k & 127)) * 32 + (k >> 7)] = k < K ? LAd[(size_t)k * K + i] : -INFINITY;
}

// The tables of the half-precision filter (k_flash_persist16, flash_persistent.cu), models up to 4096 states:
// hi16 = (half)log A, rounded from the double directly, tiled [cta][iteration of 256 states][column][256] with
// the column partition of tile_geom.h and -inf beyond K; LAc16 = log A chain-major for 256 chains of 16.
__global__ void k_build_tiled16(const double *__restrict__ LAd, __half *__restrict__ hi16, double *__restrict__ LAc16, int K,
                                int Kp16, int G)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // coalesced reads along i, scattered writes
    const int k = blockIdx.y;                             // 0 .. 4095
    if (i >= K) return;
    const double la = k < K ? LAd[(size_t)k * K + i] : -INFINITY;
    LAc16[((size_t)i * 256 + (k & 255)) * 16 + (k >> 8)] = la;
    if (k < Kp16) {
        const int b = tile_owner(K, G, i), c0 = tile_c0(K, G, b), ncols = tile_c0(K, G, b + 1) - c0;
        hi16[(size_t)c0 * Kp16 + ((size_t)(k >> 8) * ncols + (i - c0)) * 256 + (k & 255)] = __double2half(la);
    }
}

// LAcL: the chain-major double table of models wider than 4096 states (flash_persistent.cu: scan_long): 128 chains
// of clp = Kp/128 rounded up to 32 elements per destination column, -inf padding.
__global__ void k_build_long_chains(const double *__restrict__ LAd, double *__restrict__ LAcL, int K, int clp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // coalesced reads along i, scattered 8-byte writes
    const int k = blockIdx.y;                             // 0 .. 128 * clp - 1
    if (i >= K) return;
    LAcL[((size_t)i * 128 + (k & 127)) * clp + (k >> 7)] = k < K ? LAd[(size_t)k * K + i] : -INFINITY;
}

void build_tiled_slice(const double *LAd, float *hiC, int K, int Kp, int col_begin, int ncol, int G, cudaStream_t st)
{
    k_build_tiled<<<dim3((ncol + 255) / 256, Kp), 256, 0, st>>>(LAd, hiC, K, Kp, col_begin, ncol, G);
}

static bool prep_trace() { return getenv("FLASHV_PREP_TRACE") != nullptr; }
static double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

static int host_threads(int share)
{
    int nthr = 0;
    if (const char *e = getenv("FLASHV_HOST_THREADS")) nthr = atoi(e);
    if (nthr < 1) {
        unsigned hw = std::thread::hardware_concurrency();
        nthr = (int)(hw ? hw : 8) / (share > 1 ? share : 1);  // ranks of one box share its cores
    }
    return nthr < 1 ? 1 : (nthr > 64 ? 64 : nthr);
}

// Device allocations of the tables that exist for every model (the layouts derived from LAd are
// allocated by tables_layouts()).
int tables_alloc(flashv_model *m)
{
    const int K = m->K, M = m->M, Kp = m->Kp;
    const size_t nA = (size_t)K * K;
    FV_CUDA(cudaMalloc(&m->LAd, nA * sizeof(double)));
    FV_CUDA(cudaMalloc(&m->LBf, (size_t)M * Kp * sizeof(float)));
    FV_CUDA(cudaMalloc(&m->LBmax, (size_t)M * sizeof(float)));
    FV_CUDA(cudaMalloc(&m->LBd, (size_t)M * K * sizeof(double)));
    FV_CUDA(cudaMalloc(&m->LPi, (size_t)K * sizeof(double)));
    FV_CUDA(cudaMalloc(&m->scratch_f, (size_t)4 * Kp * sizeof(float)));
    FV_CUDA(cudaMalloc(&m->scratch_i, (size_t)2 * Kp * sizeof(int32_t)));
    FV_CUDA(cudaMalloc(&m->scratch_x, (size_t)2 * Kp * 8));
    FV_CUDA(cudaMemsetAsync(m->scratch_f, 0, (size_t)4 * Kp * sizeof(float), m->ctx->stream));
    m->bytes += nA * sizeof(double) + (size_t)M * Kp * 4 + (size_t)M * K * 8 + (size_t)K * 8 + (size_t)6 * Kp * 4 +
               (size_t)2 * Kp * 8;
    return FLASHV_OK;
}

// Host libm logarithms of the rows [row_lo, row_hi) of A (F:170) and of B and Pi (F:142, F:167),
// uploaded as they are produced: the host threads fill one pinned staging buffer while the copy
// engine drains the other, so neither the whole double table (8.6 GB at K=32768) nor a pinned
// allocation of that size ever exists on the host.  Validation (every entry a probability) rides
// along in the same pass over A.
int tables_logs(flashv_model *m, const float *A, const float *B, const float *Pi, int row_lo, int row_hi, int share)
{
    const int K = m->K, M = m->M, Kp = m->Kp;
    flashv_ctx *ctx = m->ctx;
    auto t0 = std::chrono::steady_clock::now();

    // The window filter of the trellis kernels relies on every log being <= 0;
    // anything else is not a probability table.  The reference would decode garbage or NaN.
    auto unit = [](float v) { return v >= 0.0f && v <= 1.0f; };
    bool ok = true;
    for (size_t t = 0; t < (size_t)K * M && ok; ++t) ok = unit(B[t]);
    for (int t = 0; t < K && ok; ++t) ok = unit(Pi[t]);
    if (!ok) {
        set_error("flashv_model_create: A/B/Pi entries must be finite and inside [0,1]");
        return FLASHV_ERR_DOMAIN;
    }

    const size_t row_bytes = (size_t)K * sizeof(double);
    constexpr size_t STAGE_BYTES = (size_t)16 << 20;
    int chunk_rows = (int)(STAGE_BYTES / row_bytes);
    if (chunk_rows < 1) chunk_rows = 1;
    const size_t stage_bytes = (size_t)chunk_rows * row_bytes;
    if (ctx->h_prep_bytes < stage_bytes) {
        for (auto &b : ctx->h_prep)
            if (b) cudaFreeHost(b), b = nullptr;
        ctx->h_prep_bytes = 0;
        for (auto &b : ctx->h_prep)
            if (cudaMallocHost(&b, stage_bytes) != cudaSuccess) {
                cudaGetLastError();
                set_error("flashv_model_create: pinned staging allocation of %zu bytes failed", stage_bytes);
                return FLASHV_ERR_NOMEM;
            }
        ctx->h_prep_bytes = stage_bytes;
    }
    const int nthr = host_threads(share);
    if (prep_trace()) fprintf(stderr, "[flashv prep] checks + pinned staging: %.2f ms, %d host threads\n", ms_since(t0), nthr);
    std::atomic<int> bad{0};
    cudaEvent_t drained[2] = {ctx->ev_prep[0], ctx->ev_prep[1]};
    int nchunk = 0;
    for (int r0 = row_lo; r0 < row_hi; r0 += chunk_rows, ++nchunk) {
        const int r1 = r0 + chunk_rows < row_hi ? r0 + chunk_rows : row_hi;
        double *stage = reinterpret_cast<double *>(ctx->h_prep[nchunk & 1]);
        if (nchunk >= 2) FV_CUDA(cudaEventSynchronize(drained[nchunk & 1]));  // the copy that last read this buffer
        const long long cells = (long long)(r1 - r0) * K;
        int use = nthr;
        if ((long long)use * 4096 > cells) use = (int)(cells / 4096) + 1;  // tiny models: not worth the threads
        auto work = [&](int t) {
            const long long c0 = cells * t / use, c1 = cells * (t + 1) / use;
            const float *src = A + (size_t)r0 * K;
            bool good = true;
            for (long long c = c0; c < c1; ++c) {
                const float v = src[c];
                good &= (v >= 0.0f && v <= 1.0f);
                // F:170.  log(+0) is -inf by definition (C Annex F), but glibc reaches it through its error
                // path (errno, FE_DIVBYZERO) at several times the cost of a regular call — and 1-p of a
                // data_script.py table is exactly 0.
                stage[c] = v == 0.0f ? -INFINITY : log((double)v);
            }
            if (!good) bad.store(1);
        };
        if (use <= 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 1; t < use; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto &th : pool) th.join();
        }
        FV_CUDA(cudaMemcpyAsync(m->LAd + (size_t)r0 * K, stage, (size_t)cells * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
        FV_CUDA(cudaEventRecord(drained[nchunk & 1], ctx->stream));
    }
    if (bad.load()) {
        cudaStreamSynchronize(ctx->stream);
        set_error("flashv_model_create: A/B/Pi entries must be finite and inside [0,1]");
        return FLASHV_ERR_DOMAIN;
    }

    if (prep_trace()) fprintf(stderr, "[flashv prep] log A rows [%d,%d) issued: %.2f ms\n", row_lo, row_hi, ms_since(t0));
    std::vector<double> hLB((size_t)M * K), hLPi((size_t)K);
    std::vector<float> hLBf((size_t)M * Kp, 0.0f), hLBmax((size_t)M, -INFINITY);
    for (int i = 0; i < K; ++i) {
        for (int o = 0; o < M; ++o) {
            double v = log((double)B[(size_t)i * M + o]);  // F:142 / F:167
            hLB[(size_t)o * K + i] = v;
            hLBf[(size_t)o * Kp + i] = (float)v;  // "tmp = log(...)" stored into a float, F:167
            hLBmax[o] = fmaxf(hLBmax[o], (float)v);  // the largest emission term of symbol o (k_flash_persist16's bound on max delta)
        }
        hLPi[i] = log((double)Pi[i]);  // F:142
    }
    FV_CUDA(cudaMemcpyAsync(m->LBf, hLBf.data(), hLBf.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(m->LBmax, hLBmax.data(), hLBmax.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(m->LBd, hLB.data(), hLB.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(m->LPi, hLPi.data(), hLPi.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));  // the vectors are on this frame; the staging buffers are reusable
    m->row_lo = row_lo, m->row_hi = row_hi;
    if (prep_trace()) fprintf(stderr, "[flashv prep] logs uploaded: %.2f ms\n", ms_since(t0));
    m->prep_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return FLASHV_OK;
}

// Every layout the kernels read is a re-arrangement of LAd and is built on the device.
int tables_layouts(flashv_model *m)
{
    const int K = m->K, Kp = m->Kp;
    flashv_ctx *ctx = m->ctx;
    auto t0 = std::chrono::steady_clock::now();
    FV_CUDA(cudaMalloc(&m->hiT, (size_t)K * Kp * sizeof(float)));
    FV_CUDA(cudaMalloc(&m->hiC, (size_t)K * Kp * sizeof(float)));
    m->bytes += (size_t)2 * K * Kp * 4;
    if (Kp <= GROUP_MAX_KP) {
        FV_CUDA(cudaMalloc(&m->hiS, (size_t)Kp * Kp * sizeof(float)));
        m->bytes += (size_t)Kp * Kp * sizeof(float);
    }
    dim3 grid((K + 31) / 32, (Kp + 31) / 32), block(32, 8);
    k_transpose_to_f32<<<grid, block, 0, ctx->stream>>>(m->LAd, m->hiT, K, Kp);
    FV_CUDA(cudaGetLastError());
    if (m->hiS) {
        k_build_source_major<<<dim3((Kp + 255) / 256, Kp), 256, 0, ctx->stream>>>(m->LAd, m->hiS, K, Kp);
        FV_CUDA(cudaGetLastError());
    }
    m->tile_G = ctx->sm_count < K ? ctx->sm_count : K;
    build_tiled_slice(m->LAd, m->hiC, K, Kp, 0, K, m->tile_G, ctx->stream);
    FV_CUDA(cudaGetLastError());
    if (Kp > 4096) {
        // wider models: long chains, if the table (as large as LAd) is affordable; FLASHV_LACL_MAX_GB moves the limit
        const int clp = ((Kp >> 7) + 31) / 32 * 32;
        const size_t bytes = (size_t)K * 128 * clp * sizeof(double);
        const char *e = getenv("FLASHV_LACL_MAX_GB");
        const size_t max_gb = e ? (size_t)atol(e) : 32;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (bytes <= max_gb << 30 && bytes + ((size_t)8 << 30) <= free_b && 128 * clp <= 65535) {
            FV_CUDA(cudaMalloc(&m->LAcL, bytes));
            m->bytes += bytes, m->clp = clp;
            k_build_long_chains<<<dim3((K + 255) / 256, 128 * clp), 256, 0, ctx->stream>>>(m->LAd, m->LAcL, K, clp);
            FV_CUDA(cudaGetLastError());
        }
    }
    if (Kp <= 4096) {
        FV_CUDA(cudaMalloc(&m->LAc, (size_t)K * 4096 * sizeof(double)));
        m->bytes += (size_t)K * 4096 * sizeof(double);
        k_build_chains<<<dim3((K + 255) / 256, 4096), 256, 0, ctx->stream>>>(m->LAd, m->LAc, K);
        FV_CUDA(cudaGetLastError());
        m->Kp16 = (K + 255) / 256 * 256;
        FV_CUDA(cudaMalloc(&m->LAc16, (size_t)K * 4096 * sizeof(double)));
        FV_CUDA(cudaMalloc(&m->hi16, (size_t)K * m->Kp16 * sizeof(__half)));
        m->bytes += (size_t)K * 4096 * sizeof(double) + (size_t)K * m->Kp16 * sizeof(__half);
        k_build_tiled16<<<dim3((K + 255) / 256, 4096), 256, 0, ctx->stream>>>(m->LAd, m->hi16, m->LAc16, K, m->Kp16, m->tile_G);
        FV_CUDA(cudaGetLastError());
        // the largest log A of the model: with the largest emission term it bounds a step's best delta from the previous one
        double *d_res = nullptr;
        void *tmp_store = nullptr;
        size_t tmp_bytes = 0;
        FV_CUDA(cub::DeviceReduce::Max(tmp_store, tmp_bytes, m->LAd, d_res, (size_t)K * K, ctx->stream));  // size query
        FV_CUDA(cudaMalloc(&tmp_store, tmp_bytes));
        FV_CUDA(cudaMalloc(&d_res, sizeof(double)));
        FV_CUDA(cub::DeviceReduce::Max(tmp_store, tmp_bytes, m->LAd, d_res, (size_t)K * K, ctx->stream));
        FV_CUDA(cudaMemcpyAsync(&m->lamax, d_res, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(tmp_store), cudaFree(d_res);
    }
    if (prep_trace()) {
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[flashv prep] dense layouts: %.2f ms\n", ms_since(t0));
    }
    const int rc = sparse_build(m);
    if (rc != FLASHV_OK) return rc;
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    if (prep_trace()) fprintf(stderr, "[flashv prep] + edge lists: %.2f ms\n", ms_since(t0));
    m->ready = true;
    m->prep_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return FLASHV_OK;
}

}  // namespace flashv

// Program-shell ingest, F:56-95: every probability goes through fscanf("%f") into a float, i.e.
// strtof — one rounding from the decimal text.  Read the file whole and walk it with strtof.
// Large float files (A is K*K values, 300 MB of text at K=3965) are cut at whitespace into one piece
// per host thread: a counting pass finds how many tokens each piece holds, then the pieces are
// converted in parallel into their slots — the same strtof on the same characters, so the values
// are bit-identical to the serial walk.
static bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

static long parse_floats_parallel(char *buf, size_t size, long n, float *out)
{
    unsigned hw = std::thread::hardware_concurrency();
    int nthr = (int)(hw ? hw : 4);
    if (nthr > 64) nthr = 64;
    std::vector<size_t> cut((size_t)nthr + 1, size);
    cut[0] = 0;
    for (int t = 1; t < nthr; ++t) {
        size_t p = size / nthr * t;
        while (p < size && !is_space(buf[p])) ++p;  // never split a token
        cut[t] = p;
    }
    std::vector<long> count((size_t)nthr, 0);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthr; ++t)
            pool.emplace_back([&, t]() {
                long c = 0;
                bool in = false;
                for (size_t p = cut[t]; p < cut[t + 1]; ++p) {
                    const bool sp = is_space(buf[p]);
                    if (!sp && !in) ++c;
                    in = !sp;
                }
                count[t] = c;
            });
        for (auto &th : pool) th.join();
    }
    std::vector<long> first((size_t)nthr + 1, 0);
    for (int t = 0; t < nthr; ++t) first[t + 1] = first[t] + count[t];
    std::vector<long> done((size_t)nthr, 0);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthr; ++t)
            pool.emplace_back([&, t]() {
                char *p = buf + cut[t], *end = nullptr;
                char *stop = buf + cut[t + 1];
                const char saved = *stop;  // pieces end on whitespace or at the terminating NUL
                (void)saved;
                long c = 0;
                while (p < stop && first[t] + c < n) {
                    const float v = strtof(p, &end);
                    if (end == p || end > stop) break;  // not a number: the serial walk would stop here too
                    out[first[t] + c] = v;
                    p = end;
                    ++c;
                    while (p < stop && is_space(*p)) ++p;
                }
                done[t] = c;
            });
        for (auto &th : pool) th.join();
    }
    long total = 0;
    for (int t = 0; t < nthr; ++t) {
        total += done[t];
        if (done[t] != count[t] && first[t] + done[t] < n) break;  // a malformed token: everything after it is unread
    }
    return total < n ? total : n;
}

static long read_text(const char *path, long n, float *fout, int32_t *iout)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        flashv::set_error("cannot open %s", path);
        return FLASHV_ERR_ARG;
    }
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)size + 1);
    if (!buf) {
        fclose(fp);
        flashv::set_error("out of memory reading %s", path);
        return FLASHV_ERR_NOMEM;
    }
    size_t got = fread(buf, 1, (size_t)size, fp);
    fclose(fp);
    buf[got] = 0;
    long cnt = 0;
    if (fout && n >= (1 << 20)) {
        cnt = parse_floats_parallel(buf, got, n, fout);
        free(buf);
        return cnt;
    }
    char *p = buf, *end = nullptr;
    while (cnt < n) {
        if (fout) {
            float v = strtof(p, &end);
            if (end == p) break;
            fout[cnt] = v;
        } else {
            long v = strtol(p, &end, 10);
            if (end == p) break;
            iout[cnt] = (int32_t)v;
        }
        p = end;
        ++cnt;
    }
    free(buf);
    return cnt;
}

extern "C" long flashv_read_floats_text(const char *path, long n, float *out) { return read_text(path, n, out, nullptr); }
extern "C" long flashv_read_ints_text(const char *path, long n, int32_t *out) { return read_text(path, n, nullptr, out); }

// Binary side-car of a text table: the text stays canonical, <path>.f32cache holds the
// float32 values one parse of it gave, valid as long as the text file's size and mtime are the ones
// recorded in the header.  Anything wrong with the cache (missing, stale, short, unwritable
// directory) silently falls back to parsing the text.
struct CacheHeader {
    char magic[8];  // "FLASHVC1"
    long long n, text_size, text_mtime_ns;
};

extern "C" long flashv_read_floats_cached(const char *path, long n, float *out)
{
    struct stat st;
    if (stat(path, &st) != 0) {
        flashv::set_error("cannot open %s", path);
        return FLASHV_ERR_ARG;
    }
    const long long mtime_ns = (long long)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    const std::string cpath = std::string(path) + ".f32cache";
    if (FILE *fp = fopen(cpath.c_str(), "rb")) {
        CacheHeader h;
        bool ok = fread(&h, sizeof(h), 1, fp) == 1 && memcmp(h.magic, "FLASHVC1", 8) == 0 && h.n == n &&
                  h.text_size == (long long)st.st_size && h.text_mtime_ns == mtime_ns;
        ok = ok && fread(out, sizeof(float), (size_t)n, fp) == (size_t)n;
        fclose(fp);
        if (ok) return n;
    }
    const long got = read_text(path, n, out, nullptr);
    if (got == n) {
        const std::string tmp = cpath + ".tmp" + std::to_string((long long)getpid());
        if (FILE *fp = fopen(tmp.c_str(), "wb")) {
            CacheHeader h;
            memcpy(h.magic, "FLASHVC1", 8);
            h.n = n, h.text_size = (long long)st.st_size, h.text_mtime_ns = mtime_ns;
            const bool ok = fwrite(&h, sizeof(h), 1, fp) == 1 && fwrite(out, sizeof(float), (size_t)n, fp) == (size_t)n;
            if (fclose(fp) != 0 || !ok || rename(tmp.c_str(), cpath.c_str()) != 0) remove(tmp.c_str());
        }
    }
    return got;
}
