// api.cu — the extern "C" surface of libflashv.so (include/flashv.h): context, model, plan and the
// decode entry points that stand in for the reference's calc() (F:338-368, S:548-577).
//   F: = /root/reference/src/FLASH_Viterbi_multithread.c
//   S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "flashv_internal.h"

namespace flashv {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return FLASHV_ERR_CUDA;
}

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace flashv

using namespace flashv;

extern "C" const char *flashv_last_error(void) { return g_err; }
extern "C" const char *flashv_version(void) { return "flashv-b200 0.1 (sm_100a)"; }

// ---- context -------------------------------------------------------------------------------
extern "C" int flashv_ctx_create(int device, void *stream, flashv_ctx **out)
{
    if (!out) {
        set_error("flashv_ctx_create: out is NULL");
        return FLASHV_ERR_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("flashv_ctx_create: no CUDA device (%s) - this library has no CPU path",
                  e == cudaSuccess ? "count 0" : cudaGetErrorString(e));
        return FLASHV_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("flashv_ctx_create: device %d of %d", device, ndev);
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(device));
    flashv_ctx *c = new flashv_ctx();
    c->device = device;
    cudaDeviceProp prop;
    FV_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    c->coop = prop.cooperativeLaunch;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        FV_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    for (auto &ev : c->ev) FV_CUDA(cudaEventCreate(&ev));
    *out = c;
    return FLASHV_OK;
}

extern "C" void flashv_ctx_destroy(flashv_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void *flashv_ctx_stream(const flashv_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int flashv_ctx_sm_count(const flashv_ctx *c) { return c ? c->sm_count : 0; }
extern "C" int flashv_ctx_sync(flashv_ctx *c)
{
    if (!c) return FLASHV_ERR_ARG;
    FV_CUDA(cudaStreamSynchronize(c->stream));
    return FLASHV_OK;
}

static int stage_reserve(flashv_ctx *c, size_t bytes)
{
    if (c->h_stage_bytes >= bytes) return FLASHV_OK;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr, c->h_stage_bytes = 0;
    FV_CUDA(cudaMallocHost(&c->h_stage, bytes));
    c->h_stage_bytes = bytes;
    return FLASHV_OK;
}

// ---- model ---------------------------------------------------------------------------------
extern "C" int flashv_model_create(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi,
                                   flashv_model **out)
{
    if (!ctx || !A || !B || !Pi || !out || K < 1 || M < 1) {
        set_error("flashv_model_create: bad argument");
        return FLASHV_ERR_ARG;
    }
    *out = nullptr;
    FV_CUDA(cudaSetDevice(ctx->device));
    flashv_model *m = new flashv_model();
    m->ctx = ctx, m->K = K, m->M = M;
    m->Kp = (K + 127) / 128 * 128;
    int rc = tables_build(m, A, B, Pi);
    if (rc != FLASHV_OK) {
        flashv_model_destroy(m);
        return rc;
    }
    *out = m;
    return FLASHV_OK;
}

extern "C" void flashv_model_destroy(flashv_model *m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (flashv_plan *p : m->plan_cache) flashv_plan_destroy(p);
    cudaFree(m->hiT), cudaFree(m->hiC), cudaFree(m->hiS), cudaFree(m->csc_ptr), cudaFree(m->csc_k), cudaFree(m->csc_la), cudaFree(m->csr_cut), cudaFree(m->csr_i), cudaFree(m->csr_la), cudaFree(m->LAd), cudaFree(m->LBf), cudaFree(m->LBd), cudaFree(m->LPi);
    cudaFree(m->scratch_f), cudaFree(m->scratch_i), cudaFree(m->scratch_x);
    delete m;
}

extern "C" int flashv_model_K(const flashv_model *m) { return m ? m->K : 0; }
extern "C" int flashv_model_M(const flashv_model *m) { return m ? m->M : 0; }
extern "C" double flashv_model_prep_ms(const flashv_model *m) { return m ? m->prep_ms : 0; }

// ---- plan ----------------------------------------------------------------------------------
static void add_pass(flashv_plan *p, std::vector<VecDesc> &all, const std::vector<Task> &tasks, int flags)
{
    Pass ps;
    ps.vec_offset = all.size();
    ps.full_range = (flags & VEC_FULL_RANGE) != 0;
    int row = 0;
    for (const Task &t : tasks)  // tasks arrive longest first; batch innermost keeps the order
        for (int b = 0; b < p->batch; ++b) {
            VecDesc vd{b, t.L, t.R, t.mid, row, flags};
            row += t.R - t.mid;
            if (all.size() == ps.vec_offset) ps.first_vec = vd;
            all.push_back(vd);
            ps.max_steps = std::max(ps.max_steps, t.R - t.L);
        }
    ps.nvec = (int)tasks.size() * p->batch;
    ps.psi_rows = row;
    ps.nactive.assign(ps.max_steps + 2, 0);
    for (int s = 1; s <= ps.max_steps; ++s) {
        int n = 0;
        for (const Task &t : tasks)
            if (t.R - t.L >= s) n += p->batch;
        ps.nactive[s] = n;
    }
    p->passes.push_back(std::move(ps));
}

extern "C" int flashv_plan_create(flashv_model *m, int T, int N, int batch, int B, int engine, flashv_plan **out)
{
    if (!m || !out || batch < 1 || B < 0) {
        set_error("flashv_plan_create: bad argument");
        return FLASHV_ERR_ARG;
    }
    *out = nullptr;
    if (B > m->K) {
        set_error("flashv_plan_create: BeamSearchWidth %d > K %d (the reference scans stale slots there)", B, m->K);
        return FLASHV_ERR_ARG;
    }
    flashv_ctx *ctx = m->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    flashv_plan *p = new flashv_plan();
    p->model = m, p->T = T, p->N = N, p->batch = batch, p->B = B;
    if (!build_schedule(T, N, &p->sched)) {
        delete p;
        set_error("flashv_plan_create: unsupported (T=%d, N=%d): need T >= 2, N >= 1 and not T == 2N with N > 2", T, N);
        return FLASHV_ERR_ARG;
    }
    if (engine == FLASHV_ENGINE_AUTO) engine = ctx->coop ? FLASHV_ENGINE_PERSISTENT : FLASHV_ENGINE_STEP;
    if (engine == FLASHV_ENGINE_SPARSE && (!sparse_engine_available(m) || !ctx->coop || B > 0)) {
        delete p;
        set_error("flashv_plan_create: the sparse engine needs a FLASH plan, a cooperative-launch device and a model whose "
                  "transition table is at most half non-zero with K < 65536");
        return FLASHV_ERR_ARG;
    }
    p->engine = engine;
    // FLASH-BS keeps one flag bit in every backpointer entry (bs_kernels.cu)
    p->psi16 = (B > 0 ? m->K < 32768 : m->K < 65535) ? 1 : 0;

    std::vector<VecDesc> all;
    if (p->sched.first_pass) {
        // nvviterNdivide over (0,T-1): backpointers are needed from the first boundary on
        std::vector<Task> fp{{0, T - 1, p->sched.mids[0]}};
        add_pass(p, all, fp, VEC_FULL_RANGE | VEC_FIRST_PASS);
    }
    for (const auto &lvl : p->sched.levels) {
        const bool root = !p->sched.first_pass && &lvl == &p->sched.levels[0];
        add_pass(p, all, lvl, root ? VEC_FULL_RANGE : 0);
    }
    int max_rows = 1;
    for (const Pass &ps : p->passes) {
        p->max_vec = std::max(p->max_vec, ps.nvec);
        max_rows = std::max(max_rows, ps.psi_rows);
    }
    const int K = m->K, Kp = m->Kp;
    const size_t psi_bytes = (size_t)max_rows * K * (p->psi16 ? 2 : 4) + 64;  // + slack: the staged backtrack reads whole 16-byte words
    const size_t delta_bytes = (size_t)(2 * p->max_vec + 4) * Kp * sizeof(float);  // two delta sets + the persistent engine's {value,step} exchange buffers
    std::vector<uint8_t> ismid((size_t)T, 0);
    for (int mid : p->sched.mids) ismid[mid] = 1;

    int rc = FLASHV_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        rc = cuda_fail(e, what, __FILE__, __LINE__);
    };
    cudaError_t e;
#define PL_ALLOC(ptr, nb_)                                               \
    if (rc == FLASHV_OK && (e = cudaMalloc(&(ptr), (nb_))) != cudaSuccess) fail(e, "cudaMalloc " #ptr); \
    else if (rc == FLASHV_OK) p->bytes += (nb_);
    PL_ALLOC(p->d_ob, (size_t)batch * T * 4);
    PL_ALLOC(p->d_ans, (size_t)batch * T * 4);
    PL_ALLOC(p->d_score, (size_t)batch * 4);
    PL_ALLOC(p->d_psi, psi_bytes);
    PL_ALLOC(p->d_vecs, all.size() * sizeof(VecDesc));
    PL_ALLOC(p->d_ismid, (size_t)T);
    PL_ALLOC(p->d_endstate, (size_t)p->max_vec * 4);
    PL_ALLOC(p->d_sync, 256);
    if (B == 0) {
        PL_ALLOC(p->d_delta, delta_bytes);
    } else {
        PL_ALLOC(p->d_bs_score, (size_t)max_rows * K * sizeof(float));  // score vectors for backtrack-time repairs
    }
#undef PL_ALLOC
    if (rc == FLASHV_OK && p->d_delta && (e = cudaMemsetAsync(p->d_delta, 0, delta_bytes, ctx->stream)) != cudaSuccess)
        fail(e, "memset delta");  // padding lanes k >= K must stay finite (hiT pads with -inf)
    if (rc == FLASHV_OK && (e = cudaMemsetAsync(p->d_sync, 0, 256, ctx->stream)) != cudaSuccess) fail(e, "memset sync");
    if (rc == FLASHV_OK && (e = cudaMemsetAsync(p->d_ans, 0xff, (size_t)batch * T * 4, ctx->stream)) != cudaSuccess)
        fail(e, "memset ans");
    if (rc == FLASHV_OK && (e = cudaMemcpyAsync(p->d_vecs, all.data(), all.size() * sizeof(VecDesc),
                                                 cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        fail(e, "upload vecs");
    if (rc == FLASHV_OK &&
        (e = cudaMemcpyAsync(p->d_ismid, ismid.data(), (size_t)T, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
        fail(e, "upload ismid");
    if (rc == FLASHV_OK && (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) fail(e, "sync");
    if (rc != FLASHV_OK) {
        flashv_plan_destroy(p);
        return rc;
    }
    *out = p;
    return FLASHV_OK;
}

extern "C" void flashv_plan_destroy(flashv_plan *p)
{
    if (!p) return;
    cudaSetDevice(p->model->ctx->device);
    cudaStreamSynchronize(p->model->ctx->stream);
    for (int r = 0; r < 8; ++r)
        if (p->peer_ipc[r]) cudaIpcCloseMemHandle(p->peer_delta[r]), cudaIpcCloseMemHandle(p->peer_psi[r]);
    cudaFree(p->hiC_shard);
    cudaFree(p->d_ob), cudaFree(p->d_ans), cudaFree(p->d_score), cudaFree(p->d_delta), cudaFree(p->d_psi);
    cudaFree(p->d_vecs), cudaFree(p->d_ismid), cudaFree(p->d_endstate), cudaFree(p->d_sync), cudaFree(p->d_bs_score);
    delete p;
}

// ---- state sharding across GPUs ------------------------------------------------------
extern "C" int flashv_plan_shard_init(flashv_plan *p, int rank, int world)
{
    if (!p || world < 1 || world > 8 || rank < 0 || rank >= world) {
        set_error("flashv_plan_shard_init: rank %d of %d (at most 8 GPUs)", rank, world);
        return FLASHV_ERR_ARG;
    }
    if (p->B != 0 || p->batch != 1 || p->engine != FLASHV_ENGINE_PERSISTENT) {
        set_error("flashv_plan_shard_init: only single-sequence FLASH plans on the persistent engine shard their states");
        return FLASHV_ERR_ARG;
    }
    if (p->hiC_shard) {
        set_error("flashv_plan_shard_init: plan is already sharded");
        return FLASHV_ERR_STATE;
    }
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    p->shard_rank = rank, p->shard_world = world;
    if (world == 1) return FLASHV_OK;
    int rc = shard_build_table(p);
    if (rc != FLASHV_OK) return rc;
    p->peer_delta[rank] = p->d_delta, p->peer_psi[rank] = p->d_psi;
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_buffers(flashv_plan *p, void **delta_base, void **psi_base)
{
    if (!p || !delta_base || !psi_base) return FLASHV_ERR_ARG;
    *delta_base = p->d_delta, *psi_base = p->d_psi;
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_ipc_handles(flashv_plan *p, void *out128)
{
    if (!p || !out128) return FLASHV_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two handles fill 128 bytes");
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    cudaIpcMemHandle_t h[2];
    FV_CUDA(cudaIpcGetMemHandle(&h[0], p->d_delta));
    FV_CUDA(cudaIpcGetMemHandle(&h[1], p->d_psi));
    memcpy(out128, h, sizeof(h));
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_set_peer(flashv_plan *p, int peer_rank, int peer_device, void *delta_base, void *psi_base)
{
    if (!p || peer_rank < 0 || peer_rank >= p->shard_world || !delta_base || !psi_base) {
        set_error("flashv_plan_shard_set_peer: bad argument");
        return FLASHV_ERR_ARG;
    }
    flashv_ctx *ctx = p->model->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    if (peer_device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        cudaGetLastError();
    }
    p->peer_delta[peer_rank] = (float *)delta_base, p->peer_psi[peer_rank] = psi_base;
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_open_peer(flashv_plan *p, int peer_rank, const void *handles128)
{
    if (!p || peer_rank < 0 || peer_rank >= p->shard_world || !handles128 || peer_rank == p->shard_rank) {
        set_error("flashv_plan_shard_open_peer: bad argument");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    cudaIpcMemHandle_t h[2];
    memcpy(h, handles128, sizeof(h));
    void *d = nullptr, *s = nullptr;
    FV_CUDA(cudaIpcOpenMemHandle(&d, h[0], cudaIpcMemLazyEnablePeerAccess));
    FV_CUDA(cudaIpcOpenMemHandle(&s, h[1], cudaIpcMemLazyEnablePeerAccess));
    p->peer_delta[peer_rank] = (float *)d, p->peer_psi[peer_rank] = s, p->peer_ipc[peer_rank] = true;
    return FLASHV_OK;
}

extern "C" int flashv_plan_upload(flashv_plan *p, const int32_t *ob)
{
    if (!p || !ob) {
        set_error("flashv_plan_upload: bad argument");
        return FLASHV_ERR_ARG;
    }
    const int M = p->model->M;
    const size_t n = (size_t)p->batch * p->T;
    for (size_t i = 0; i < n; ++i)
        if (ob[i] < 0 || ob[i] >= M) {
            set_error("flashv_plan_upload: observation %zu = %d outside [0,%d)", i, ob[i], M);
            return FLASHV_ERR_ARG;
        }
    flashv_ctx *ctx = p->model->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    FV_CUDA(cudaMemcpyAsync(p->d_ob, ob, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    p->uploaded = true;
    return FLASHV_OK;
}

extern "C" int flashv_plan_run(flashv_plan *p)
{
    if (!p) return FLASHV_ERR_ARG;
    if (!p->uploaded) {
        set_error("flashv_plan_run: no observations uploaded");
        return FLASHV_ERR_STATE;
    }
    flashv_ctx *ctx = p->model->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    p->launches = 0;
    FV_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    bool first = true;
    for (const Pass &ps : p->passes) {
        int rc = p->B > 0 ? bs_run_pass(p, ps) : flash_run_pass(p, ps, first);
        if (rc != FLASHV_OK) return rc;
        first = false;
    }
    FV_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    p->ran = true;
    return FLASHV_OK;
}

extern "C" int flashv_plan_download(flashv_plan *p, int32_t *path_out, float *score_out)
{
    if (!p || !path_out) {
        set_error("flashv_plan_download: bad argument");
        return FLASHV_ERR_ARG;
    }
    if (!p->ran) {
        set_error("flashv_plan_download: plan has not run");
        return FLASHV_ERR_STATE;
    }
    flashv_ctx *ctx = p->model->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    FV_CUDA(cudaMemcpyAsync(path_out, p->d_ans, (size_t)p->batch * p->T * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (score_out)
        FV_CUDA(cudaMemcpyAsync(score_out, p->d_score, (size_t)p->batch * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    return FLASHV_OK;
}

extern "C" int flashv_plan_report(flashv_plan *p, flashv_report *r)
{
    if (!p || !r) return FLASHV_ERR_ARG;
    flashv_ctx *ctx = p->model->ctx;
    flashv_report rep = p->rep;
    if (p->ran) {
        FV_CUDA(cudaSetDevice(ctx->device));
        FV_CUDA(cudaEventSynchronize(ctx->ev[1]));
        float ms = 0;
        FV_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        rep.decode_ms = ms;
        if (p->B == 0 && !p->passes.empty() && p->passes[0].full_range) {
            FV_CUDA(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
            rep.first_pass_ms = ms;
        }
    }
    rep.executed_steps = p->sched.executed_steps;
    rep.device_bytes = (long long)(p->bytes + p->model->bytes);
    rep.memory_bytes = p->B > 0 ? flashv_bs_memory_bytes(p->T, p->N, p->B) : flashv_memory_bytes(p->model->K, p->T, p->N);
    rep.first_pass = p->sched.first_pass ? 1 : 0;
    rep.n_tasks = (int)p->sched.fifo.size();
    rep.n_levels = (int)p->sched.levels.size();
    rep.kernel_launches = p->launches;
    rep.engine = p->engine;
    *r = rep;
    return FLASHV_OK;
}

// ---- one-call decodes ------------------------------------------------------------------------
static int cached_plan(flashv_model *m, int T, int N, int batch, int B, flashv_plan **out)
{
    for (flashv_plan *p : m->plan_cache)
        if (p->T == T && p->N == N && p->batch == batch && p->B == B) {
            *out = p;
            return FLASHV_OK;
        }
    flashv_plan *p = nullptr;
    int rc = flashv_plan_create(m, T, N, batch, B, FLASHV_ENGINE_AUTO, &p);
    if (rc != FLASHV_OK) return rc;
    if (m->plan_cache.size() >= 8) {  // small LRU-less cache: drop the oldest
        flashv_plan_destroy(m->plan_cache.front());
        m->plan_cache.erase(m->plan_cache.begin());
    }
    m->plan_cache.push_back(p);
    *out = p;
    return FLASHV_OK;
}

static int decode_common(flashv_model *m, const int32_t *ob, int batch, int T, int N, int B, int32_t *path_out,
                         float *score_out, flashv_report *report)
{
    if (!m || !ob || !path_out) {
        set_error("flashv decode: bad argument");
        return FLASHV_ERR_ARG;
    }
    flashv_plan *p = nullptr;
    int rc = cached_plan(m, T, N, batch, B, &p);
    if (rc != FLASHV_OK) return rc;
    flashv_ctx *ctx = m->ctx;
    // stage through pinned memory so both copies are true async DMA on the context's stream
    const size_t n = (size_t)batch * T;
    rc = stage_reserve(ctx, (2 * n + batch) * 4);
    if (rc != FLASHV_OK) return rc;
    double t0 = now_ms();
    memcpy(ctx->h_stage, ob, n * 4);
    rc = flashv_plan_upload(p, ctx->h_stage);
    if (rc != FLASHV_OK) return rc;
    double t1 = now_ms();
    rc = flashv_plan_run(p);
    if (rc != FLASHV_OK) return rc;
    int32_t *h_path = ctx->h_stage + n;
    float *h_score = reinterpret_cast<float *>(ctx->h_stage + 2 * n);
    rc = flashv_plan_download(p, h_path, h_score);
    if (rc != FLASHV_OK) return rc;
    memcpy(path_out, h_path, n * 4);
    if (score_out) memcpy(score_out, h_score, (size_t)batch * 4);
    double t2 = now_ms();
    if (report) {
        rc = flashv_plan_report(p, report);
        if (rc != FLASHV_OK) return rc;
        report->h2d_ms = t1 - t0;
        report->d2h_ms = (t2 - t1) - report->decode_ms;
    }
    return FLASHV_OK;
}

extern "C" int flashv_decode(flashv_model *m, const int32_t *ob, int T, int N, int32_t *path_out, float *score_out,
                             flashv_report *report)
{
    return decode_common(m, ob, 1, T, N, 0, path_out, score_out, report);
}

extern "C" int flashv_bs_decode(flashv_model *m, const int32_t *ob, int T, int N, int B, int32_t *path_out,
                                float *score_out, flashv_report *report)
{
    if (B < 1) {
        set_error("flashv_bs_decode: BeamSearchWidth must be >= 1");
        return FLASHV_ERR_ARG;
    }
    return decode_common(m, ob, 1, T, N, B, path_out, score_out, report);
}

extern "C" int flashv_decode_batch(flashv_model *m, const int32_t *ob, int batch, int T, int N, int32_t *path_out,
                                   float *score_out, flashv_report *report)
{
    return decode_common(m, ob, batch, T, N, 0, path_out, score_out, report);
}

extern "C" int flashv_bs_decode_batch(flashv_model *m, const int32_t *ob, int batch, int T, int N, int B,
                                      int32_t *path_out, float *score_out, flashv_report *report)
{
    if (B < 1) {
        set_error("flashv_bs_decode_batch: BeamSearchWidth must be >= 1");
        return FLASHV_ERR_ARG;
    }
    return decode_common(m, ob, batch, T, N, B, path_out, score_out, report);
}

// ---- per-step test hooks -----------------------------------------------------------------------
extern "C" int flashv_trellis_init(flashv_model *m, int prev_state, int ob0, float *delta_out)
{
    if (!m || !delta_out || prev_state >= m->K || ob0 < 0 || ob0 >= m->M) {
        set_error("flashv_trellis_init: bad argument");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(m->ctx->device));
    int rc = flash_single_init(m, prev_state, ob0, m->scratch_f);
    if (rc != FLASHV_OK) return rc;
    FV_CUDA(cudaMemcpyAsync(delta_out, m->scratch_f, (size_t)m->K * 4, cudaMemcpyDeviceToHost, m->ctx->stream));
    FV_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return FLASHV_OK;
}

extern "C" int flashv_trellis_step(flashv_model *m, const float *delta_in, int o, float *delta_out, int32_t *psi_out,
                                   int engine)
{
    if (!m || !delta_in || !delta_out || !psi_out || o < 0 || o >= m->M) {
        set_error("flashv_trellis_step: bad argument");
        return FLASHV_ERR_ARG;
    }
    flashv_ctx *ctx = m->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    if (engine == FLASHV_ENGINE_AUTO) engine = FLASHV_ENGINE_STEP;
    float *din = m->scratch_f, *dout = m->scratch_f + m->Kp;
    int32_t *psi = m->scratch_i + 64;  // past the descriptor words of the hooks
    FV_CUDA(cudaMemsetAsync(din, 0, (size_t)m->Kp * 4, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(din, delta_in, (size_t)m->K * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = flash_single_step(m, din, o, dout, psi, engine);
    if (rc != FLASHV_OK) return rc;
    FV_CUDA(cudaMemcpyAsync(delta_out, dout, (size_t)m->K * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(psi_out, psi, (size_t)m->K * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    return FLASHV_OK;
}

extern "C" int flashv_bs_score_step(flashv_model *m, const float *heap_val, const int32_t *heap_state, int B, int o,
                                    float *score_out, int32_t *arg_slot_out)
{
    if (!m || !heap_val || !heap_state || !score_out || !arg_slot_out || B < 1 || B > m->Kp || o < 0 || o >= m->M) {
        set_error("flashv_bs_score_step: bad argument");
        return FLASHV_ERR_ARG;
    }
    for (int c = 0; c < B; ++c)
        if (heap_state[c] < 0 || heap_state[c] >= m->K) {
            set_error("flashv_bs_score_step: heap state %d outside [0,K)", heap_state[c]);
            return FLASHV_ERR_ARG;
        }
    flashv_ctx *ctx = m->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    float *hv = m->scratch_f, *score = m->scratch_f + m->Kp;
    int32_t *hs = m->scratch_i, *arg = m->scratch_i + m->Kp;
    FV_CUDA(cudaMemcpyAsync(hv, heap_val, (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(hs, heap_state, (size_t)B * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = bs_single_score(m, hv, hs, B, o, score, arg);
    if (rc != FLASHV_OK) return rc;
    FV_CUDA(cudaMemcpyAsync(score_out, score, (size_t)m->K * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FV_CUDA(cudaMemcpyAsync(arg_slot_out, arg, (size_t)m->K * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FV_CUDA(cudaStreamSynchronize(ctx->stream));
    return FLASHV_OK;
}

extern "C" int flashv_bs_heap_replay(flashv_ctx *ctx, const float *score, int K, int B, float *heap_val_out,
                                     int32_t *heap_state_out)
{
    if (!ctx || !score || !heap_val_out || !heap_state_out || B < 1 || B > K) {
        set_error("flashv_bs_heap_replay: bad argument");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(ctx->device));
    float *d_score = nullptr, *d_hv = nullptr;
    int32_t *d_hs = nullptr;
    FV_CUDA(cudaMalloc(&d_score, (size_t)K * 4));
    FV_CUDA(cudaMalloc(&d_hv, (size_t)B * 4));
    FV_CUDA(cudaMalloc(&d_hs, (size_t)B * 4));
    FV_CUDA(cudaMemcpyAsync(d_score, score, (size_t)K * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = bs_single_replay(ctx, d_score, K, B, d_hv, d_hs);
    if (rc == FLASHV_OK) {
        FV_CUDA(cudaMemcpyAsync(heap_val_out, d_hv, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaMemcpyAsync(heap_state_out, d_hs, (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
        FV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(d_score), cudaFree(d_hv), cudaFree(d_hs);
    return rc;
}
