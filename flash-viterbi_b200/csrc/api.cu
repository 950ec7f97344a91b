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
    for (auto &ev : c->ev_prep) FV_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
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
    for (auto &ev : c->ev_prep)
        if (ev) cudaEventDestroy(ev);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto &b : c->h_prep)
        if (b) cudaFreeHost(b);
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
static int model_new(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi, flashv_model **out,
                     const char *who)
{
    if (!ctx || !A || !B || !Pi || !out || K < 1 || M < 1) {
        set_error("%s: bad argument", who);
        return FLASHV_ERR_ARG;
    }
    *out = nullptr;
    FV_CUDA(cudaSetDevice(ctx->device));
    flashv_model *m = new flashv_model();
    m->ctx = ctx, m->K = K, m->M = M;
    m->Kp = (K + 127) / 128 * 128;
    int rc = tables_alloc(m);
    if (rc != FLASHV_OK) {
        flashv_model_destroy(m);
        return rc;
    }
    *out = m;
    return FLASHV_OK;
}

extern "C" int flashv_model_create(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi,
                                   flashv_model **out)
{
    flashv_model *m = nullptr;
    int rc = model_new(ctx, K, M, A, B, Pi, &m, "flashv_model_create");
    if (rc != FLASHV_OK) return rc;
    rc = tables_logs(m, A, B, Pi, 0, K, 1);
    if (rc == FLASHV_OK) rc = tables_layouts(m);
    if (rc != FLASHV_OK) {
        flashv_model_destroy(m);
        *out = nullptr;
        return rc;
    }
    *out = m;
    return FLASHV_OK;
}

// A model created in parts (include/flashv.h "model creation shared by the ranks of a box"): the host
// logarithms are what model creation costs (one libm call per table entry, F:170), so `world` ranks
// each compute K/world rows and fetch the rest from the peers' device tables over NVLink.
extern "C" int flashv_model_create_rows(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi,
                                        int rank, int world, flashv_model **out)
{
    if (world < 1 || rank < 0 || rank >= world) {
        set_error("flashv_model_create_rows: rank %d of %d", rank, world);
        return FLASHV_ERR_ARG;
    }
    flashv_model *m = nullptr;
    int rc = model_new(ctx, K, M, A, B, Pi, &m, "flashv_model_create_rows");
    if (rc != FLASHV_OK) return rc;
    const int lo = (int)((long long)rank * K / world), hi = (int)((long long)(rank + 1) * K / world);
    rc = tables_logs(m, A, B, Pi, lo, hi, world);
    if (rc != FLASHV_OK) {
        flashv_model_destroy(m);
        *out = nullptr;
        return rc;
    }
    m->part_rank = rank, m->part_world = world;
    *out = m;
    return FLASHV_OK;
}

extern "C" int flashv_model_rows_handle(flashv_model *m, void *out64)
{
    if (!m || !out64) {
        set_error("flashv_model_rows_handle: bad argument");
        return FLASHV_ERR_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "one handle fills 64 bytes");
    FV_CUDA(cudaSetDevice(m->ctx->device));
    cudaIpcMemHandle_t h;
    FV_CUDA(cudaIpcGetMemHandle(&h, m->LAd));
    memcpy(out64, &h, sizeof(h));
    return FLASHV_OK;
}

static int pull_rows(flashv_model *m, int peer_rank, const double *peer_LAd)
{
    const int K = m->K;
    const int lo = (int)((long long)peer_rank * K / m->part_world), hi = (int)((long long)(peer_rank + 1) * K / m->part_world);
    if (hi > lo)
        FV_CUDA(cudaMemcpyAsync(m->LAd + (size_t)lo * K, peer_LAd + (size_t)lo * K, (size_t)(hi - lo) * K * sizeof(double),
                                cudaMemcpyDefault, m->ctx->stream));
    return FLASHV_OK;
}

extern "C" int flashv_model_pull_rows(flashv_model *m, int peer_rank, const void *handle64)
{
    if (!m || !handle64 || m->ready || peer_rank < 0 || peer_rank >= m->part_world || peer_rank == m->part_rank) {
        set_error("flashv_model_pull_rows: bad argument or model already finished");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(m->ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *peer = nullptr;
    FV_CUDA(cudaIpcOpenMemHandle(&peer, h, cudaIpcMemLazyEnablePeerAccess));
    int rc = pull_rows(m, peer_rank, reinterpret_cast<const double *>(peer));
    cudaError_t e = cudaStreamSynchronize(m->ctx->stream);
    cudaIpcCloseMemHandle(peer);
    if (rc == FLASHV_OK && e != cudaSuccess) rc = cuda_fail(e, "pull rows", __FILE__, __LINE__);
    return rc;
}

extern "C" int flashv_model_pull_rows_from(flashv_model *m, int peer_rank, const flashv_model *peer)
{
    if (!m || !peer || m->ready || peer_rank < 0 || peer_rank >= m->part_world || peer->K != m->K) {
        set_error("flashv_model_pull_rows_from: bad argument or model already finished");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(m->ctx->device));
    if (peer->ctx->device != m->ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
        cudaGetLastError();
    }
    int rc = pull_rows(m, peer_rank, peer->LAd);
    if (rc == FLASHV_OK) FV_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return rc;
}

extern "C" int flashv_model_finish(flashv_model *m)
{
    if (!m || m->ready) {
        set_error("flashv_model_finish: no model, or already finished");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(m->ctx->device));
    return tables_layouts(m);
}

extern "C" void flashv_model_destroy(flashv_model *m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (flashv_plan *p : m->plan_cache) flashv_plan_destroy(p);
    cudaFree(m->hiT), cudaFree(m->hiC), cudaFree(m->LAc), cudaFree(m->LAcL), cudaFree(m->LBmax), cudaFree(m->hi16), cudaFree(m->LAc16), cudaFree(m->hiS), cudaFree(m->csc_ptr), cudaFree(m->csc_k), cudaFree(m->csc_la), cudaFree(m->csr_cut), cudaFree(m->csr_i), cudaFree(m->csr_la), cudaFree(m->LAd), cudaFree(m->LBf), cudaFree(m->LBd), cudaFree(m->LPi);
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

// Device copy of every pass's nactive[] (the persistent level kernel walks it step by step).
static int upload_nactive(flashv_plan *p)
{
    std::vector<int> all;
    for (Pass &ps : p->passes) {
        ps.nactive_off = all.size();
        all.insert(all.end(), ps.nactive.begin(), ps.nactive.end());
        all.push_back(0);
    }
    cudaFree(p->d_nactive);
    p->d_nactive = nullptr;
    FV_CUDA(cudaMalloc(&p->d_nactive, (all.size() + 1) * sizeof(int)));
    FV_CUDA(cudaMemcpy(p->d_nactive, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice));
    return FLASHV_OK;
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
    if (!m->ready) {
        set_error("flashv_plan_create: the model was created in parts and flashv_model_finish has not run");
        return FLASHV_ERR_STATE;
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
    // Engine feasibility is decided here, not at run time: the persistent kernel needs the delta vector
    // plus two ring stages in one CTA's shared memory (K up to ~42 k), the per-step kernels the delta
    // vector alone (K <= 56,320, the largest K a FLASH plan accepts; include/flashv.h).
    const bool persist_ok = persistent_engine_fits(ctx, m->Kp);
    if (B == 0 && m->Kp > STEP_MAX_KP) {
        delete p;
        set_error("flashv_plan_create: K=%d is beyond the largest supported K (%d): one delta vector must fit shared memory", m->K, STEP_MAX_KP);
        return FLASHV_ERR_ARG;
    }
    if (engine == FLASHV_ENGINE_AUTO) engine = persist_ok ? FLASHV_ENGINE_PERSISTENT : FLASHV_ENGINE_STEP;
    if (engine == FLASHV_ENGINE_PERSISTENT && B == 0 && !persist_ok) {
        delete p;
        set_error("flashv_plan_create: the persistent engine does not fit K=%d on this device (no cooperative launch, or delta + two ring "
                  "stages exceed %d bytes of shared memory); use FLASHV_ENGINE_AUTO or FLASHV_ENGINE_STEP", m->K, ctx->smem_optin);
        return FLASHV_ERR_ARG;
    }
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
    {
        // scratch of the parallel walk back (k_flash_backtrack_par): one composed map of K states per window of rows
        const size_t row_bytes = (size_t)K * (p->psi16 ? 2 : 4);
        const int win_rows = (int)((200 * 1024 - 32) / row_bytes);
        if (B == 0 && batch == 1 && win_rows >= 4) {
            p->bt_windows = (T + win_rows - 1) / win_rows + 1;
            PL_ALLOC(p->d_btmap, (size_t)p->bt_windows * ((size_t)K + 1) * 4);
        }
    }
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
    if (rc == FLASHV_OK) rc = upload_nactive(p);
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
        if (p->peer_ipc[r]) cudaIpcCloseMemHandle(p->peer_region[r]);
    cudaFree(p->hiC_shard), cudaFree(p->shard_region), cudaFree(p->d_lvl_mid);
    cudaFree(p->d_ob), cudaFree(p->d_ans), cudaFree(p->d_score), cudaFree(p->d_delta), cudaFree(p->d_psi);
    cudaFree(p->d_vecs), cudaFree(p->d_ismid), cudaFree(p->d_endstate), cudaFree(p->d_sync), cudaFree(p->d_btmap), cudaFree(p->d_bs_score);
    cudaFree(p->d_nactive);
    delete p;
}

// ---- state sharding across GPUs ------------------------------------------------------
// Rebuild the level passes of a sharded plan so that this rank runs every world-th task of a level
// (tasks are sorted longest first, so round-robin balances the steps), and record which Ans entries
// every level produces: they are exchanged after the level (shard_ans_exchange).
static int shard_spread_levels(flashv_plan *p)
{
    const int first = p->sched.first_pass ? 1 : 0;
    std::vector<VecDesc> all;
    std::vector<Pass> old;
    old.swap(p->passes);
    std::vector<int32_t> mids;
    p->lvl_off.clear(), p->lvl_cnt.clear();
    size_t li = 0;
    if (first) {
        std::vector<Task> fp{{0, p->T - 1, p->sched.mids[0]}};
        add_pass(p, all, fp, VEC_FULL_RANGE | VEC_FIRST_PASS);
    }
    for (const auto &lvl : p->sched.levels) {
        const bool root = !first && li == 0;
        p->lvl_off.push_back((int)mids.size()), p->lvl_cnt.push_back((int)lvl.size());
        for (const Task &t : lvl) mids.push_back(t.mid);
        std::vector<Task> mine;
        if (root) mine = lvl;  // pass 0 of a plan without an N-way pass: the state-sharded one, run by every rank
        else
            for (size_t t = 0; t < lvl.size(); ++t)
                if ((int)(t % (size_t)p->shard_world) == p->shard_rank) mine.push_back(lvl[t]);
        add_pass(p, all, mine, root ? VEC_FULL_RANGE : 0);  // possibly empty: the rank still takes part in the exchange
        ++li;
    }
    (void)old;
    FV_CUDA(cudaMemcpyAsync(p->d_vecs, all.data(), all.size() * sizeof(VecDesc), cudaMemcpyHostToDevice, p->model->ctx->stream));
    FV_CUDA(cudaMalloc(&p->d_lvl_mid, (mids.size() + 1) * sizeof(int32_t)));
    FV_CUDA(cudaMemcpyAsync(p->d_lvl_mid, mids.data(), mids.size() * sizeof(int32_t), cudaMemcpyHostToDevice, p->model->ctx->stream));
    FV_CUDA(cudaStreamSynchronize(p->model->ctx->stream));
    return upload_nactive(p);
}

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
    if (p->hiC_shard || p->shard_region) {
        set_error("flashv_plan_shard_init: plan is already sharded");
        return FLASHV_ERR_STATE;
    }
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    p->shard_rank = rank, p->shard_world = world;
    if (world == 1) return FLASHV_OK;
    int rc = shard_build_table(p);
    if (rc != FLASHV_OK) return rc;
    // the region peers store into: exchange words, Ans exchange words, backpointer rows of pass 0
    const int K = p->model->K, Kp = p->model->Kp;
    const size_t xch_bytes = (size_t)2 * Kp * 8;
    const size_t ans_bytes = ((size_t)p->T * 8 + 255) & ~(size_t)255;
    const size_t psi_bytes = (size_t)p->passes[0].psi_rows * K * (p->psi16 ? 2 : 4) + 64;
    p->shard_ans_off = xch_bytes, p->shard_psi_off = xch_bytes + ans_bytes;
    p->shard_region_bytes = xch_bytes + ans_bytes + psi_bytes;
    FV_CUDA(cudaMalloc(&p->shard_region, p->shard_region_bytes));
    FV_CUDA(cudaMemsetAsync(p->shard_region, 0, xch_bytes + ans_bytes, p->model->ctx->stream));
    p->bytes += p->shard_region_bytes;
    p->peer_region[rank] = p->shard_region;
    return shard_spread_levels(p);
}

extern "C" int flashv_plan_shard_buffers(flashv_plan *p, void **region_base, size_t *region_bytes)
{
    if (!p || !region_base || !p->shard_region) {
        set_error("flashv_plan_shard_buffers: plan is not sharded");
        return FLASHV_ERR_ARG;
    }
    *region_base = p->shard_region;
    if (region_bytes) *region_bytes = p->shard_region_bytes;
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_ipc_handle(flashv_plan *p, void *out64)
{
    if (!p || !out64 || !p->shard_region) {
        set_error("flashv_plan_shard_ipc_handle: plan is not sharded");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    cudaIpcMemHandle_t h;
    FV_CUDA(cudaIpcGetMemHandle(&h, p->shard_region));
    memcpy(out64, &h, sizeof(h));
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_set_peer(flashv_plan *p, int peer_rank, int peer_device, void *region_base)
{
    if (!p || !p->shard_region || peer_rank < 0 || peer_rank >= p->shard_world || !region_base) {
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
    p->peer_region[peer_rank] = (unsigned char *)region_base;
    return FLASHV_OK;
}

extern "C" int flashv_plan_shard_open_peer(flashv_plan *p, int peer_rank, const void *handle64)
{
    if (!p || !p->shard_region || peer_rank < 0 || peer_rank >= p->shard_world || !handle64 || peer_rank == p->shard_rank) {
        set_error("flashv_plan_shard_open_peer: bad argument");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(p->model->ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *d = nullptr;
    FV_CUDA(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer_region[peer_rank] = (unsigned char *)d, p->peer_ipc[peer_rank] = true;
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
    const bool sharded = p->shard_world > 1;
    if (sharded) {
        // Peers store into this plan's region during a run, and the tags that make those stores
        // self-validating count runs: every rank must have finished run n before any starts run n+1
        // (include/flashv.h).  What can be checked locally is that this rank's previous work is done.
        if (cudaStreamQuery(ctx->stream) == cudaErrorNotReady) {
            set_error("flashv_plan_run: a state-sharded plan needs an idle stream (sync, then barrier across ranks, then run)");
            return FLASHV_ERR_STATE;
        }
        cudaGetLastError();
        for (int r = 0; r < p->shard_world; ++r)
            if (!p->peer_region[r]) {
                set_error("flashv_plan_run: peer %d of the sharded plan has not been connected", r);
                return FLASHV_ERR_STATE;
            }
        ++p->shard_run;
    }
    FV_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    bool first = true;
    const int first_level_pass = p->sched.first_pass ? 1 : 0;
    for (size_t pi = 0; pi < p->passes.size(); ++pi) {
        const Pass &ps = p->passes[pi];
        int rc = FLASHV_OK;
        if (ps.nvec > 0) rc = p->B > 0 ? bs_run_pass(p, ps) : flash_run_pass(p, ps, first);
        if (rc != FLASHV_OK) return rc;
        first = false;
        if (sharded && pi > 0) {  // pass 0 is the state-sharded one: every rank already holds its results
            rc = shard_ans_exchange(p, (int)pi - first_level_pass);
            if (rc != FLASHV_OK) return rc;
        }
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

extern "C" int flashv_trellis_step_columns_dev(flashv_model *m, const void *delta_in_dev, int o, int col_begin, int col_end,
                                               void *delta_out_dev, void *psi_out_dev)
{
    if (!m || !delta_in_dev || !delta_out_dev || !psi_out_dev || o < 0 || o >= m->M || col_begin < 0 || col_end > m->K ||
        col_begin >= col_end || m->Kp > STEP_MAX_KP) {
        set_error("flashv_trellis_step_columns_dev: bad argument");
        return FLASHV_ERR_ARG;
    }
    FV_CUDA(cudaSetDevice(m->ctx->device));
    return flash_step_columns(m, (const float *)delta_in_dev, o, col_begin, col_end, (float *)delta_out_dev, (int32_t *)psi_out_dev);
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
