// mgpu.cu — every GPU of a box behind the C ABI (the reference is single-process pthreads and
// has no counterpart: F:264-308 spreads tasks over MAX_THREADS host threads, here the same independence
// spreads sequences, destination states and tree tasks over GPUs).
//
//   * flashv_shard_count / flashv_decode_batch_shard: the per-rank form (one process per GPU, e.g. under
//     torchrun): rank r decodes the sequences r, r+G, r+2G, ... and leaves their rows in place.
//   * flashv_mgpu_*: the one-process form, one host thread per device per call.  The model's log tables
//     are computed once — every device's thread takes K/G rows of host logarithms — and exchanged over
//     NVLink; batches shard b mod G; a single sequence runs state-sharded (api.cu: flashv_plan_shard_*).
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "flashv_internal.h"

using namespace flashv;

extern "C" int flashv_shard_count(int total, int rank, int world)
{
    if (total < 0 || world < 1 || rank < 0 || rank >= world) return FLASHV_ERR_ARG;
    return total > rank ? (total - rank + world - 1) / world : 0;
}

static int batch_shard(flashv_model *m, const int32_t *ob, int total, int T, int N, int B, int rank, int world, int32_t *path_out,
                       float *score_out, flashv_report *report)
{
    const int mine = flashv_shard_count(total, rank, world);
    if (!m || !ob || !path_out || mine < 0 || T < 1) {
        set_error("flashv_decode_batch_shard: bad argument");
        return FLASHV_ERR_ARG;
    }
    if (report) memset(report, 0, sizeof(*report));
    if (mine == 0) return FLASHV_OK;
    // the strided rows, packed: the batched plan wants ob[batch][T] contiguous
    std::vector<int32_t> lob((size_t)mine * T), lpath((size_t)mine * T);
    std::vector<float> lscore((size_t)mine);
    for (int q = 0; q < mine; ++q) memcpy(&lob[(size_t)q * T], ob + ((size_t)rank + (size_t)q * world) * T, (size_t)T * 4);
    int rc = B > 0 ? flashv_bs_decode_batch(m, lob.data(), mine, T, N, B, lpath.data(), lscore.data(), report)
                   : flashv_decode_batch(m, lob.data(), mine, T, N, lpath.data(), lscore.data(), report);
    if (rc != FLASHV_OK) return rc;
    for (int q = 0; q < mine; ++q) {
        const size_t b = (size_t)rank + (size_t)q * world;
        memcpy(path_out + b * T, &lpath[(size_t)q * T], (size_t)T * 4);
        if (score_out) score_out[b] = lscore[q];
    }
    return FLASHV_OK;
}

extern "C" int flashv_decode_batch_shard(flashv_model *m, const int32_t *ob, int total, int T, int N, int rank, int world,
                                         int32_t *path_out, float *score_out, flashv_report *report)
{
    return batch_shard(m, ob, total, T, N, 0, rank, world, path_out, score_out, report);
}

extern "C" int flashv_bs_decode_batch_shard(flashv_model *m, const int32_t *ob, int total, int T, int N, int B, int rank, int world,
                                            int32_t *path_out, float *score_out, flashv_report *report)
{
    if (B < 1) {
        set_error("flashv_bs_decode_batch_shard: BeamSearchWidth must be >= 1");
        return FLASHV_ERR_ARG;
    }
    return batch_shard(m, ob, total, T, N, B, rank, world, path_out, score_out, report);
}

// ---- one process, all GPUs -------------------------------------------------------------------------
struct ShardedPlanSet {
    int T = 0, N = 0;
    std::vector<flashv_plan *> plans;  // one per device
};

struct flashv_mgpu {
    std::vector<int> devices;
    std::vector<flashv_ctx *> ctx;
    std::vector<flashv_model *> model;
    std::vector<ShardedPlanSet> sharded;  // small cache, like the one-call decodes keep
};

// Run fn(rank) on one host thread per device; the first failure's code and message come back on the
// calling thread (flashv_last_error is thread-local).
template <class F>
static int for_each_device(flashv_mgpu *g, F fn)
{
    const int W = (int)g->devices.size();
    std::vector<int> rc((size_t)W, FLASHV_OK);
    std::vector<std::string> msg((size_t)W);
    auto body = [&](int r) {
        rc[r] = fn(r);
        if (rc[r] != FLASHV_OK) msg[r] = flashv_last_error();
    };
    std::vector<std::thread> pool;
    for (int r = 1; r < W; ++r) pool.emplace_back(body, r);
    body(0);
    for (auto &t : pool) t.join();
    for (int r = 0; r < W; ++r)
        if (rc[r] != FLASHV_OK) {
            set_error("device %d: %s", g->devices[r], msg[r].c_str());
            return rc[r];
        }
    return FLASHV_OK;
}

extern "C" int flashv_mgpu_create(int ndev, const int *devices, flashv_mgpu **out)
{
    if (!out || ndev < 1 || ndev > 8) {
        set_error("flashv_mgpu_create: 1 to 8 devices");
        return FLASHV_ERR_ARG;
    }
    *out = nullptr;
    flashv_mgpu *g = new flashv_mgpu();
    for (int r = 0; r < ndev; ++r) g->devices.push_back(devices ? devices[r] : r);
    g->ctx.assign((size_t)ndev, nullptr), g->model.assign((size_t)ndev, nullptr);
    for (int r = 0; r < ndev; ++r) {
        int rc = flashv_ctx_create(g->devices[r], nullptr, &g->ctx[r]);
        if (rc != FLASHV_OK) {
            flashv_mgpu_destroy(g);
            return rc;
        }
    }
    // every device reads and writes every other one's tables and regions
    for (int r = 0; r < ndev; ++r) {
        cudaSetDevice(g->devices[r]);
        for (int q = 0; q < ndev; ++q)
            if (q != r) {
                cudaError_t e = cudaDeviceEnablePeerAccess(g->devices[q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    int rc = cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                    flashv_mgpu_destroy(g);
                    return rc;
                }
                cudaGetLastError();
            }
    }
    *out = g;
    return FLASHV_OK;
}

static void drop_models(flashv_mgpu *g)
{
    for (auto &set : g->sharded)
        for (flashv_plan *p : set.plans) flashv_plan_destroy(p);
    g->sharded.clear();
    for (auto &m : g->model) {
        if (m) flashv_model_destroy(m);
        m = nullptr;
    }
}

extern "C" void flashv_mgpu_destroy(flashv_mgpu *g)
{
    if (!g) return;
    drop_models(g);
    for (flashv_ctx *c : g->ctx)
        if (c) flashv_ctx_destroy(c);
    delete g;
}

extern "C" int flashv_mgpu_world(const flashv_mgpu *g) { return g ? (int)g->devices.size() : 0; }
extern "C" flashv_ctx *flashv_mgpu_ctx(flashv_mgpu *g, int rank)
{
    return g && rank >= 0 && rank < (int)g->ctx.size() ? g->ctx[rank] : nullptr;
}
extern "C" flashv_model *flashv_mgpu_model(flashv_mgpu *g, int rank)
{
    return g && rank >= 0 && rank < (int)g->model.size() ? g->model[rank] : nullptr;
}

extern "C" int flashv_mgpu_model_create(flashv_mgpu *g, int K, int M, const float *A, const float *B, const float *Pi)
{
    if (!g || !A || !B || !Pi) {
        set_error("flashv_mgpu_model_create: bad argument");
        return FLASHV_ERR_ARG;
    }
    drop_models(g);
    const int W = (int)g->devices.size();
    // 1. every device's thread: host logarithms of its rows, uploaded to its own table
    int rc = for_each_device(g, [&](int r) { return flashv_model_create_rows(g->ctx[r], K, M, A, B, Pi, r, W, &g->model[r]); });
    // 2. (threads joined = barrier) fetch the other rows over NVLink; 3. (barrier) layouts on the device
    if (rc == FLASHV_OK)
        rc = for_each_device(g, [&](int r) {
            for (int q = 0; q < W; ++q)
                if (q != r) {
                    int e = flashv_model_pull_rows_from(g->model[r], q, g->model[q]);
                    if (e != FLASHV_OK) return e;
                }
            return FLASHV_OK;
        });
    if (rc == FLASHV_OK) rc = for_each_device(g, [&](int r) { return flashv_model_finish(g->model[r]); });
    if (rc != FLASHV_OK) drop_models(g);
    return rc;
}

static void merge_report(flashv_report *into, const flashv_report &r, bool first)
{
    if (first) {
        *into = r;
        return;
    }
    into->decode_ms = std::max(into->decode_ms, r.decode_ms);
    into->first_pass_ms = std::max(into->first_pass_ms, r.first_pass_ms);
    into->h2d_ms = std::max(into->h2d_ms, r.h2d_ms), into->d2h_ms = std::max(into->d2h_ms, r.d2h_ms);
    into->device_bytes += r.device_bytes;
    into->kernel_launches += r.kernel_launches;
}

static int mgpu_batch(flashv_mgpu *g, const int32_t *ob, int batch, int T, int N, int B, int32_t *path_out, float *score_out,
                      flashv_report *report)
{
    if (!g || !ob || !path_out || batch < 1 || g->model.empty() || !g->model[0]) {
        set_error("flashv_mgpu_decode_batch: bad argument, or no model (flashv_mgpu_model_create first)");
        return FLASHV_ERR_ARG;
    }
    const int W = (int)g->devices.size();
    std::vector<flashv_report> reps((size_t)W);
    int rc = for_each_device(
        g, [&](int r) { return batch_shard(g->model[r], ob, batch, T, N, B, r, W, path_out, score_out, &reps[r]); });
    if (rc != FLASHV_OK) return rc;
    if (report) {
        bool first = true;
        for (int r = 0; r < W; ++r)
            if (flashv_shard_count(batch, r, W) > 0) merge_report(report, reps[r], first), first = false;
    }
    return FLASHV_OK;
}

extern "C" int flashv_mgpu_decode_batch(flashv_mgpu *g, const int32_t *ob, int batch, int T, int N, int32_t *path_out,
                                        float *score_out, flashv_report *report)
{
    return mgpu_batch(g, ob, batch, T, N, 0, path_out, score_out, report);
}

extern "C" int flashv_mgpu_bs_decode_batch(flashv_mgpu *g, const int32_t *ob, int batch, int T, int N, int B, int32_t *path_out,
                                           float *score_out, flashv_report *report)
{
    if (B < 1) {
        set_error("flashv_mgpu_bs_decode_batch: BeamSearchWidth must be >= 1");
        return FLASHV_ERR_ARG;
    }
    return mgpu_batch(g, ob, batch, T, N, B, path_out, score_out, report);
}

static int sharded_plans(flashv_mgpu *g, int T, int N, ShardedPlanSet **out)
{
    for (auto &set : g->sharded)
        if (set.T == T && set.N == N) {
            *out = &set;
            return FLASHV_OK;
        }
    const int W = (int)g->devices.size();
    ShardedPlanSet set;
    set.T = T, set.N = N;
    set.plans.assign((size_t)W, nullptr);
    auto fail = [&](int rc) {
        for (flashv_plan *p : set.plans) flashv_plan_destroy(p);
        return rc;
    };
    for (int r = 0; r < W; ++r) {
        int rc = flashv_plan_create(g->model[r], T, N, 1, 0, FLASHV_ENGINE_PERSISTENT, &set.plans[r]);
        if (rc == FLASHV_OK) rc = flashv_plan_shard_init(set.plans[r], r, W);
        if (rc != FLASHV_OK) return fail(rc);
    }
    if (W > 1)
        for (int r = 0; r < W; ++r)
            for (int q = 0; q < W; ++q)
                if (q != r) {
                    void *base = nullptr;
                    int rc = flashv_plan_shard_buffers(set.plans[q], &base, nullptr);
                    if (rc == FLASHV_OK) rc = flashv_plan_shard_set_peer(set.plans[r], q, g->devices[q], base);
                    if (rc != FLASHV_OK) return fail(rc);
                }
    if (g->sharded.size() >= 4) {
        for (flashv_plan *p : g->sharded.front().plans) flashv_plan_destroy(p);
        g->sharded.erase(g->sharded.begin());
    }
    g->sharded.push_back(std::move(set));
    *out = &g->sharded.back();
    return FLASHV_OK;
}

extern "C" int flashv_mgpu_decode(flashv_mgpu *g, const int32_t *ob, int T, int N, int32_t *path_out, float *score_out,
                                  flashv_report *report)
{
    if (!g || !ob || !path_out || g->model.empty() || !g->model[0]) {
        set_error("flashv_mgpu_decode: bad argument, or no model (flashv_mgpu_model_create first)");
        return FLASHV_ERR_ARG;
    }
    const int W = (int)g->devices.size();
    ShardedPlanSet *set = nullptr;
    int rc = sharded_plans(g, T, N, &set);
    if (rc != FLASHV_OK) return rc;
    // upload everywhere, make every stream idle (the sharded run's contract), launch everywhere, then
    // read rank 0's result: all ranks hold the same path
    for (int r = 0; r < W && rc == FLASHV_OK; ++r) rc = flashv_plan_upload(set->plans[r], ob);
    for (int r = 0; r < W && rc == FLASHV_OK; ++r) rc = flashv_ctx_sync(g->ctx[r]);
    for (int r = 0; r < W && rc == FLASHV_OK; ++r) rc = flashv_plan_run(set->plans[r]);
    if (rc != FLASHV_OK) return rc;
    std::vector<int32_t> other((size_t)T);
    for (int r = W - 1; r >= 0 && rc == FLASHV_OK; --r) rc = flashv_plan_download(set->plans[r], r == 0 ? path_out : other.data(), r == 0 ? score_out : nullptr);
    if (rc != FLASHV_OK) return rc;
    if (report) {
        for (int r = 0; r < W; ++r) {
            flashv_report rr;
            rc = flashv_plan_report(set->plans[r], &rr);
            if (rc != FLASHV_OK) return rc;
            merge_report(report, rr, r == 0);
        }
    }
    return FLASHV_OK;
}
