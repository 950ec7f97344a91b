// flashv_internal.h — shared declarations of the libflashv translation units.
// Not part of the ABI (include/flashv.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "flashv.h"

namespace flashv {

// ---- error plumbing --------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FV_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t fv_e_ = (call);                                            \
        if (fv_e_ != cudaSuccess) return ::flashv::cuda_fail(fv_e_, #call, __FILE__, __LINE__); \
    } while (0)

// ---- schedule (host, pure function of T and N) -----------------------------------------
struct Task {
    int L, R, mid;
};

// The reference's FIFO (F:284-304, F:349-359) flattened into tree levels: every task of level
// l+1 is a child of a task of level l, and tasks of one level are independent (each reads only
// Ans[] entries written by its ancestors), so a level advances in lock-step on the device.
struct Schedule {
    int T = 0, N = 0;
    bool first_pass = false;   // F:342
    std::vector<int> mids;     // F:129-136, N-1 entries when first_pass
    std::vector<Task> fifo;    // queue order, exactly the tasks the reference runs
    std::vector<std::vector<Task>> levels;  // same tasks, grouped by depth, longest first
    long long executed_steps = 0;
};
// returns false for the domain the reference mishandles (T < 2, N < 1, T == 2N with N > 2)
bool build_schedule(int T, int N, Schedule *out);
int pool_struct_bytes(int N);

// ---- device-side descriptors -------------------------------------------------------------
// One trellis vector in flight: sequence `seq` walking the interval (L,R).  Backpointers are
// kept only for steps j >= mid+1 (F:242: the tracker latches at j == mid+1), in rows
// psi_row .. psi_row + (R-mid-1) of the pass's backpointer store.
struct VecDesc {
    int seq, L, R, mid;
    int psi_row;
    int flags;
};
enum : int {
    VEC_FULL_RANGE = 1,   // (L,R) == (0,T-1): the pass also picks Ans[T-1] (F:186-196, F:249-259)
    VEC_FIRST_PASS = 2,   // nvviterNdivide: record every segment boundary on the way back (F:198-201)
};

struct Pass {
    int nvec = 0;        // vectors (tasks of the level x batch), sorted by steps descending
    int max_steps = 0;   // max (R-L)
    int psi_rows = 0;    // rows of the backpointer store this pass needs
    size_t vec_offset = 0;  // into the plan's device VecDesc array
    std::vector<int> nactive;  // nactive[s] = vectors still stepping at step s (1-based), prefix of the order
    size_t nactive_off = 0;    // where this pass's nactive[] starts in the plan's device copy
    bool full_range = false;
    VecDesc first_vec{};  // host copy of vector 0 (single-vector passes are driven from it)
};

}  // namespace flashv

struct flashv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int smem_optin = 0;
    int coop = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // pinned staging for the one-call API
    int32_t *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    // pinned double buffer + its events for the chunked log-table upload of flashv_model_create (tables.cu)
    void *h_prep[2] = {nullptr, nullptr};
    size_t h_prep_bytes = 0;
    cudaEvent_t ev_prep[2] = {nullptr, nullptr};
};

struct flashv_model {
    flashv_ctx *ctx = nullptr;
    int K = 0, M = 0;
    int Kp = 0;               // K rounded up to a multiple of 128 (one warp x float4)
    float *hiT = nullptr;     // [K][Kp]  (float)log A, destination-major: hiT[i][k] = log A[k][i]; pad = -inf
    float *hiC = nullptr;     // K*Kp     the same, CTA-tiled for the persistent engine (tile_geom.h)
    double *LAc = nullptr;    // K*4096   log A chain-major for the persistent engine's window scan (K <= 4096 only)
    double *LAcL = nullptr;   // K*128*clp the same for wider models: chains of clp = Kp/128 (rounded up to 32) elements
    int clp = 0;
    __half *hi16 = nullptr;   // K*Kp16   (half)log A, CTA-tiled for the half-precision filter (K <= 4096 only)
    double *LAc16 = nullptr;  // K*4096   log A chain-major for that filter: 256 chains of 16
    int Kp16 = 0;             // K rounded up to 256
    float *LBmax = nullptr;   // M         the largest emission term of every symbol
    double lamax = 0.0;       //           the largest log A (host copy)
    int *csc_ptr = nullptr;       // in-edge lists of the transition graph (flash_sparse.cu); null when the table is dense or K >= 65536
    uint16_t *csc_k = nullptr;
    double *csc_la = nullptr;
    long long csc_nnz = 0;
    int *csr_cut = nullptr;       // out-edge lists for FLASH-BS (bs_kernels.cu): row s, cut q = first entry with destination >= q*ceil(K/8); [K][9]
    uint16_t *csr_i = nullptr;    // [nnz] destination state, ascending within a row
    double *csr_la = nullptr;     // [nnz] log A[s][i]
    int csc_max_cta_nnz = 0;
    float *hiS = nullptr;     // [Kp][Kp] the same, source-major (hiS[k][i]), for the group engine; only when Kp <= 1536
    int tile_G = 0;           // grid the tiling was built for (min(#SM, K))
    double *LAd = nullptr;    // [K][K]   log A, source-major as the reference stores A (F:27)
    float *LBf = nullptr;     // [M][Kp]  (float)log B, symbol-major: LBf[o][i] — the per-step "tmp" (F:167)
    double *LBd = nullptr;    // [M][K]   log B (double), symbol-major — start vectors (F:142, F:220)
    double *LPi = nullptr;    // [K]      log Pi
    float *scratch_f = nullptr;   // test hooks: 2*Kp floats
    int32_t *scratch_i = nullptr; // test hooks: 2*Kp ints
    void *scratch_x = nullptr;    // test hooks: exchange buffers [2][Kp] x 8 B
    size_t bytes = 0;
    double prep_ms = 0;
    int row_lo = 0, row_hi = 0;   // rows of LAd whose logarithms this process computed (all of them unless the model was created in parts)
    bool ready = false;           // layouts built: the model can decode
    int part_rank = 0, part_world = 1;  // flashv_model_create_rows: which rows are local
    std::vector<flashv_plan *> plan_cache;  // owned; used by the one-call decodes
};

struct flashv_plan {
    flashv_model *model = nullptr;
    int T = 0, N = 0, batch = 0, B = 0, engine = 0;
    flashv::Schedule sched;
    std::vector<flashv::Pass> passes;  // [0] = first pass when sched.first_pass, then one per level
    int max_vec = 0;
    int psi16 = 0;
    // device
    int32_t *d_ob = nullptr;      // [batch][T]
    int32_t *d_ans = nullptr;     // [batch][T]
    float *d_score = nullptr;     // [batch]
    float *d_delta = nullptr;     // [2][max_vec][Kp] + exchange buffers [2][Kp] x 8 B of the persistent engine
    void *d_psi = nullptr;        // [max psi_rows][K] u16 or i32
    flashv::VecDesc *d_vecs = nullptr;
    uint8_t *d_ismid = nullptr;   // [T] 1 where a first-pass segment boundary sits
    int *d_nactive = nullptr;     // every pass's nactive[] back to back (the persistent level kernel reads it)
    int32_t *d_endstate = nullptr;  // [max_vec]
    int32_t *d_btmap = nullptr;     // [bt_windows][K] composed backpointer maps + [bt_windows] start states (parallel walk back)
    int bt_windows = 0;
    unsigned int *d_sync = nullptr; // 256 zeroed bytes: the level kernel's grid-barrier counter (64-bit, monotone)
    unsigned long long bar_count = 0;  // what that counter holds once every launch issued so far has finished
    // state sharding of single-vector passes
    int shard_rank = 0, shard_world = 1;
    int shard_c0 = 0, shard_ncol = 0;   // destination columns this GPU owns
    float *hiC_shard = nullptr;         // their slice of the tiled table
    // One device region per sharded plan holds everything peers store into, and nothing the local
    // passes reuse: [exchange words 2 x Kp x 8 B][Ans exchange words T x 8 B][backpointer rows of pass 0]
    unsigned char *shard_region = nullptr;
    size_t shard_region_bytes = 0, shard_ans_off = 0, shard_psi_off = 0;
    unsigned char *peer_region[8] = {};  // every GPU's region (own entry included)
    bool peer_ipc[8] = {};              // opened with cudaIpcOpenMemHandle (to be closed)
    unsigned shard_run = 0;             // counts the runs of a sharded plan: the same on every rank, tags all cross-GPU words
    int32_t *d_lvl_mid = nullptr;       // midpoints of every task, level after level (all ranks' tasks): the Ans entries a level produces
    std::vector<int> lvl_off, lvl_cnt;  // per tree level: offset into d_lvl_mid, task count
    unsigned run_epoch = 0;             // tags the exchange words of the local (unsharded) persistent passes
    // FLASH-BS
    float *d_bs_score = nullptr;  // [max_vec][Kp]
    size_t bytes = 0;
    // last run
    flashv_report rep{};
    bool uploaded = false, ran = false;
    int launches = 0;
};

// ---- kernels / stages (defined in the .cu files) -------------------------------------------
namespace flashv {

int tables_alloc(flashv_model *m);
int tables_logs(flashv_model *m, const float *A, const float *B, const float *Pi, int row_lo, int row_hi, int share);
int tables_layouts(flashv_model *m);
void build_tiled_slice(const double *LAd, float *hiC, int K, int Kp, int col_begin, int ncol, int G, cudaStream_t st);
int shard_build_table(flashv_plan *p);

int flash_run_pass(flashv_plan *p, const Pass &pass, bool time_it);
int sparse_build(flashv_model *m);                                  // flash_sparse.cu: edge lists from the device table
bool sparse_engine_available(const flashv_model *m);
int sparse_pass(flashv_plan *p, const Pass &pass);
int sparse_level_step(flashv_plan *p, const Pass &pass, int s, int nact, const float *din, float *dout);
bool group_engine_fits(const flashv_model *m);                       // flash_group.cu
int group_run_pass(flashv_plan *p, const Pass &pass, float *dfinal);  // flash_group.cu
constexpr int GROUP_MAX_KP = 1536;  // largest padded K the group engine's shared-memory buffers hold
int bs_run_pass(flashv_plan *p, const Pass &pass);
bool persistent_engine_fits(const flashv_ctx *ctx, int Kp);           // flash_persistent.cu
constexpr int STEP_MAX_KP = 220 * 1024 / 4;  // the per-step kernels keep one delta vector (Kp floats) in shared memory
int shard_ans_exchange(flashv_plan *p, int level);                    // flash_persistent.cu: Ans entries of a level to every rank
inline bool pass_is_sharded(const flashv_plan *p, const Pass &pass) { return p->shard_world > 1 && &pass == &p->passes[0]; }
inline void *pass_psi(const flashv_plan *p, const Pass &pass)
{
    return pass_is_sharded(p, pass) ? (void *)(p->shard_region + p->shard_psi_off) : p->d_psi;
}

int flash_single_step(flashv_model *m, const float *d_in_dev, int o, float *d_out_dev, int32_t *psi_dev, int engine);
int flash_single_init(flashv_model *m, int prev_state, int o, float *d_out_dev);
int flash_step_columns(flashv_model *m, const float *d_in_dev, int o, int c_begin, int c_end, float *d_out_dev, int32_t *psi_dev);
int bs_single_score(flashv_model *m, const float *hv_dev, const int32_t *hs_dev, int B, int o, float *score_dev,
                    int32_t *arg_dev);
int bs_single_replay(flashv_ctx *ctx, const float *score_dev, int K, int B, float *hv_dev, int32_t *hs_dev);

}  // namespace flashv
