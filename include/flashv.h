/*
 * flashv.h — C ABI of the B200-native FLASH / FLASH-BS Viterbi decoder (libflashv.so).
 *
 * Drop-in boundary for the reference's hot path.  The reference exports no library symbol:
 * its interface is the program shell around calc().  Each entry point below
 * names the reference function it replaces:
 *
 *   F: = /root/reference/src/FLASH_Viterbi_multithread.c
 *   S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
 *
 * Plain C types only; all pointers are HOST pointers unless a name says "dev".  Every
 * function returns FLASHV_OK (0) or a negative FLASHV_ERR_* code; flashv_last_error() gives
 * the message of the last failure on the calling thread.  There is no CPU fallback: every
 * decode runs on the CUDA device of its context or fails.
 */
#ifndef FLASHV_H
#define FLASHV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLASHV_OK 0
#define FLASHV_ERR_ARG (-1)     /* bad argument (NULL, sizes, T < 2, T == 2N with N > 2, B > K ...) */
#define FLASHV_ERR_DOMAIN (-2)  /* A/B/Pi entry outside [0,1] or not finite */
#define FLASHV_ERR_CUDA (-3)    /* CUDA runtime failure (message holds cudaGetErrorString) */
#define FLASHV_ERR_NOMEM (-4)   /* host or device allocation failed */
#define FLASHV_ERR_STATE (-5)   /* call order (run before upload, ...) */

typedef struct flashv_ctx flashv_ctx;     /* one GPU + one stream + scratch; one host thread at a time */
typedef struct flashv_model flashv_model; /* device-resident log tables of one HMM */
typedef struct flashv_plan flashv_plan;   /* decode schedule + workspace for one (T, N, batch[, B]) */

/* Which trellis engine the FLASH plan uses for single-vector passes (the N-way first pass and
 * the root task).  All are exact.  AUTO picks the persistent kernel when it fits the device
 * (cooperative launch; delta vector + two ring stages in one CTA's shared memory: K up to ~42,000 on
 * B200), else the per-step kernels.  Feasibility is decided by flashv_plan_create: FLASH plans accept
 * K <= 56,320 (one delta vector in shared memory); a forced engine that does not fit is FLASHV_ERR_ARG. */
#define FLASHV_ENGINE_AUTO 0
#define FLASHV_ENGINE_STEP 1       /* one kernel launch per trellis step */
#define FLASHV_ENGINE_PERSISTENT 2 /* one cooperative launch per pass; steps hand over through tagged delta words, no grid barrier.  Models up
                                      to 4096 states: the table slice of every SM stays in tensor memory and the sweep is a half-precision
                                      filter (every candidate inside its window is then evaluated with the reference's arithmetic: results
                                      are bit-identical); wider models: float sweep, TMA-fed ring, first 2048 states in tensor memory */
#define FLASHV_ENGINE_SPARSE 3     /* opt-in: the pass walks the in-edge lists (entries with A[k][i] == 0 can never win,
                                      F:171), resident in shared memory; same results, not the dense roofline's bytes */

typedef struct flashv_report {
    double decode_ms;          /* device time of the decode proper (CUDA events, tables resident) */
    double first_pass_ms;      /* device time of the full-length pass alone (0 if none) */
    double h2d_ms, d2h_ms;     /* observation upload / path download, host wall clock */
    long long executed_steps;  /* S(T,N) per sequence: (T-1 if first pass) + sum (R-L) over tasks */
    long long device_bytes;    /* tables + plan workspace actually allocated on the device */
    int memory_bytes;          /* the reference's "memory:" formula (F:355,364-367 / S:564,573-576) */
    int first_pass;            /* 1 iff the N-way pass ran (F:342) */
    int n_tasks;               /* queue tasks per sequence (T-N or T-1) */
    int n_levels;              /* task-tree levels executed */
    int kernel_launches;       /* kernels launched by this decode */
    int engine;                /* FLASHV_ENGINE_* actually used for the full-length pass */
} flashv_report;

const char *flashv_last_error(void);
const char *flashv_version(void);

/* ---- context ------------------------------------------------------------------------- */
/* stream: a cudaStream_t created by the caller on `device` (e.g. torch's), or NULL to let
 * the context create its own non-blocking stream.  All work of the context is issued there. */
int flashv_ctx_create(int device, void *stream, flashv_ctx **out);
void flashv_ctx_destroy(flashv_ctx *ctx);
void *flashv_ctx_stream(const flashv_ctx *ctx);
int flashv_ctx_sync(flashv_ctx *ctx);
int flashv_ctx_sm_count(const flashv_ctx *ctx);

/* ---- model: replaces create_vit() (F:97-107) and the in-loop log() calls (F:142,167,170) */
/* A[K][K] (row = source, column = destination), B[K][M], Pi[K]: the float values the
 * reference's loader leaves in VIT (F:25-34).  Log tables are computed on the HOST with the
 * host libm (the same log() the reference calls), then laid out on the device. */
int flashv_model_create(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi,
                        flashv_model **out);
void flashv_model_destroy(flashv_model *model);
int flashv_model_K(const flashv_model *model);
int flashv_model_M(const flashv_model *model);
double flashv_model_prep_ms(const flashv_model *model); /* host log tables + upload + layout */

/* Program-shell loader (F:56-95): fscanf("%f") / fscanf("%d") semantics. Returns count read or <0. */
long flashv_read_floats_text(const char *path, long n, float *out);
long flashv_read_ints_text(const char *path, long n, int32_t *out);
/* The same values through a binary side-car (<path>.f32cache, written after the first parse and valid while the
 * text file's size and mtime are unchanged).  The text stays canonical (data_script.py:98-101); parsing K*K
 * decimal numbers is what dominates a run of the reference-shaped program (seconds against a millisecond decode). */
long flashv_read_floats_cached(const char *path, long n, float *out);

/* ---- one-call decodes: replace calc() ------------------------------------------------ */
/* FLASH, calc() of F:338-368.  ob[T] in [0,M); N = MAX_THREADS (the segment count).
 * path_out[T]; score_out (may be NULL) = max_i delta_{T-1}[i] of the full-range pass. */
int flashv_decode(flashv_model *model, const int32_t *ob, int T, int N, int32_t *path_out, float *score_out,
                  flashv_report *report);
/* FLASH-BS, calc() of S:548-577.  B = BeamSearchWidth (1 <= B <= K).  path entries may be -1. */
int flashv_bs_decode(flashv_model *model, const int32_t *ob, int T, int N, int B, int32_t *path_out,
                     float *score_out, flashv_report *report);
/* `batch` independent sequences of one HMM: ob[batch][T] -> path_out[batch][T], score_out[batch]. */
int flashv_decode_batch(flashv_model *model, const int32_t *ob, int batch, int T, int N, int32_t *path_out,
                        float *score_out, flashv_report *report);
int flashv_bs_decode_batch(flashv_model *model, const int32_t *ob, int batch, int T, int N, int B,
                           int32_t *path_out, float *score_out, flashv_report *report);

/* Sanity path: viterbi() of "Base_line/C implementations/vanilla Viterbi.c":124-171, the textbook
 * decoder the reference ships as its baseline — every backpointer kept, one backtrack, and ITS arithmetic
 * (T1 + log A + log B in double, one rounding, :140), which is not FLASH's (F:170): bit-equal to the reference's
 * vanilla program, equal to FLASH where the reference's own two programs agree.  report->memory_bytes is that
 * program's `memory:` line (sizeof(T1)+sizeof(T2), :172).  FLASHV_ERR_DOMAIN if the path runs through a state
 * no transition reaches (the reference reads T2[-1] there). */
int flashv_vanilla_decode(flashv_model *model, const int32_t *ob, int T, int32_t *path_out, float *score_out,
                          flashv_report *report);

/* ---- staged decodes (what the one-call forms do; lets a caller keep inputs resident) ---- */
/* B == 0 selects FLASH, B >= 1 FLASH-BS. */
int flashv_plan_create(flashv_model *model, int T, int N, int batch, int B, int engine, flashv_plan **out);
void flashv_plan_destroy(flashv_plan *plan);
int flashv_plan_upload(flashv_plan *plan, const int32_t *ob);                 /* H2D, async on the ctx stream */
int flashv_plan_run(flashv_plan *plan);                                       /* kernels only, async */
int flashv_plan_download(flashv_plan *plan, int32_t *path_out, float *score_out); /* D2H + stream sync */
int flashv_plan_report(flashv_plan *plan, flashv_report *report);            /* of the last run (syncs) */

/* ---- model creation shared by the ranks of a box ------------------------------------------------ */
/* What model creation costs is one host libm log() per table entry (F:170).  With `world` ranks on a
 * box (one process per GPU) each rank computes the rows [rank*K/world, (rank+1)*K/world) only, and the
 * rest arrives from the peers' device tables over NVLink:
 *     flashv_model_create_rows(...)                 on every rank (A may be a shared mapping; only the
 *                                                   rank's rows of it are read)
 *     flashv_model_rows_handle(model, h)            64-byte cudaIpc handle; all-gather them; BARRIER
 *     flashv_model_pull_rows(model, r, handle_r)    for every peer r (flashv_model_pull_rows_from inside
 *                                                   one process, no handle)
 *     BARRIER (peers may still be pulling from this rank's table) ; flashv_model_finish(model)
 * The model decodes only after flashv_model_finish (layouts and edge lists are built on the device). */
int flashv_model_create_rows(flashv_ctx *ctx, int K, int M, const float *A, const float *B, const float *Pi, int rank,
                             int world, flashv_model **out);
int flashv_model_rows_handle(flashv_model *model, void *out64);
int flashv_model_pull_rows(flashv_model *model, int peer_rank, const void *handle64);
int flashv_model_pull_rows_from(flashv_model *model, int peer_rank, const flashv_model *peer);
int flashv_model_finish(flashv_model *model);

/* ---- batches sharded over ranks (sequence b -> GPU b mod G, no data-path collective) --- */
/* How many of `total` sequences rank `rank` of `world` decodes, and which: rank, rank+world, ... */
int flashv_shard_count(int total, int rank, int world);
/* Decode this rank's share of ob[total][T] and leave the rows in place in path_out[total][T] /
 * score_out[total] (rows of other ranks are not touched): the caller gathers with whatever transport it
 * has (MPI, torch.distributed, shared memory) — or uses flashv_mgpu_decode_batch inside one process. */
int flashv_decode_batch_shard(flashv_model *model, const int32_t *ob, int total, int T, int N, int rank, int world,
                              int32_t *path_out, float *score_out, flashv_report *report);
/* The same for FLASH-BS (S:548-577 per sequence); a single FLASH-BS sequence does not shard (replicas only). */
int flashv_bs_decode_batch_shard(flashv_model *model, const int32_t *ob, int total, int T, int N, int B, int rank, int world,
                                 int32_t *path_out, float *score_out, flashv_report *report);

/* ---- state sharding of one huge K across the GPUs of a box -------------------- */
/* Every rank holds the full model and an identical FLASH plan (batch 1, persistent engine).  After
 * flashv_plan_shard_init(plan, rank, world) a rank computes only its slice of destination states
 * in the plan's pass 0 (the N-way pass of F:126-202, or the root task); each step it stores its slice
 * of delta and of the backpointer row into the regions of ALL ranks with in-kernel peer stores over
 * NVLink and polls only its own copy — the per-step all-gather, with no host or NCCL call.  The task
 * tree that follows is spread over the ranks as well: rank r runs every world-th task of a level, and
 * the level's Ans[] entries travel the same way (one tagged 64-bit word each) before the next level
 * starts.  Ranks exchange their region once: the raw pointer inside one process (set_peer enables peer
 * access), a 64-byte cudaIpc handle between processes (e.g. through torch.distributed.all_gather).
 * CONTRACT: peers store into a plan's region during a run, so every rank must have COMPLETED run n
 * (stream synchronised) before ANY rank calls flashv_plan_run for run n+1 — sync, barrier across ranks,
 * run.  flashv_plan_run returns FLASHV_ERR_STATE if this rank's stream is still busy.  A rank whose peers
 * do not show up within FLASHV_WATCHDOG_MS (default 30,000 across GPUs) traps instead of hanging. */
int flashv_plan_shard_init(flashv_plan *plan, int rank, int world);
int flashv_plan_shard_buffers(flashv_plan *plan, void **region_base, size_t *region_bytes);
int flashv_plan_shard_ipc_handle(flashv_plan *plan, void *out64);
int flashv_plan_shard_set_peer(flashv_plan *plan, int peer_rank, int peer_device, void *region_base);
int flashv_plan_shard_open_peer(flashv_plan *plan, int peer_rank, const void *handle64);

/* ---- all GPUs of a box from ONE host process (what a C program like the reference's shell uses) ---- */
/* A flashv_mgpu owns one context (and one host thread per call) per device.  The model is created once:
 * every device's thread computes its share of the log tables and the shares are exchanged over NVLink. */
typedef struct flashv_mgpu flashv_mgpu;
int flashv_mgpu_create(int ndev, const int *devices /* NULL: 0..ndev-1 */, flashv_mgpu **out);
void flashv_mgpu_destroy(flashv_mgpu *g);
int flashv_mgpu_world(const flashv_mgpu *g);
flashv_ctx *flashv_mgpu_ctx(flashv_mgpu *g, int rank);
flashv_model *flashv_mgpu_model(flashv_mgpu *g, int rank);
int flashv_mgpu_model_create(flashv_mgpu *g, int K, int M, const float *A, const float *B, const float *Pi);
/* Batch of independent sequences: sequence b on device b mod G, paths gathered into path_out[batch][T].
 * report: decode_ms = the slowest device's. */
int flashv_mgpu_decode_batch(flashv_mgpu *g, const int32_t *ob, int batch, int T, int N, int32_t *path_out,
                             float *score_out, flashv_report *report);
int flashv_mgpu_bs_decode_batch(flashv_mgpu *g, const int32_t *ob, int batch, int T, int N, int B, int32_t *path_out,
                                float *score_out, flashv_report *report);
/* One sequence over a K too large for one GPU to be fast: pass 0 state-sharded, tree levels spread (above). */
int flashv_mgpu_decode(flashv_mgpu *g, const int32_t *ob, int T, int N, int32_t *path_out, float *score_out,
                       flashv_report *report);

/* ---- pieces of the pass, exposed for per-step parity tests ---------------------------- */
/* Start vector of nvviter / nvviterNdivide (F:142, F:220): prev_state < 0 selects the pi form. */
int flashv_trellis_init(flashv_model *model, int prev_state, int ob0, float *delta_out);
/* One max-plus step (F:165-174): delta_in[K], symbol o -> delta_out[K], psi_out[K] (-1 = dead). */
int flashv_trellis_step(flashv_model *model, const float *delta_in, int o, float *delta_out, int32_t *psi_out,
                        int engine);
/* The same step for the destination states [col_begin, col_end) only, DEVICE pointers, asynchronous on the
 * context's stream: delta_in_dev[Kp] (padding beyond K finite), delta_out_dev[Kp] and psi_out_dev[K] (int32)
 * receive the entries of that range.  What a state-sharded pass looks like when the per-step exchange is
 * left to a library collective (tools/nccl_baseline.py: ncclAllGather between launches, the baseline of
 * SURVEY 8e) instead of the in-kernel peer stores of flashv_plan_shard_*. */
int flashv_trellis_step_columns_dev(flashv_model *model, const void *delta_in_dev, int o, int col_begin, int col_end,
                                    void *delta_out_dev, void *psi_out_dev);
/* One FLASH-BS score step (S:437-446) and the heap rebuild (S:167-211). */
int flashv_bs_score_step(flashv_model *model, const float *heap_val, const int32_t *heap_state, int B, int o,
                         float *score_out, int32_t *arg_slot_out);
int flashv_bs_heap_replay(flashv_ctx *ctx, const float *score, int K, int B, float *heap_val_out,
                          int32_t *heap_state_out);

/* ---- host-side pure functions ---------------------------------------------------------- */
/* The FIFO task list of F:284-304 / F:349-359: returns the task count, fills L[], R[] (capacity
 * T), *first_pass and mids[N-1] (may be NULL). */
int flashv_task_list(int T, int N, int *L, int *R, int *first_pass, int *mids);
long long flashv_executed_steps(int T, int N);
int flashv_memory_bytes(int K, int T, int N);    /* F:355, F:364-367 */
int flashv_bs_memory_bytes(int T, int N, int B); /* S:564, S:573-576 */

#ifdef __cplusplus
}
#endif
#endif
