/*
 * flashv_oracle.h — CPU restatement of the reference FLASH / FLASH-BS decoders.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under flash-viterbi_b200/ may include, link or call this.
 * The only legitimate users are tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs, and there only as the checker (or the timed CPU baseline), never as
 * the product path.
 *
 * Parity status: PINNED.  Every entry point is checked against the unmodified reference
 * binaries (built by oracle/build_ref.py from /root/reference/src/*.c with the exact
 * src/run.py:29-54 recipe) on the golden vectors under tests/golden/ — see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py.
 *
 * Citation shorthand:  F: = /root/reference/src/FLASH_Viterbi_multithread.c
 *                      S: = /root/reference/src/FLASH_BS_Viterbi_multithread.c
 */
#ifndef FLASHV_ORACLE_H
#define FLASHV_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fvo_model fvo_model;

/* K states, M observation symbols; A[K][K] row=source col=dest, B[K][M], Pi[K] — the float
 * arrays the reference holds in VIT (F:25-34) after fscanf("%f") (F:85-91).  Builds the
 * libm log() tables once; log(x) of a given x is deterministic, so hoisting it out of the
 * inner loop (F:170) does not change a single bit. */
fvo_model *fvo_model_create(int K, int M, const float *A, const float *B, const float *Pi);
/* The same model for FLASH decodes of shapes whose K x K double tables do not fit the host (K=32768:
 * 2 x 8.6 GB): only the entries with A[k][i] > 0 are kept, as ascending per-column lists.  A zero entry
 * gives ktmp = -inf, which never passes the strict '>' from -FLT_MAX (F:167, F:171), so the decode is the
 * same function; checked against the literal form on the golden models.  A is BORROWED (start vectors
 * read it, F:220) and must outlive the model.  FLASH only: the fvo_bs_* entry points need the full form. */
fvo_model *fvo_model_create_lean(int K, int M, const float *A, const float *B, const float *Pi);
void fvo_model_free(fvo_model *m);

/* calc() of F:338-368 for one observation sequence.  path[T]; *score = max_i delta_{T-1}[i] of
 * the full-range pass (F:188-193); *memory_bytes = the formula of F:355,364-367.
 * Returns 0, or -1 for the T == 2N case the reference mishandles (SURVEY §8a "edge"). */
int fvo_flash_decode(const fvo_model *m, const int *ob, int T, int N,
                     int *path, float *score, int *memory_bytes);

/* calc() of S:548-577.  Bw = BeamSearchWidth (requires Bw <= K). path entries may be -1
 * (S:73-86).  *score = Value of the slot chosen by the end scan (S:376-383 / S:456-463). */
int fvo_bs_decode(const fvo_model *m, const int *ob, int T, int N, int Bw,
                  int *path, float *score, int *memory_bytes);

/* viterbi() of "Base_line/C implementations/vanilla Viterbi.c":124-171 — the textbook O(K*T)-memory decoder the
 * reference ships as a baseline, with ITS arithmetic (candidate = T1 + log A + log B in double, one rounding),
 * which is not FLASH's (F:170), so paths may legitimately differ on near-ties: a sanity path, never the parity
 * oracle of FLASH.  Returns -1 if the decoded path would run through a dead column (the reference reads T2[-1]). */
int fvo_vanilla_decode(const fvo_model *m, const int *ob, int T, int *path, float *score, int *memory_bytes);

/* One dense trellis step (F:165-174): d_out[i], psi[i] from d_in[k] and symbol o. */
void fvo_flash_step(const fvo_model *m, const float *d_in, int o, float *d_out, int *psi);

/* Start vector of a (sub-)pass: F:142 when prev_state < 0 (the pi form), else F:220. */
void fvo_flash_init(const fvo_model *m, int prev_state, int o, float *d_out);

/* One beam step (S:437-449) on a heap given as parallel arrays of Bw slots (1-based slot s is
 * index s-1): scores[i], arg_slot[i] (0-based slot or -1). */
void fvo_bs_score_step(const fvo_model *m, const float *hval, const int *hstate, int Bw, int o,
                       float *score, int *arg_slot);

/* Heap replay of S:167-211 over score[0..K-1] (payload[i] carried along): fills hval/hstate/
 * hpay[Bw] in heap-array order. */
void fvo_bs_heap_replay(int K, int Bw, const float *score, const int *payload,
                        float *hval, int *hstate, int *hpay);

/* The FIFO task list (F:284-304, F:349-359): writes up to T entries of (L,R) in queue order,
 * returns the number of tasks; *first_pass = 1 iff the N-way pass runs (F:342); mids[N-1]. */
int fvo_task_list(int T, int N, int *L, int *R, int *first_pass, int *mids);

/* Executed trellis steps S(T,N) = (T-1 if first pass) + sum over tasks (R-L). */
long fvo_executed_steps(int T, int N);

/* The reference's "memory:" formulas. */
int fvo_flash_memory_bytes(int K, int T, int N);
int fvo_bs_memory_bytes(int T, int N, int Bw);

/* fscanf("%f") / fscanf("%d") ingest exactly as F:76-91; returns the number of values read. */
long fvo_read_floats(const char *path, long n, float *out);
long fvo_read_ints(const char *path, long n, int *out);

#ifdef __cplusplus
}
#endif
#endif
