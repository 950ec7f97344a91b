// vanilla.cu — the reference's vanilla Viterbi baseline on the device, as a SANITY path beside FLASH
// (SURVEY §8f-4).  Restates viterbi() of "Base_line/C implementations/vanilla Viterbi.c":124-171 (V: below):
// one forward pass with every backpointer kept (T2[K][T], V:123) and one backtrack (V:167-170).  Its
// arithmetic is NOT FLASH's: the candidate is  T1[k][j-1] + log A[k][i] + log B[i][o_j]  (V:140) — float
// promoted to double, two double adds left to right, one rounding — so its score differs from FLASH's in the
// last bits and its path may differ on near-ties; tests compare it with the reference's own vanilla program,
// and with FLASH only where the reference's two programs agree themselves.
// Nothing clever here on purpose: an exact double evaluation of every (k, i) pair, no estimate, no window.
#include <string.h>

#include "flashv_internal.h"
#include "trellis_common.cuh"

namespace flashv {

// T1[i][0] = log(Pi[i]) + log(B[i][o_0])  (V:120, V:128): double + double, one rounding
__global__ void k_vanilla_init(const double *__restrict__ LPi, const double *__restrict__ LBd, int K, int o, float *__restrict__ d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) d[i] = __double2float_rn(__dadd_rn(LPi[i], LBd[(size_t)o * K + i]));
}

// One time step (V:133-150): one thread per destination state i, source states k ascending, the previous
// column in shared memory; reads of log A[k][i] are coalesced across the threads of a warp.
__global__ void __launch_bounds__(128) k_vanilla_step(const double *__restrict__ LAd, const double *__restrict__ LBd, int K, int o,
                                                      const float *__restrict__ din, float *__restrict__ dout,
                                                      int32_t *__restrict__ psi_row)
{
    extern __shared__ float sprev[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) sprev[k] = din[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const double lb = LBd[(size_t)o * K + i];
    float best = -FLT_MAX;
    int arg = -1;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float cand = __double2float_rn(__dadd_rn(__dadd_rn((double)sprev[k], __ldg(LAd + (size_t)k * K + i)), lb));  // V:140
        if (cand > best) best = cand, arg = k;  // V:141-145
    }
    dout[i] = best;
    psi_row[i] = arg;
}

// Last column (V:152-161) and the backtrack (V:167-170), one CTA: first maximum from (-FLT_MAX, -1).
__global__ void __launch_bounds__(256) k_vanilla_finish(const float *__restrict__ d, const int32_t *__restrict__ psi, int K, int T,
                                                        int32_t *__restrict__ path, float *__restrict__ score, int *__restrict__ status)
{
    __shared__ float sx[8];
    __shared__ int sk[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Best b{-FLT_MAX, 0x7fffffff};
    for (int i = tid; i < K; i += 256) {
        const float x = d[i];
        if (x > b.x || (x == b.x && x > -FLT_MAX && i < b.k)) b.x = x, b.k = i;
    }
    b = warp_best(b);
    if (lane == 0) sx[warp] = b.x, sk[warp] = b.k;
    __syncthreads();
    if (tid == 0) {
        Best r{sx[0], sk[0]};
        for (int w = 1; w < 8; ++w) best_take(r, sx[w], sk[w]);
        int state = r.x > -FLT_MAX ? r.k : -1;
        *score = r.x;
        int ok = state >= 0;
        path[T - 1] = state;
        for (int j = T - 1; j > 0 && ok; --j) {
            state = psi[(size_t)(j - 1) * K + state];  // row j-1 of the store holds T2[.][j]
            ok = state >= 0;
            path[j - 1] = state;
        }
        *status = ok ? 0 : 1;  // 1: the path runs through a dead column — the reference reads T2[-1] there
    }
}

}  // namespace flashv

using namespace flashv;

extern "C" int flashv_vanilla_decode(flashv_model *m, const int32_t *ob, int T, int32_t *path_out, float *score_out,
                                     flashv_report *report)
{
    if (!m || !ob || !path_out || T < 1 || !m->ready) {
        set_error("flashv_vanilla_decode: bad argument");
        return FLASHV_ERR_ARG;
    }
    const int K = m->K;
    for (int j = 0; j < T; ++j)
        if (ob[j] < 0 || ob[j] >= m->M) {
            set_error("flashv_vanilla_decode: observation %d = %d outside [0,%d)", j, ob[j], m->M);
            return FLASHV_ERR_ARG;
        }
    if ((size_t)K * 4 > 200 * 1024) {
        set_error("flashv_vanilla_decode: K=%d does not fit the sanity path's shared-memory column", K);
        return FLASHV_ERR_ARG;
    }
    flashv_ctx *ctx = m->ctx;
    FV_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    float *d = nullptr, *d_score = nullptr;
    int32_t *psi = nullptr, *d_path = nullptr;
    int *d_status = nullptr;
    const size_t psi_rows = T > 1 ? (size_t)(T - 1) : 1;
    cudaError_t e = cudaMalloc(&d, (size_t)2 * K * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&psi, psi_rows * K * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&d_path, (size_t)T * sizeof(int32_t) + sizeof(float) + sizeof(int));
    int rc = FLASHV_OK;
    if (e != cudaSuccess) rc = cuda_fail(e, "vanilla workspace", __FILE__, __LINE__);
    int status = 0;
    float score = 0.f;
    float ms = 0.f;
    if (rc == FLASHV_OK) {
        d_score = reinterpret_cast<float *>(d_path + T);
        d_status = reinterpret_cast<int *>(d_score + 1);
        cudaFuncSetAttribute(k_vanilla_step, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaEventRecord(ctx->ev[0], st);
        k_vanilla_init<<<(K + 255) / 256, 256, 0, st>>>(m->LPi, m->LBd, K, ob[0], d);
        for (int j = 1; j < T; ++j)
            k_vanilla_step<<<(K + 127) / 128, 128, (size_t)K * sizeof(float), st>>>(m->LAd, m->LBd, K, ob[j], d + (size_t)((j - 1) & 1) * K,
                                                                                 d + (size_t)(j & 1) * K, psi + (size_t)(j - 1) * K);
        k_vanilla_finish<<<1, 256, 0, st>>>(d + (size_t)((T - 1) & 1) * K, psi, K, T, d_path, d_score, d_status);
        cudaEventRecord(ctx->ev[1], st);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(path_out, d_path, (size_t)T * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&score, d_score, sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        if (e != cudaSuccess) rc = cuda_fail(e, "vanilla decode", __FILE__, __LINE__);
    }
    cudaFree(d), cudaFree(psi), cudaFree(d_path);
    if (rc != FLASHV_OK) return rc;
    if (status != 0) {
        set_error("flashv_vanilla_decode: the decoded path runs through a state no transition reaches (the reference's vanilla "
                  "program reads T2[-1] there: outside its domain)");
        return FLASHV_ERR_DOMAIN;
    }
    if (score_out) *score_out = score;
    if (report) {
        memset(report, 0, sizeof(*report));
        report->decode_ms = ms;
        report->executed_steps = T - 1;
        report->kernel_launches = T + 1;
        report->memory_bytes = (int)(sizeof(float) * (size_t)K * T + sizeof(int) * (size_t)K * T);  // V:172 sizeof(T1)+sizeof(T2)
        report->device_bytes = (long long)(m->bytes + (size_t)2 * K * 4 + psi_rows * K * 4);
    }
    return FLASHV_OK;
}
