// pipes.cu — issue/pipe throughput of the max-plus inner loop's instructions on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
// Prints warp-instructions per clock per SM for several instruction mixes (16 independent
// chains per thread, 16 warps per SM resident so that latency is hidden).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
#define FADD(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMAX(d, a, b) asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))

constexpr int NCH = 16, ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, const float *in, long long *cyc)
{
    float c[NCH], x[NCH], t = in[threadIdx.x & 7], h = in[8 + (threadIdx.x & 7)];
#pragma unroll
    for (int i = 0; i < NCH; ++i) c[i] = in[i] - 1e30f, x[i] = in[16 + i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            float a, b;
            if (MODE == 0) { FADD(c[i], c[i], x[i]); }                                   // FADD only
            if (MODE == 1) { FMAX(c[i], c[i], x[i]); }                                   // FMNMX only
            if (MODE == 2) { c[i] = fmax3(c[i], x[i], t); }                              // FMNMX3 only
            if (MODE == 3) { FADD(a, x[i], t); FADD(b, a, h); FMAX(c[i], c[i], b); }     // 2 FADD + FMNMX  (3 per update)
            if (MODE == 4) { FADD(a, x[i], h); FMAX(c[i], c[i], a); }                    // FADD + FMNMX    (2 per update)
            if (MODE == 5) { FADD(a, x[i], t); FADD(b, x[i], h); c[i] = fmax3(c[i], a, b); }  // 2 FADD + FMNMX3 = two 2-op updates
            if (MODE == 6) { float a2, b2; FADD(a, x[i], t); FADD(b, a, h); FADD(a2, x[i], h); FADD(b2, a2, t); c[i] = fmax3(c[i], b, b2); }  // 4 FADD + FMNMX3 = two 3-op updates
            if (MODE == 7) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(x[i]), "f"(t)); }  // FFMA only
        }
        t += 1.0f;  // keep the loop from being hoisted
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ninstr, float *out, float *in, long long *cyc, int sms)
{
    k<MODE><<<sms, 512>>>(out, in, cyc);
    k<MODE><<<sms, 512>>>(out, in, cyc);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double winstr = 16.0 * ITERS * NCH * ninstr;  // warp-instructions per SM
    printf("%-44s %6.3f warp-instr/clk/SM  (%lld cycles)\n", name, winstr / (double)mx, mx);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out, *in;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    cudaMalloc(&in, sizeof(float) * 64);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    float hin[64];
    for (int i = 0; i < 64; ++i) hin[i] = -1.0f - i;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    run<0>("FADD", 1, out, in, cyc, sms);
    run<7>("FFMA", 1, out, in, cyc, sms);
    run<1>("FMNMX", 1, out, in, cyc, sms);
    run<2>("FMNMX3", 1, out, in, cyc, sms);
    run<3>("2 FADD + FMNMX (one 3-op update)", 3, out, in, cyc, sms);
    run<4>("FADD + FMNMX (one 2-op update)", 2, out, in, cyc, sms);
    run<5>("2 FADD + FMNMX3 (two 2-op updates)", 3, out, in, cyc, sms);
    run<6>("4 FADD + FMNMX3 (two 3-op updates)", 5, out, in, cyc, sms);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
