// Microbenchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2, 1..16 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed)
{
    float2 a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = make_float2(seed + i, seed - i);
    float2 b = make_float2(seed * 0.5f, seed * 0.25f), c = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int r = 0; r < 8; ++r) {
            #pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, c.x, b.x); a[i].y = fmaf(a[i].y, c.y, b.y); }
                else           { a[i] = __ffma2_rn(a[i], c, b); }
            }
        }
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
    const int iters = 4000;
    printf("SMs %d, clock attr %d kHz\n", sms, khz);
    for (int mode = 0; mode < 2; ++mode)
        for (int threads = 128; threads <= 1024; threads *= 2) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<sms, threads>>>(out, iters, 1.0f); else k<1><<<sms, threads>>>(out, iters, 1.0f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)sms * threads * iters * 8 * 16 * 2;   // scalar FMAs
            printf("%s threads/SM %4d: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", mode ? "FFMA2" : "FFMA ", threads, ms,
                   2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
        }
    return 0;
}
