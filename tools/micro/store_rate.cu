// Streaming-store throughput: 4-byte-per-lane stores (one 128-byte line per warp instruction, rows `pitch` apart,
// like the transposed tensor-core epilogue) vs 16-byte-per-lane stores (512 bytes per instruction).
#include <cstdio>
#include <cuda_runtime.h>
template <int W>   // W = floats per lane per store: 1 or 4
__global__ void __launch_bounds__(256) k(float* out, long long pitch_floats, int rows_per_warp, int reps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * 8 + warp;
    for (int r = 0; r < reps; ++r) {
        float* p = out + ((gw * reps + r) * rows_per_warp) * pitch_floats;   // this warp's block of rows
        for (int i = 0; i < rows_per_warp; ++i) {
            if (W == 1) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + lane), "f"(1.f) : "memory"); p += pitch_floats; }
            else        { if (lane < 8 || true) asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p + 4 * lane), "f"(1.f) : "memory"); p += pitch_floats; }
        }
    }
}
int main()
{
    const long long pitch = 2048;            // floats: 8 KB rows like cfg5 (K = 1024)
    const int grid = 296, rows = 32, reps = 256;
    const size_t total_rows = (size_t)grid * 8 * reps * rows;
    float* out; cudaMalloc(&out, total_rows * pitch * 4 > (size_t)64 << 30 ? (size_t)64 << 30 : total_rows * pitch * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) {
        for (int it = 0; it < 2; ++it) {
            cudaEventRecord(a);
            if (w == 0) k<1><<<grid, 256>>>(out, pitch, rows, reps); else k<4><<<grid, 256>>>(out, pitch, rows, reps);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)total_rows * (w == 0 ? 128 : 512);
        printf("%s: %.3f ms, %.0f GB/s, %.1f B/clk/SM at 1.965 GHz, %.2f cyc per warp-store per SM\n", w == 0 ? "st.b32  (128 B/instr)" : "st.v4   (512 B/instr)",
               ms, bytes / ms / 1e6, bytes / (ms * 1e-3) / 148 / 1.965e9, (ms * 1e-3 * 1.965e9) / ((double)total_rows / 148));
    }
    return 0;
}
