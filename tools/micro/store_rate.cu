// Streaming-store throughput: 4-byte-per-lane stores (one 128-byte line per warp instruction, rows `pitch` apart,
// like the transposed tensor-core epilogue) vs 16-byte-per-lane stores (512 bytes per instruction).
#include <cstdio>
#include <cuda_runtime.h>
template <int W>   // W = floats per lane per store: 1 or 4
__global__ void __launch_bounds__(256) k(float* out, long long pitch_floats, int rows_per_warp, int reps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // CTA tile like the tensor-core epilogue: 128 rows x 128 floats (512 B per row); tiles of a CTA walk along the row
    // (16 tiles = one 8 KB row block), then down.  W == 1: warp = (quarter q, half h): rows h*64.., 128-byte segment q.
    // W == 4: warp w writes rows w*16.., the whole 512-byte segment per instruction.
    const int q = warp & 3, h = warp >> 2;
    for (int r = 0; r < reps; ++r) {
        const long long tile = (long long)blockIdx.x * reps + r;
        float* base = out + (tile / 16) * 128 * pitch_floats + (tile % 16) * 128;
        if (W == 1) {
            float* p = base + (long long)(h * 64) * pitch_floats + q * 32 + lane;
            for (int i = 0; i < 64; ++i) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory"); p += pitch_floats; }
        } else {
            float* p = base + (long long)(warp * 16) * pitch_floats + 4 * lane;
            for (int i = 0; i < 16; ++i) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.f) : "memory"); p += pitch_floats; }
        }
    }
}
int main()
{
    const long long pitch = 2048;            // floats: 8 KB rows like cfg5 (K = 1024)
    const int grid = 296, rows = 32, reps = 512;     // 296 x 512 tiles of 64 KB = 9.7 GB
    const size_t total_rows = ((size_t)grid * reps / 16 + 1) * 128;
    float* out; if (cudaMalloc(&out, total_rows * pitch * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 2; ++w) {
        for (int it = 0; it < 2; ++it) {
            cudaEventRecord(a);
            if (w == 0) k<1><<<grid, 256>>>(out, pitch, rows, reps); else k<4><<<grid, 256>>>(out, pitch, rows, reps);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b); if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
        const double bytes = (double)grid * reps * 65536.0;
        printf("%s: %.3f ms, %.0f GB/s, %.1f B/clk/SM at 1.965 GHz, %.2f cyc per warp-store per SM\n", w == 0 ? "st.b32  (128 B/instr)" : "st.v4   (512 B/instr)",
               ms, bytes / ms / 1e6, bytes / (ms * 1e-3) / 148 / 1.965e9, (ms * 1e-3 * 1.965e9) / (bytes / (w == 0 ? 128 : 512) / 148));
    }
    return 0;
}
