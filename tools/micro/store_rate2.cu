// Streaming-store throughput of the tensor-core epilogue's pattern as a function of resident warps and tile walk order.
//   mode 0: st.b32, warp = (quarter q, row half h), 128 rows x 512 B tile, tiles walk along the row (fd_tc mtile=128)
//   mode 1: st.b32, 64 rows x 1 KB stage (two 512 B sub-tiles), warp = (q, h): 32 rows per warp per sub-tile (fd_tc mtile=64, nsub=2)
//   mode 2: st.b32, like mode 1 but each warp walks the 8 x 128 B pieces of its rows' 1 KB before moving to the next row
//   mode 3: st.v4 (512 B per instruction), warp w owns rows w*8.., both sub-tiles (1 KB per row, two instructions)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long pitch, int reps, int tiles_per_row)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp & 3, h = warp >> 2;
    for (int r = 0; r < reps; ++r) {
        const long long tile = (long long)blockIdx.x * reps + r;         // 64 KB of output each
        if (MODE == 0) {
            float* base = out + (tile / tiles_per_row) * 128 * pitch + (tile % tiles_per_row) * 128;
            float* p = base + (long long)(h * 64) * pitch + q * 32 + lane;
            for (int i = 0; i < 64; ++i) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory"); p += pitch; }
        } else {
            const int spr = tiles_per_row / 2;                           // 1 KB stages per row
            float* base = out + (tile / spr) * 64 * pitch + (tile % spr) * 256;
            if (MODE == 1) {
                for (int sub = 0; sub < 2; ++sub) {
                    float* p = base + (long long)(h * 32) * pitch + sub * 128 + q * 32 + lane;
                    for (int i = 0; i < 32; ++i) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory"); p += pitch; }
                }
            } else if (MODE == 4) {        // 4 warps (q) x 2 row halves (h): per row 256 B contiguous per warp (two adjacent 128 B stores)
                float* p = base + (long long)(h * 32) * pitch + q * 64 + lane;
                for (int i = 0; i < 32; ++i) {
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory");
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + 32), "f"(1.f) : "memory");
                    p += pitch;
                }
            } else if (MODE == 5) {        // like mode 1 but the two sub-tiles interleaved per row (128 B pieces 512 B apart)
                float* p = base + (long long)(h * 32) * pitch + q * 32 + lane;
                for (int i = 0; i < 32; ++i) {
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory");
                    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + 128), "f"(1.f) : "memory");
                    p += pitch;
                }
            } else if (MODE == 2) {
                float* p = base + (long long)(warp * 8) * pitch + lane;
                for (int i = 0; i < 8; ++i) {
                    for (int s = 0; s < 8; ++s) asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p + s * 32), "f"(1.f) : "memory");
                    p += pitch;
                }
            } else {
                float* p = base + (long long)(warp * 8) * pitch + 4 * lane;
                for (int i = 0; i < 8; ++i) {
                    asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.f) : "memory");
                    asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p + 128), "f"(1.f) : "memory");
                    p += pitch;
                }
            }
        }
    }
}
int main()
{
    const long long pitch = 2048;            // floats: 8 KB rows like cfg5 (K = 1024)
    const int tiles_per_row = 16;
    const long long total_tiles = 296LL * 512;     // 9.7 GB
    float* out; if (cudaMalloc(&out, (size_t)(total_tiles / 16 + 2) * 128 * pitch * 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[6] = {"b32 128x512B tile", "b32 64x1KB stage (q,h)", "b32 64x1KB row-walk", "v4  64x1KB rows", "b32 64x1KB 256B/warp/row", "b32 64x1KB sub-interleaved"};
    for (int cps = 2; cps <= 2; ++cps) {
        const int grid = 148 * cps, reps = (int)(total_tiles / grid);
        for (int mode = 0; mode < 6; ++mode) {
            float ms = 0;
            for (int it = 0; it < 2; ++it) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                else if (mode == 1) k<1><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                else if (mode == 2) k<2><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                else if (mode == 3) k<3><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                else if (mode == 4) k<4><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                else k<5><<<grid, 256>>>(out, pitch, reps, tiles_per_row);
                cudaEventRecord(b); cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
            const double bytes = (double)grid * reps * 65536.0;
            printf("CTAs/SM %d  %-26s %.3f ms  %.0f GB/s\n", cps, names[mode], ms, bytes / ms * 1e-6);
        }
    }
    return 0;
}
