// Streaming-store throughput of candidate drain patterns for the flat-chunk tensor-core kernel (round 2).
// A "stage" is 64 KB of output; a "user" is `stages_per_user` consecutive stages (cfg5: 8 -> 512 KB, cfg2: 64 -> 4 MB); users are
// dealt to 2 CTAs/SM by an atomic ticket like the real kernel; 4 store warps per CTA.
//   mode 0: round-1 pattern: stage = 64 rows x 1 KB at an 8 KB row pitch, warp = 128-byte quarter of each 512 B sub-tile (b32 lanes)
//   mode 1: flat stage (64 KB contiguous): warp q writes 128 B at  i*512 + q*128,  i = 0..127  (b32 lanes, direct from registers)
//   mode 2: flat stage staged through shared memory (4 x 16 KB pieces, double-buffered), one cp.async.bulk (16 KB) per piece
//   mode 3: flat stage, 16-byte lanes: warp w writes 512 B at (4 i + w) * 512
//   mode 4: like mode 2 with an L2 evict_first cache hint on the bulk copy
//   mode 5: like mode 2 with 32 KB pieces (2 per stage)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ unsigned int g_ticket;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128, 2) k(float* out, int n_users, int stages_per_user)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned int s_user;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    uint64_t policy = 0;
    if (MODE == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    unsigned piece_no = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_user = atomicAdd(&g_ticket, 1u);
        __syncthreads();
        const unsigned user = s_user;
        if (user >= (unsigned)n_users) break;
        float* ubase = out + (size_t)user * stages_per_user * 16384;
        for (int st = 0; st < stages_per_user; ++st) {
            float* base = ubase + (size_t)st * 16384;
            if (MODE == 0) {
                // 64 rows x 1 KB (two 512 B sub-tiles), row pitch = stages_per_user * 1 KB
                const long long pitch = (long long)stages_per_user * 256;
                float* b0 = ubase + st * 256;
                for (int sub = 0; sub < 2; ++sub) {
                    float* p = b0 + sub * 128 + q * 32 + lane;
                    for (int i = 0; i < 64; ++i) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory"); p += pitch; }
                }
            } else if (MODE == 1) {
                float* p = base + q * 32 + lane;
                #pragma unroll 8
                for (int i = 0; i < 128; ++i) { asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(1.f) : "memory"); p += 128; }
            } else if (MODE == 3) {
                float* p = base + q * 128 + lane * 4;
                #pragma unroll 8
                for (int i = 0; i < 32; ++i) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.f) : "memory"); p += 512; }
            } else {
                constexpr int kPiece = (MODE == 5) ? 32768 : 16384;
                constexpr int kCols = kPiece / 512;                   // 512-byte chunks per piece
                for (int pc = 0; pc < 65536 / kPiece; ++pc, ++piece_no) {
                    unsigned char* buf = smem + (piece_no & 1) * kPiece;
                    // the bulk copy that read this buffer two pieces ago must have finished reading shared memory
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    float* sp = reinterpret_cast<float*>(buf) + q * 32 + lane;
                    #pragma unroll 8
                    for (int i = 0; i < kCols; ++i) sp[i * 128] = 1.f + i;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (threadIdx.x == 0) {
                        float* dst = base + pc * (kPiece / 4);
                        if (MODE == 4)
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                                         :: "l"(dst), "r"(smem_u32(buf)), "r"(kPiece), "l"(policy) : "memory");
                        else
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                         :: "l"(dst), "r"(smem_u32(buf)), "r"(kPiece) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
    }
    if (MODE == 2 || MODE == 4 || MODE == 5) { if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}

template <int MODE>
float run(float* out, int n_users, int spu, size_t smem)
{
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    for (int it = 0; it < 3; ++it) {
        unsigned z = 0; cudaMemcpyToSymbol(g_ticket, &z, sizeof(z));
        cudaEventRecord(a);
        k<MODE><<<296, 128, smem>>>(out, n_users, spu);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return ms;
}

int main()
{
    const size_t total = 9ull << 30;                       // 9 GiB per run
    float* out; if (cudaMalloc(&out, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    const char* names[6] = {"r1: 64 rows x 1KB, quarter warps", "flat, b32 direct (128B x4 warps)", "flat, smem + bulk 16KB",
                            "flat, v4 direct (512B/warp)", "flat, smem + bulk 16KB evict_first", "flat, smem + bulk 32KB"};
    for (int spu : {8, 64}) {
        const int n_users = (int)(total / ((size_t)spu * 65536));
        printf("user = %d stages (%d KB), %d users\n", spu, spu * 64, n_users);
        for (int mode = 0; mode < 6; ++mode) {
            float ms = 0;
            switch (mode) {
                case 0: ms = run<0>(out, n_users, spu, 0); break;
                case 1: ms = run<1>(out, n_users, spu, 0); break;
                case 2: ms = run<2>(out, n_users, spu, 32768); break;
                case 3: ms = run<3>(out, n_users, spu, 0); break;
                case 4: ms = run<4>(out, n_users, spu, 32768); break;
                case 5: ms = run<5>(out, n_users, spu, 65536); break;
            }
            printf("  %-36s %.3f ms  %.0f GB/s\n", names[mode], ms, (double)n_users * spu * 65536.0 / ms * 1e-6);
        }
    }
    return 0;
}
