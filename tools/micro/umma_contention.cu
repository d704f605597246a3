// How slow do SS-mode tcgen05 MMAs get when other warps hammer shared memory / when two CTAs share the SM?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// mode bit0: warps 1..7 do smem traffic (STS.128 + LDS.128) while the MMAs run; kind: 0 = f16 (K=16), 1 = tf32 (K=8)
template <int KIND>
__global__ void __launch_bounds__(256) k(int nmma, int mode, long long* t_out, float* gbuf)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar; __shared__ uint32_t tb; __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    for (int e = tid; e < 16384; e += 256) reinterpret_cast<uint32_t*>(sm)[e] = 0x3c003c00u;   // 64 KB operand area
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb)), "r"(128));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    if (tid == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (KIND ? ((2u << 7) | (2u << 10)) : 0u) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = desc(smem_u32(sm)), db = desc(smem_u32(sm + 16384));
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const uint32_t acc = i != 0;
            if (KIND) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" :: "r"(tb), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(idesc), "r"(acc) : "memory");
            else      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" :: "r"(tb), "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        const long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        const long long t2 = clock64();
        stop = 1;
        if (blockIdx.x == gridDim.x / 2) { t_out[0] = t1 - t0; t_out[1] = t2 - t0; }
    } else if ((mode & 1) && warp >= 1) {
        uint4* area = reinterpret_cast<uint4*>(sm + 32768) + tid;       // separate 32 KB scratch area
        uint4 v = make_uint4(tid, 1, 2, 3);
        while (!stop) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) { area[i * 256 & 1023] = v; v.x += area[(i * 256 + 512) & 1023].y; }
        }
        if (v.x == 0x12345) t_out[3] = v.x;
    } else if ((mode & 2) && warp >= 1) {
        // global streaming stores: one 128-byte line per warp instruction, rows 8 KB apart (like the channel epilogue)
        float* g = gbuf + ((size_t)blockIdx.x * 8 + warp) * (1 << 20) + (tid & 31);
        int i = 0;
        while (!stop) {
            #pragma unroll
            for (int r = 0; r < 16; ++r) asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(g + ((i + r) & 255) * 2048), "f"(1.0f) : "memory");
            i += 16;
        }
    } else if ((mode & 4) && warp >= 4) {
        // TMEM loads of the other half of the allocation (columns 64..127 are never written by the MMA here: N = 64 test)
        uint32_t acc = 0;
        while (!stop) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tb + ((uint32_t)((warp & 3) * 32) << 16) + 64));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0];
        }
        if (acc == 0x12345) t_out[3] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(128));
}
int main()
{
    long long* d; cudaMalloc(&d, 64);
    float* gbuf; cudaMalloc(&gbuf, (size_t)296 * 8 * (1 << 20) * 4);
    const int smem = 65536 + 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int kind = 0; kind < 1; ++kind)
        for (int grid : {1, 296})
            for (int mode : {0, 2, 4, 6})
                for (int n : {6, 12}) {
                    long long h[2] = {0, 0};
                    for (int it = 0; it < 2; ++it) { if (kind) k<1><<<grid, 256, smem>>>(n, mode, d, gbuf); else k<0><<<grid, 256, smem>>>(n, mode, d, gbuf); cudaDeviceSynchronize(); }
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("%s 128x128 grid %3d contention(2=STG,4=LDTM) %d: %2d MMAs issue %6lld complete %6lld  (%.0f cyc/MMA)\n", kind ? "tf32" : "f16 ", grid, mode, n, h[0], h[1], (double)h[1] / n);
                }
    return 0;
}
