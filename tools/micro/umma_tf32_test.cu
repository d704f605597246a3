// Stand-alone check of the hand-written tcgen05 (kind::tf32) path used by the FD tensor-core kernel:
// K-major SWIZZLE_128B shared-memory operand layout written by threads, smem/instruction descriptors,
// TMEM allocation, MMA issue + commit to an mbarrier, TMEM read-back (32x32b.x32) for M = 128 and M = 64.
// Every wait is bounded (no hang on a wrong descriptor).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);      // start address >> 4
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int max_spins)
{
    for (int i = 0; i < max_spins; ++i) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

template <int M, int N>
__global__ void __launch_bounds__(128) umma_test(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                 int ksteps, int* err)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sA = sm;                  // [M rows][32 tf32], 128 B per row, swizzled
    uint8_t* sB = sm + 128 * 128;      // [N rows][32 tf32]

    for (int e = tid; e < M * 32; e += 128) {
        const int r = e >> 5, k = e & 31;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(sA + off) = A[e];
    }
    for (int e = tid; e < N * 32; e += 128) {
        const int r = e >> 5, k = e & 31;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(sB + off) = B[e];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t da = make_kmajor_sw128_desc(smem_u32(sA)), db = make_kmajor_sw128_desc(smem_u32(sB));
        for (int j = 0; j < ksteps; ++j) {
            const uint32_t accum = j > 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                         :: "r"(tmem_base), "l"(da + 2 * j), "l"(db + 2 * j), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
    }
    const bool ok = mbar_wait(smem_u32(&mbar), 0, 1 << 22);
    if (!ok && lane == 0) atomicAdd(err, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (ok) {
        for (int c = 0; c < N / 32; ++c) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // M = 128: TMEM lane = row.  M = 64: row m lives in lane (m % 16) + 32 * (m / 16).
            int row = -1;
            if (M == 128) row = warp * 32 + lane;
            else if (lane < 16) row = warp * 16 + lane;
            if (row >= 0)
                for (int i = 0; i < 32; ++i) D[row * N + c * 32 + i] = __uint_as_float(v[i]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(N));
}

template <int M, int N>
__global__ void __launch_bounds__(128) umma_time(int nrep, long long* t_out)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    for (int e = tid; e < (128 + N) * 32; e += 128) reinterpret_cast<float*>(sm)[e] = 0.001f * (e & 255);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t da = make_kmajor_sw128_desc(smem_u32(sm)), db = make_kmajor_sw128_desc(smem_u32(sm + 128 * 128));
        const long long t0 = clock64();
        for (int r = 0; r < nrep; ++r)
            for (int j = 0; j < 4; ++j) {
                const uint32_t accum = (r | j) != 0;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                             :: "r"(tmem_base), "l"(da + 2 * j), "l"(db + 2 * j), "r"(idesc), "r"(accum) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar)) : "memory");
        const long long t1 = clock64();
        mbar_wait(smem_u32(&mbar), 0, 1 << 24);
        const long long t2 = clock64();
        t_out[0] = t1 - t0; t_out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(N));
}

template <int M, int N>
static void run_time()
{
    long long* d; cudaMalloc(&d, 16);
    const int smem = 128 * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(umma_time<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nrep : {1, 3, 12, 48}) {
        long long h[2] = {0, 0};
        for (int it = 0; it < 2; ++it) { umma_time<M, N><<<1, 128, smem>>>(nrep, d); cudaDeviceSynchronize(); }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("tf32 SS MMA M=%d N=%d K=8: %3d MMAs: issue %6lld cyc, complete %6lld cyc  (%.1f cyc/MMA)\n", M, N, 4 * nrep, h[0], h[1], (double)h[1] / (4 * nrep));
    }
    cudaFree(d);
}

static float to_tf32(float x)
{
    uint32_t u; memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xffffe000u;      // round to nearest (ties away), 10-bit mantissa
    memcpy(&x, &u, 4);
    return x;
}

template <int M, int N>
static int run(int ksteps)
{
    std::vector<float> A(128 * 32, 0.f), B(N * 32, 0.f), D(128 * N, -777.f);
    srand(7 + M + ksteps);
    for (int i = 0; i < M * 32; ++i) A[i] = to_tf32((rand() / (float)RAND_MAX) * 2 - 1);
    for (int i = 0; i < N * 32; ++i) B[i] = to_tf32((rand() / (float)RAND_MAX) * 2 - 1);
    float *dA, *dB, *dD; int* derr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(derr, 0, 4);
    const int smem = 128 * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(umma_test<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_test<M, N><<<1, 128, smem>>>(dA, dB, dD, ksteps, derr);
    cudaError_t e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 8 * ksteps; ++k) ref += (double)A[m * 32 + k] * B[n * 32 + k];
            double d = fabs(ref - D[m * N + n]);
            if (d > maxerr) maxerr = d;
            if (d > 1e-4) ++bad;
        }
    printf("M=%d N=%d ksteps=%d: cuda=%s timeout_warps=%d max|err|=%.3e bad=%d  D[0][0]=%f D[1][5]=%f\n", M, N, ksteps,
           cudaGetErrorString(e), herr, maxerr, bad, D[0], D[N + 5]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(derr);
    return (e != cudaSuccess) || herr || bad;
}

int main()
{
    int rc = 0;
    rc |= run<128, 128>(4);
    rc |= run<128, 128>(3);
    rc |= run<128, 256>(4);
    rc |= run<64, 128>(4);
    rc |= run<64, 128>(1);
    run_time<128, 128>(); run_time<128, 64>(); run_time<128, 256>(); run_time<64, 128>();
    printf(rc ? "UMMA TEST FAILED\n" : "UMMA TEST PASSED\n");
    return rc;
}
