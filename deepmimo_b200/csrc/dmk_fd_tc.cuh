// dmk_fd_tc.cuh -- FD channel contraction on the 5th-generation tensor cores (tcgen05, kind::f16, TMEM): the arithmetic, the
// operand layouts and the one-CTA-per-user kernel.  The production kernel (persistent, warp-specialised) is in dmk_fd_ws.cuh and
// reuses everything here except the kernel body.
//
// Why: with P ~ 12 contributing paths per user the FP32 FMA pipe, not HBM, bounds the CUDA-core kernel
// (profiles/r01_ncu_fd_fast_kernel_*.txt: top stall math_pipe_throttle, DRAM = algorithmic bytes).  The
// rank-P sum is a small-K complex GEMM per user,
//     H[m, k] = sum_p A[m,p] W[p,k]           A: M x P,  W: P x K   (complex)
// written as one real GEMM whose rows are already the interleaved complex64 layout of a 64-subcarrier segment:
//     Dt[2k+s, m] = sum_{p,e} B'[2k+s, 2p+e] A'[m, 2p+e]        (the tensor core computes the TRANSPOSED tile)
//     A'[m,2p] = Re A, A'[m,2p+1] = Im A;  B'[2k,2p] = Re W, B'[2k,2p+1] = -Im W, B'[2k+1,2p] = Im W, B'[2k+1,2p+1] = Re W.
// TMEM lanes = the 128 floats (re/im of 64 subcarriers) of an output row segment, TMEM columns = antenna rows m:
// register i of a tcgen05.ld holds, across the 32 lanes of a warp, 32 consecutive floats of output row m_i, so every
// warp-wide 4-byte store is one full 128-byte line -- no shared-memory staging -- and the MMA N dimension adapts to
// small arrays (M = 64 -> N = 64, M = 8 -> N = 16) without idle tensor rows.
// Precision: FP16 operands keep 11 significant bits, so every operand is split x = hi + lo (hi = fp16(x), lo = fp16(x - hi)) and
// D = B_hi A_hi + B_lo A_hi + B_hi A_lo accumulates in FP32 in tensor memory.  Operands are scaled into [-1, 1] (unit phasors;
// path gains divided by the user's largest |c_p| component, undone in the epilogue), so the absolute operand error is
// <= 2^-23 relative to the user's strongest path -- measured per-user relative Frobenius error <= 4.3e-7 (tests), far inside
// the 1e-5 budget.  All 32 path slots (re, im) of a row are one 128-byte K-major SWIZZLE_128B row; one MMA covers 8 slots (K = 16).
// (A first version used 3xTF32, K = 8 per MMA: same accuracy, twice the MMAs and operand bytes.)
//
// fd_tc_kernel (one CTA = one user or a slice of its column stages, 288 threads, 2 CTAs per SM):
//   warps 0-2: per-path float64 prologue chains -> warp 0 combines and compacts;  all threads: separable phasor tables
//   per stage (64 or 2 x 64 subcarriers) and row tile of mtile rows:
//     workers write A_hi/A_lo (mtile x 64 halves) and B_hi/B_lo (128 x 64 halves) straight into shared memory in the
//     K-major SWIZZLE_128B UMMA layout (one complex multiply of table entries + split per entry);
//     fence.proxy.async; barrier; warp 8 issues 3 x ksteps tcgen05.mma (128 x mtile x 16) and commits to an mbarrier.
//   epilogue: every worker warp pulls its TMEM lane quarter with tcgen05.ld 32x32b.xN and writes each register as one
//   128-byte row segment (streaming stores).  The accumulator is double-buffered in TMEM (2 x 128 columns): the
//   epilogue of stage s-1 runs while the tensor core works on stage s.
// Used for shapes whose double-buffered tables do not fit next to the operand tiles, and as DMK_FD_KERNEL=tc1.
#pragma once
#include <cuda_fp16.h>
#include "dmk_fd.cuh"

namespace dmk {

constexpr int kTcWorkers = 256;   // 8 warps build operands and drain accumulators
constexpr int kTcThreads = 288;   // + warp 8: issues the tcgen05.mma groups (never shares a warp with worker code)
constexpr int kTcN       = 128;   // real output columns per tile = 64 subcarriers
constexpr int kTcSlots   = 32;    // path slots per operand row: 64 fp16 (re, im per path) = one 128-byte swizzle row

struct TcCfg {
    int off_A, off_B, off_tY, off_tQ, off_wA, off_wB;   // byte offsets from the 1024-aligned base
    int nA, pcap, mtile;                                 // mtile = antenna rows per tile = tcgen05 N: 16, 32, 64 or 128
    int nsub;                                            // 64-subcarrier sub-tiles per pipeline stage: 2 when mtile <= 64, else 1
    int off_seed, wa_table;                              // seed tables [pcap][41]; wa_table = 1: coarse table wA materialised (K <= 1024)
    int off_tab, tab_bytes, sS;                          // persistent kernel: table buffer base, bytes per buffer, seed stride
    int sY, sQ, sA, sB;                                  // per-path table strides (float2 units), odd -> lanes that differ
                                                         // in the path index hit different shared-memory banks
    unsigned mul_mt, mul_bs0;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);      // start address >> 4
    d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

// x = hi + lo in FP16: hi = fp16(x), lo = fp16(x - hi).  Operands are scaled to |x| <= 1 (unit phasors; path gains divided
// by the user's largest |c_p|), so the absolute error per operand is <= max(2^-23 |x|, 2^-25) -- FP32-class relative to
// the user's strongest path, which is what the per-user Frobenius criterion measures.  Four path slots (re,im x 4 =
// 8 halves = 16 bytes) per store.
__device__ __forceinline__ void st_split8_f16(unsigned char* hi, unsigned char* lo, int off, const float2 (&a)[4])
{
    uint32_t h[4], l[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(a[i].x, a[i].y);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a[i].x - back.x, a[i].y - back.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();     // a lost commit must fail loudly, never hang the device
    }
}

// tcgen05.ld 32x32b: lane t of the warp reads TMEM lane (quarter base + t), NR consecutive columns.
template <int NR> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[32]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// Write NR antenna rows of one accumulator: register i = row (r_first + i), lanes = 32 consecutive floats of that row.
template <int NR>
__device__ __forceinline__ void tc_store_rows(uint32_t taddr, float* out_rows, long long pitch, int r_first, int M, float scale)
{
    uint32_t v[32];
    tmem_ld<NR>(taddr, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < NR; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * scale);      // undo the per-user operand scale
    // explicit 64-bit pointer bump per row (2 instructions): left to itself the compiler re-derives base + i*pitch
    const long long pitch_bytes = pitch * 4;
    if (r_first + NR <= M) {
        #pragma unroll
        for (int i = 0; i < NR; ++i) {
            asm volatile("st.global.cs.b32 [%0], %1;" :: "l"(out_rows), "r"(v[i]) : "memory");
            asm volatile("add.s64 %0, %0, %1;" : "+l"(out_rows) : "l"(pitch_bytes));
        }
    } else {
        #pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (r_first + i < M) asm volatile("st.global.cs.b32 [%0], %1;" :: "l"(out_rows), "r"(v[i]) : "memory");
            asm volatile("add.s64 %0, %0, %1;" : "+l"(out_rows) : "l"(pitch_bytes));
        }
    }
}

#ifdef DMK_TC_TRACE
__device__ long long g_tc_trace[4096];
#define TC_TRACE(slot) do { if (trace_on && (slot) < 4096) g_tc_trace[(slot)] = clock64(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#endif

struct TcTile { int row0, ct, acc; };   // ct: index of the 64-subcarrier segment, acc: first TMEM column of the accumulator

__device__ __noinline__ void tc_epilogue(const TcTile& t, uint32_t tmem_base, float* out_u, long long pitch, int M, int mtile,
                                            int warp, int lane, float scale)
{
    const int q = warp & 3, h = warp >> 2;                 // TMEM lane quarter (32 floats of the segment), half of the rows
    const int rows_half = mtile >> 1;                      // 64, 32, 16 or 8 antenna rows per warp
    const int r_base = t.row0 + h * rows_half;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t.acc + h * rows_half);
    float* o = out_u + (long long)r_base * pitch + t.ct * kTcN + q * 32 + lane;
    if (rows_half == 64) {
        tc_store_rows<32>(taddr, o, pitch, r_base, M, scale);
        tc_store_rows<32>(taddr + 32, o + 32 * pitch, pitch, r_base + 32, M, scale);
    } else if (rows_half == 32) {
        tc_store_rows<32>(taddr, o, pitch, r_base, M, scale);
    } else if (rows_half == 16) {
        tc_store_rows<16>(taddr, o, pitch, r_base, M, scale);
    } else {
        tc_store_rows<8>(taddr, o, pitch, r_base, M, scale);
    }
}

__global__ void __launch_bounds__(kTcThreads, 2)
fd_tc_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ TcCfg cfg, const int ksplit)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ FdShared sh;
    __shared__ PrologueScratch psc;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float scale_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_worker = tid < kTcWorkers;          // warp 8 only issues MMAs
    const long long user = blockIdx.x / ksplit;
    const int ks = blockIdx.x % ksplit;
    const int mtile = cfg.mtile, nsub = cfg.nsub;

    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* sAhi = sm + cfg.off_A;                    // [mtile rows][128 B]: 32 path slots x (re, im) fp16
    unsigned char* sAlo = sAhi + mtile * 128;
    unsigned char* sBhi = sm + cfg.off_B;                    // per sub-tile: [128 rows][128 B] hi, then lo
    unsigned char* sBlo = sBhi + kTcN * 128;
    float2* tY = reinterpret_cast<float2*>(sm + cfg.off_tY);
    float2* tQ = reinterpret_cast<float2*>(sm + cfg.off_tQ);
    float2* wA = reinterpret_cast<float2*>(sm + cfg.off_wA);
    float2* wB = reinterpret_cast<float2*>(sm + cfg.off_wB);

    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(2 * 128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 64) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#ifdef DMK_TC_TRACE
    const bool trace_user = (blockIdx.x >= gridDim.x / 2 && blockIdx.x < gridDim.x / 2 + 6) && tid == 0;
    const int tu = 4000 + 8 * (int)(blockIdx.x - gridDim.x / 2);
    if (trace_user) g_tc_trace[tu + 0] = clock64();
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    fd_cta_prologue_outlined(d, user, sh, psc, ks == 0);      // (two CTA barriers inside)
#ifdef DMK_TC_TRACE
    if (trace_user) { g_tc_trace[tu + 1] = clock64(); g_tc_trace[tu + 5] = sh.np; }
#endif
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int np = sh.np;
    const int K = d.K, M = d.M;
    const int n_seg = K / (kTcN / 2);                         // 64-subcarrier segments per row
    const int n_ct = (n_seg + nsub - 1) / nsub;               // pipeline stages (column super-tiles) per row tile
    const int n_rt = (M + mtile - 1) / mtile;
    float* out_u = reinterpret_cast<float*>(d.out + user * (long long)M * K);
    const long long pitch = 2LL * K;                          // floats per output row
    const int nq = d.Mr * d.bs1;

    if (np == 0) {
        // users without contributing paths: zeros (channel.py:257,:269-271), coalesced
        if (is_worker) {
            float4* o = reinterpret_cast<float4*>(out_u);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int c4 = tid & 31, r0 = tid >> 5;           // 32 float4 per row of a segment, 8 rows per pass
            for (int ct = ks; ct < n_ct; ct += ksplit)
                for (int sub = 0; sub < nsub && ct * nsub + sub < n_seg; ++sub) {
                    float4* ot = o + (ct * nsub + sub) * (kTcN / 4) + c4;
                    for (int m = r0; m < M; m += kTcWorkers / 32) __stcs(ot + (long long)m * (pitch / 4), z);
                }
        }
    } else {
        // ---- per-user operand scale: largest |c_p| component (FP16 operands live in [-1, 1])
        if (warp == 0) {
            float mx = (lane < np) ? fmaxf(fabsf(sh.c[lane].x), fabsf(sh.c[lane].y)) : 0.f;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) scale_s = mx;
        }
        __syncthreads();
        const float scale = scale_s;
        const float inv_scale = 1.0f / scale;
        // ---- per-user tables (phase reduced in float64 for every entry)
        {
            const int bs0 = d.bs0, bs1 = d.bs1;
            for (int e = tid; e < np * bs0; e += kTcThreads) {
                const int p = e / bs0, y = e - p * bs0;
                tY[p * cfg.sY + y] = phasor_cycles((double)y * sh.u[0][p]);
            }
            for (int e = tid; e < np * nq; e += kTcThreads) {
                const int p = e / nq, q = e - p * nq;
                const int r = q / bs1, z = q - r * bs1;
                const int yr = r % d.ue0, zr = r / d.ue0;
                const float2 cs = make_float2(sh.c[p].x * inv_scale, sh.c[p].y * inv_scale);
                tQ[p * cfg.sQ + q] = cmul(cs, phasor_cycles((double)z * sh.v[0][p] + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
            }
            // delay phasors: wB[b] (16 entries) and two seed tables per path (8 fine + up to 32 coarse entries) are reduced in
            // float64; the nA-entry table wA[a] = seed_hi[a >> 3] * seed_lo[a & 7] is then one complex multiply per entry
            // (unit-modulus float32 product) instead of nA sincos evaluations.  K <= 4096 (host-checked).
            float2* seed = reinterpret_cast<float2*>(sm + cfg.off_seed);  // [np][41]: 8 fine + up to 32 coarse entries
            const int n_hi = (cfg.nA + 7) >> 3;
            for (int e = tid; e < np * 16; e += kTcThreads) {
                const int p = e >> 4, b = e & 15;
                wB[p * cfg.sB + b] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * b)));
            }
            for (int e = tid; e < np * 40; e += kTcThreads) {
                const int p = e / 40, b = e - p * 40;
                if (b < 8)              seed[p * 41 + b] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * 16 * b)));                           // seed_lo[b]
                else if (b - 8 < n_hi)  seed[p * 41 + b] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_start + d.subc_step * 128 * (b - 8))));   // seed_hi[b - 8]
            }
        }
        __syncthreads();
        if (cfg.wa_table) {
            for (int e = tid; e < np * cfg.nA; e += kTcThreads) {
                const int p = e / cfg.nA, a = e - p * cfg.nA;
                const float2* sd = reinterpret_cast<const float2*>(sm + cfg.off_seed) + p * 41;
                wA[p * cfg.sA + a] = cmul(sd[8 + (a >> 3)], sd[a & 7]);
            }
            __syncthreads();
        }
#ifdef DMK_TC_TRACE
        if (trace_user) g_tc_trace[tu + 2] = clock64();
#endif
        const int ksteps = (np + 7) >> 3;                     // 16 fp16 (8 path slots) per MMA
        const int nslot = ksteps * 8;                         // slots the tensor core reads; slots >= np are written as zeros
        // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = antenna rows of the tile, M = 128 floats
        const uint32_t idesc = (1u << 4) | ((uint32_t)(mtile >> 3) << 17) | ((uint32_t)(kTcN >> 4) << 24);
        const uint64_t dAhi = umma_desc_kmajor_sw128(smem_u32(sAhi)), dAlo = umma_desc_kmajor_sw128(smem_u32(sAlo));
        const uint64_t dBhi = umma_desc_kmajor_sw128(smem_u32(sBhi)), dBlo = umma_desc_kmajor_sw128(smem_u32(sBlo));
        uint32_t phase = 0;
#ifdef DMK_TC_TRACE
        const bool trace_cta = (blockIdx.x == gridDim.x / 2);
        const bool trace_on = trace_cta && (tid == 0 || tid == kTcWorkers);
        const int tb = (tid == 0) ? 0 : 8;
        int stage_no = 0;
        if (trace_on && tid == 0) { g_tc_trace[4095] = np; }
#endif
        bool pending = false, have_prev = false;      // an MMA commit is outstanding / a stage waits for its epilogue
        TcTile prev = {0, 0, 0};
        int prev_nsub = 0;
        const int acc_stride = 128 / nsub;            // TMEM columns per accumulator; 2 * nsub accumulators in 256 columns
        int tile_idx = 0;

        // Operand builders: a thread owns one antenna row (A) / one subcarrier (B) and a run of consecutive path slots, so the
        // row -> (rx element, z, y) decomposition is done once per tile, table reads of neighbouring lanes are consecutive
        // or broadcast, and four path slots (8 halves) go out as one conflict-free 16-byte store.
        const int a_row  = tid & (mtile - 1);
        const int a_ngrp = kTcWorkers / mtile;                // thread groups per antenna row: 2, 4, 8 or 16
        const int a_grp  = tid / mtile;
        const int a_off0 = (a_row >> 3) * 1024 + (a_row & 7) * 128;
        const int b_col  = tid & 63;
        const int b_grp  = tid >> 6;                          // 0..3 -> slots 8*b_grp .. +7
        const int b_row0 = 2 * b_col;                         // rows 2c (Re H) and 2c + 1 (Im H)
        const int b_off0 = (b_row0 >> 3) * 1024 + (b_row0 & 7) * 128;

        bool a_valid = false;                         // A tile in smem is still current (single row tile)
        for (int ct = ks; ct < n_ct; ct += ksplit) {
            const int seg0 = ct * nsub;
            const int nsub_here = min(nsub, n_seg - seg0);
            bool b_valid = false;                     // B tiles of this column stage are in smem (reused by every row tile)
            for (int rt = 0; rt < n_rt; ++rt) {
                const int row0 = rt * mtile;
                TC_TRACE(16 * stage_no + tb + 0);
                if (pending) {                        // the previous MMA group has finished reading A/B (and writing its accumulator)
                    if (!is_worker) mbar_wait_parity(smem_u32(&mbar), phase);
                    asm volatile("bar.sync 1, %0;" :: "n"(kTcThreads) : "memory");
                    phase ^= 1;
                    pending = false;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                TC_TRACE(16 * stage_no + tb + 1);
                // ---- A_hi / A_lo: antenna rows x path slots
                if (is_worker && !(a_valid && n_rt == 1)) {
                    const int am = row0 + a_row;
                    const bool a_ok = am < M;
                    int a_q = 0, a_y = 0;
                    if (a_ok) {
                        const unsigned mm = (unsigned)am;
                        const unsigned rr = cfg.mul_mt ? __umulhi(mm, cfg.mul_mt) : mm;
                        const unsigned t = mm - rr * (unsigned)d.Mt;
                        const unsigned zt = cfg.mul_bs0 ? __umulhi(t, cfg.mul_bs0) : t;
                        a_y = (int)(t - zt * (unsigned)d.bs0);
                        a_q = (int)(rr * (unsigned)d.bs1 + zt);
                    }
                    // slot quads are dealt round-robin to the thread groups of a row so that every thread has work
                    // whatever the number of paths: quad = a_grp, a_grp + a_ngrp, ... < nslot / 4
                    #pragma unroll 1
                    for (int qd = a_grp; qd * 4 < nslot; qd += a_ngrp) {
                        float2 a[4];
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = qd * 4 + i;
                            a[i] = (a_ok && p < np) ? cmul(tQ[p * cfg.sQ + a_q], tY[p * cfg.sY + a_y]) : make_float2(0.f, 0.f);
                        }
                        st_split8_f16(sAhi, sAlo, a_off0 + (((qd ^ (a_row & 7)) & 7) << 4), a);
                    }
                }
                // ---- B_hi / B_lo per sub-tile (rows 2c -> Re H, 2c+1 -> Im H)
                if (is_worker && !b_valid) {
                    // (sub-tile, slot quad) pairs dealt round-robin to the four thread groups of a subcarrier
                    const int nq4 = nslot >> 2;
                    #pragma unroll 1
                    for (int pi = b_grp; pi < nsub_here * nq4; pi += 4) {
                        const int sub = pi / nq4, qd = pi - sub * nq4;
                        unsigned char* sBh = sBhi + sub * (2 * kTcN * 128);
                        unsigned char* sBl = sBh + kTcN * 128;
                        const int col = (seg0 + sub) * (kTcN / 2) + b_col;
                        float2 re[4], im[4];
                        #pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = qd * 4 + i;
                            float2 w = make_float2(0.f, 0.f);
                            if (p < np) {
                                const int a = col >> 4;
                                const float2* sd = reinterpret_cast<const float2*>(sm + cfg.off_seed) + p * 41;
                                const float2 wa = cfg.wa_table ? wA[p * cfg.sA + a] : cmul(sd[8 + (a >> 3)], sd[a & 7]);
                                w = cmul(wa, wB[p * cfg.sB + (col & 15)]);
                            }
                            re[i] = make_float2(w.x, -w.y);
                            im[i] = make_float2(w.y, w.x);
                        }
                        st_split8_f16(sBh, sBl, b_off0 + (((qd ^ (b_row0 & 7)) & 7) << 4), re);
                        st_split8_f16(sBh, sBl, b_off0 + 128 + (((qd ^ ((b_row0 + 1) & 7)) & 7) << 4), im);
                    }
                }
                TC_TRACE(16 * stage_no + tb + 2);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                TC_TRACE(16 * stage_no + tb + 3);
                if (tid == kTcWorkers) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    #pragma unroll 1
                    for (int sub = 0; sub < nsub_here; ++sub) {
                        const uint32_t acc_col = (uint32_t)((((tile_idx & 1) * nsub) + sub) * acc_stride);
                        const uint64_t sub_off = (uint64_t)(sub * (2 * kTcN * 128) >> 4);       // descriptor address field: 16-byte units
                        #pragma unroll 1
                        for (int s = 0; s < 3; ++s) {                                           // hi*hi, lo*hi, hi*lo
                            const uint64_t da = (s == 2) ? dAlo : dAhi;
                            const uint64_t db = ((s == 1) ? dBlo : dBhi) + sub_off;
                            #pragma unroll 1
                            for (int kk = 0; kk < ksteps; ++kk) {
                                const uint32_t accum = (s | kk) != 0;
                                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                             :: "r"(tmem_base + acc_col), "l"(db + 2 * kk), "l"(da + 2 * kk), "r"(idesc), "r"(accum) : "memory");
                            }
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 :: "r"(smem_u32(&mbar)) : "memory");
                }
                pending = true;
                if (!is_worker) TC_TRACE(16 * stage_no + tb + 4);
                if (have_prev && is_worker)           // drain the previous stage while the tensor core works on this one
                    for (int sub = 0; sub < prev_nsub; ++sub) {
                        TcTile t = {prev.row0, prev.ct + sub, prev.acc + sub * acc_stride};
                        tc_epilogue(t, tmem_base, out_u, pitch, M, mtile, warp, lane, scale);
                    }
                if (is_worker) TC_TRACE(16 * stage_no + tb + 4);
#ifdef DMK_TC_TRACE
                ++stage_no;
#endif
                b_valid = true;
                a_valid = true;
                prev.row0 = row0; prev.ct = seg0; prev.acc = (tile_idx & 1) * nsub * acc_stride; prev_nsub = nsub_here;
                have_prev = true;
                ++tile_idx;
            }
        }
        if (pending) {
            if (!is_worker) mbar_wait_parity(smem_u32(&mbar), phase);
            asm volatile("bar.sync 1, %0;" :: "n"(kTcThreads) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (have_prev && is_worker)
            for (int sub = 0; sub < prev_nsub; ++sub) {
                TcTile t = {prev.row0, prev.ct + sub, prev.acc + sub * acc_stride};
                tc_epilogue(t, tmem_base, out_u, pitch, M, mtile, warp, lane, scale);
            }
    }

#ifdef DMK_TC_TRACE
    if (trace_user) g_tc_trace[tu + 3] = clock64();
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(2 * 128));
}


// Shared with the warp-specialised persistent kernel (dmk_fd_ws.cuh): device-side ticket counters and the per-user record.
constexpr int kTcTickets = 512;      // > resident CTA slots (2 x 148) and > the 128 concurrent grids of a device: launches in flight never share a slot
__device__ unsigned int g_tc_ticket[kTcTickets];   // zero between launches: the CTA that draws the last ticket resets it

struct TcUserBuf {
    FdShared sh;
    float scale;
    unsigned int item;
    unsigned char m_fov[kMaxPaths], m_valid[kMaxPaths], m_clip[kMaxPaths];   // mask bytes of the user: written to global memory by the drain warps
};

}  // namespace dmk
