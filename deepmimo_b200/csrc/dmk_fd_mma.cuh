// dmk_fd_mma.cuh -- FD channel kernel for small per-user outputs (a few KB to ~100 KB): warp-level tensor-core accumulate.
//
// fd_small2_kernel (dmk_fd_small.cuh) spends 880 of its 1 670 warp-instructions per user on the CUDA-core accumulate
// (profiles/r02_ncu_fd_small2_cfg1.txt): 512 complex MACs per path and user is the floor of that formulation.  The persistent
// tcgen05 kernel (dmk_fd_ws.cuh) cannot take these users either: its per-user operand (128 rows of fine phasors) costs more shared
// memory traffic than a 4 KB user has output, and below ~400 KB per user it is bound by its helper warps.  Here the accumulate is a
// warp-level mma.sync (m16n8k16, FP16 hi/lo split, FP32 accumulation) fed from registers:
//
//   a user's output is a flat array of R = M * K / J chunks of J subcarriers (J = 16 or 32; chunk r = m * S + seg, S = K / J);
//   H[chunk r, j] = sum_p L[r, p] * F[p, j],   L[r, p] = c_p * a[m, p] * exp(-j 2 pi wcyc_p k0(seg)),   F[p, j] = exp(-j 2 pi wcyc_p step j)
//   real form:  Re H[r, j] = sum_p Lr Fr + Li (-Fi),   Im H[r, j] = sum_p Lr Fi + Li Fr:   A = (Lr, Li) per (chunk, path); an n-tile is
//   8 columns j of a chunk, once with B = (Fr, -Fi) (real parts) and once with B = (Fi, Fr) (imaginary parts) -- the second is the
//   first with its halves swapped and one sign flipped, formed in registers, so the pool holds every F once
//   -> MMA rows = chunks, the two accumulators of a column group interleave to a contiguous piece of the output,
//      MMA k = (path, re/im): 8 paths per k-step.
//
// Phases of a pass (one warp, no CTA-wide synchronisation; 1-3 are those of fd_small2_kernel):
//   1. window: power rows of the next 4 users -> bit masks of the columns whose chain must run;
//   2. whole users are taken while their (user, column) pairs fit ONE round of 32 lanes;
//   3. chain round, lane = dense pair: the three float64 chains + combine; the user's largest |c_p| (a power of two, by a masked
//      warp reduction over the lanes of the same user) scales the operands into FP16 range and is undone at the store;
//   4. operand rows, still lane = path: F (J columns: products of two phasor levels) and, per group of G m-tiles, L (blocks of SB
//      consecutive chunks -- of one antenna row, or of SB / S consecutive elements along y when a row has S < SB chunks -- share a base phasor -- gain x steering x coarse delay in ONE float64-reduced argument --
//      times SB block phasors), each split into FP16 hi + lo and stored in fragment order, path index fastest;
//   5. per user and m-tile: fragments by 16- and 8-byte shared loads that land in consecutive registers (a lane's two paths of a
//      k-step are adjacent pool slots), 3 MMAs (hi hi + lo hi + hi lo) per (n-tile, k-step), 16-byte stores.
// Pool rows are XOR-swizzled (32 slots of 8 or 4 bytes are exactly 256 / 128 bytes: no padding) so that producer stores and fragment
// loads are both conflict-free.  The pool is zeroed once: slots past a user's last path in its last k-step then hold finite stale
// values on the B side, and the A side is zeroed in registers.
#pragma once
#include <cuda_fp16.h>
#include "dmk_fd.cuh"

namespace dmk {

constexpr int kMmWarps  = 3;                       // 8.4 - 12.4 KB of pool per warp (J = 16): 6 CTAs = 18 warps per SM at 96 registers
constexpr int kMmWindow = 4;                       // users examined per pass
constexpr int kMmSlots  = 32;                      // pool slots (paths) per pass: one chain round

struct MmaCfg {
    int warp_bytes;                   // bytes of a warp's shared-memory slice
    int off_warps;                    // bytes of the CTA-wide tables in front of the warp slices
    int off_A, off_B, off_list, off_meta;
    int G;                            // m-tiles (16 chunks each) whose L rows are resident at a time
    int n_mt;                         // m-tiles per user = ceil(R / 16)
    int S, R;                         // chunks per antenna row (ceil(K / J)), chunks per user
    int ragged;                       // K % J != 0
    int lg_blk_seg;                   // log2 of the segments a block of SB chunks spans: SB when S % SB == 0 (one antenna row), else S (a block
                                      // then holds SB / S consecutive elements along the panel's y axis -- 16 or 32 subcarriers: S = 1 or 2)
    int users_per_warp;               // 0: guided draws (8 / 4 / 2 users); > 0 pins the draw size
    unsigned draw8_above, draw4_above; // users left in the launch above which a warp draws 8 / 4 users
    unsigned mul_s;                   // ceil(2^32 / S) (S > 1): chunk -> antenna row by a multiply-high
};

__device__ __forceinline__ void mma_m16n8k16_f16(float (&d)[4], const uint4& a, const uint2& b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}

// (x, y) -> FP16 pairs hi = (f16(x), f16(y)) and lo = (f16(x - hi.x), f16(y - hi.y)); x in the low half.
__device__ __forceinline__ void split_f16x2(float x, float y, unsigned& hi, unsigned& lo)
{
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const unsigned*>(&h);
    lo = *reinterpret_cast<const unsigned*>(&l);
}

// Pool layout of a warp (bytes).  A side (L): [hi | lo][G * 8 row pairs (m-tile, g)][32 slots] 8 B = chunks g and g + 8 of the slot's
// path, row = 256 B, 16-byte unit index XOR-ed with (g & 1) << 2.  B side (F): [hi | lo][NT / 2 column groups][8 columns][32 slots]
// 4 B = (Fr, -Fi) of the slot's path, row = 128 B, byte offset XOR-ed with (column & 3) << 5.
__device__ __forceinline__ int mm_a_off(int rowpair, int slot) { return rowpair * 256 + ((slot * 8) ^ ((rowpair & 1) << 6)); }
__device__ __forceinline__ int mm_b_off(int row, int slot)     { return row * 128 + ((slot * 4) ^ ((row & 3) << 5)); }

// K not a multiple of the chunk width: the chunks of an antenna row are S = ceil(K / J), the last one is stored up to column K only.
// Compiled into the generic instantiations only (carrying it costs the plain ones 4 %): such shapes take the generic family.
struct MmRagged { int on, K, S, vec2; unsigned mul_s; };

// Phase 5 for MP consecutive m-tiles of one user: they share the B fragments of a k-step (loaded once, imaginary-part operand formed
// once).  k-step = 8 pool slots; this lane's paths are slots 2t, 2t + 1 (MMA k indices (2t, 2t + 1) = (re, im) of the first,
// (2t + 8, 2t + 9) of the second).
template <int NT, int MP, bool kRag>
__device__ __forceinline__ void mm_consume(const unsigned char* sAh, const unsigned char* sAl, const unsigned char* bh_row, const unsigned char* bl_row,
                                           int rowpair0, unsigned a_sw, unsigned b_sw, int qb, int np, int t, float sc_up,
                                           float2* out_u, int r0, int R, const MmRagged& rg)
{
    constexpr int J = 4 * NT;
    float acc[MP][NT][4];
    #pragma unroll
    for (int i = 0; i < MP; ++i)
        #pragma unroll
        for (int n = 0; n < NT; ++n) { acc[i][n][0] = 0.f; acc[i][n][1] = 0.f; acc[i][n][2] = 0.f; acc[i][n][3] = 0.f; }
    const unsigned char* ah_row = sAh + rowpair0 * 256;
    const unsigned char* al_row = sAl + rowpair0 * 256;
    #pragma unroll 1
    for (int k0 = 0; k0 < np; k0 += 8) {
        const unsigned sl = (unsigned)min(qb + k0, kMmSlots - 2);       // past the row's end only when both paths are >= np (zeroed below)
        const unsigned ao = (sl * 8u) ^ a_sw, bo = (sl * 4u) ^ b_sw;
        uint2 bh[NT], bl[NT];                                        // [2 q]: (Fr, -Fi) of column group q, [2 q + 1]: (Fi, Fr)
        #pragma unroll
        for (int q = 0; q < NT / 2; ++q) {
            bh[2 * q] = *reinterpret_cast<const uint2*>(bh_row + q * 1024 + bo);
            bl[2 * q] = *reinterpret_cast<const uint2*>(bl_row + q * 1024 + bo);
            bh[2 * q + 1] = make_uint2(__byte_perm(bh[2 * q].x, 0, 0x1032) ^ 0x8000u, __byte_perm(bh[2 * q].y, 0, 0x1032) ^ 0x8000u);
            bl[2 * q + 1] = make_uint2(__byte_perm(bl[2 * q].x, 0, 0x1032) ^ 0x8000u, __byte_perm(bl[2 * q].y, 0, 0x1032) ^ 0x8000u);
        }
        const bool z0 = k0 + 2 * t >= np, z1 = k0 + 2 * t + 1 >= np;     // the user's last, partial k-step
        uint4 ah[MP], al[MP];
        #pragma unroll
        for (int i = 0; i < MP; ++i) {
            ah[i] = *reinterpret_cast<const uint4*>(ah_row + i * 2048 + ao);           // chunks g, g + 8 of path 2t; of path 2t + 1
            al[i] = *reinterpret_cast<const uint4*>(al_row + i * 2048 + ao);
            if (k0 + 8 > np) {                                           // warp-uniform
                if (z0) { ah[i].x = 0u; ah[i].y = 0u; al[i].x = 0u; al[i].y = 0u; }
                if (z1) { ah[i].z = 0u; ah[i].w = 0u; al[i].z = 0u; al[i].w = 0u; }
            }
        }
        // MP * NT independent accumulators between two MMAs into the same one
        #pragma unroll
        for (int i = 0; i < MP; ++i)
            #pragma unroll
            for (int n = 0; n < NT; ++n) mma_m16n8k16_f16(acc[i][n], ah[i], bh[n]);
        #pragma unroll
        for (int i = 0; i < MP; ++i)
            #pragma unroll
            for (int n = 0; n < NT; ++n) mma_m16n8k16_f16(acc[i][n], al[i], bh[n]);
        #pragma unroll
        for (int i = 0; i < MP; ++i)
            #pragma unroll
            for (int n = 0; n < NT; ++n) mma_m16n8k16_f16(acc[i][n], ah[i], bl[n]);
    }
    // chunk r = 16 mt + g (+ 8): J complex values at r * J; this lane holds columns 8 q + 2 t, + 1 of both chunks
    #pragma unroll
    for (int i = 0; i < MP; ++i)
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + 16 * i + 8 * h;
            if (r >= R) continue;
            if (!kRag || !rg.on) {                                      // K a multiple of J: chunk r starts at r * J
                float4* o4 = reinterpret_cast<float4*>(out_u + (size_t)r * J);
                #pragma unroll
                for (int q = 0; q < NT / 2; ++q)                        // columns 8 q + 2 t, + 1: (re, im, re, im)
                    __stcs(o4 + q * 4, make_float4(acc[i][2 * q][2 * h] * sc_up, acc[i][2 * q + 1][2 * h] * sc_up,
                                                   acc[i][2 * q][2 * h + 1] * sc_up, acc[i][2 * q + 1][2 * h + 1] * sc_up));
            } else {                                                    // the last chunk of an antenna row is cut off at column K
                const unsigned m = rg.mul_s ? __umulhi((unsigned)r, rg.mul_s) : (unsigned)r;
                const int col0 = (r - (int)m * rg.S) * J + 2 * t;       // out_u already points at this lane's column 2 t
                float2* row = out_u + (size_t)m * rg.K + (col0 - 2 * t);
                #pragma unroll
                for (int q = 0; q < NT / 2; ++q) {
                    const int col = col0 + 8 * q;
                    const float2 v0 = make_float2(acc[i][2 * q][2 * h] * sc_up, acc[i][2 * q + 1][2 * h] * sc_up);
                    const float2 v1 = make_float2(acc[i][2 * q][2 * h + 1] * sc_up, acc[i][2 * q + 1][2 * h + 1] * sc_up);
                    if (rg.vec2 && col + 1 < rg.K) __stcs(reinterpret_cast<float4*>(row + 8 * q), make_float4(v0.x, v0.y, v1.x, v1.y));
                    else {
                        if (col < rg.K)     __stcs(row + 8 * q, v0);
                        if (col + 1 < rg.K) __stcs(row + 8 * q + 1, v1);
                    }
                }
            }
        }
}

// NT: n-tiles of 8 floats per chunk, J = 4 NT subcarriers.  SB: chunks per base phasor (0: one phasor per chunk).  MP: m-tiles per
// k-step in phase 5 (2: they share the B fragments -- 8 % faster on users of >= 4 m-tiles, but 123 instead of 96 registers, which
// costs the small users more than it saves: cfg1 0.241 against 0.214 ms).  kPlain: float32 inputs, no FoV filter, isotropic patterns --
// the instantiation without the float64-input, angle (acos / atan2) and dipole code is half the size, which the instruction fetch of
// 18 warps in different phases of a ~2 500-instruction pass feels.
template <int NT, int SB, int MP, bool kPlain>
// MP = 1: 96 registers: 6 CTAs = 18 warps per SM with the 8.4 - 12.4 KB pools of J = 16 (a minimum-blocks launch bound makes ptxas stop at
// 80 and spill: measured 13 % slower; 80 registers without spilling buy 24 warps and no time); J = 32 is limited by its pools.
__global__ void __maxnreg__(NT == 4 ? (MP == 2 ? 128 : 96) : 168)
fd_mma_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ MmaCfg cfg, unsigned int* ticket)
{
    constexpr int J = 4 * NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // element m -> (y_t, z_t, y_r, z_r) panel coordinates: float64 [M][4] up to 256 elements, 16-bit [M][4] for larger panels (8 KB
    // instead of 32 KB at M = 1024; converted at every use)
    const bool wide = d.M > 256;
    double (*s_coef)[4] = reinterpret_cast<double (*)[4]>(smem_raw);
    ushort4* s_coef16 = reinterpret_cast<ushort4*>(smem_raw);
    double* s_k0 = reinterpret_cast<double*>(smem_raw + (wide ? (((size_t)d.M * 8 + 15) & ~size_t(15)) : (size_t)d.M * 32));   // [S] chunk segment -> subcarrier offset of its first column
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int m = tid; m < d.M; m += kMmWarps * 32) {
        const int r = m / d.Mt, t = m - r * d.Mt;
        if (wide) s_coef16[m] = make_ushort4((unsigned short)(t % d.bs0), (unsigned short)(t / d.bs0), (unsigned short)(r % d.ue0), (unsigned short)(r / d.ue0));
        else {
            s_coef[m][0] = (double)(t % d.bs0); s_coef[m][1] = (double)(t / d.bs0);
            s_coef[m][2] = (double)(r % d.ue0); s_coef[m][3] = (double)(r / d.ue0);
        }
    }
    for (int s = tid; s < cfg.S; s += kMmWarps * 32) s_k0[s] = (double)d.subc_start + (double)d.subc_step * (double)(J * s);

    unsigned char* wsm = smem_raw + cfg.off_warps + warp * cfg.warp_bytes;
    unsigned char* sAh  = wsm + cfg.off_A;                                     // A side, hi halves; lo halves follow
    unsigned char* sAl  = sAh + cfg.G * 8 * 256;
    unsigned char* sBh  = wsm + cfg.off_B;                                     // B side, hi halves; lo halves follow
    unsigned char* sBl  = sBh + (NT / 2) * 8 * 128;
    unsigned char* list = wsm + cfg.off_list;                                  // [32] (user in window << 5) | column
    int* s_base         = reinterpret_cast<int*>(wsm + cfg.off_meta);          // [kMmWindow + 1] first pool slot of the user (even)
    int* s_cnt          = s_base + kMmWindow + 1;                              // [kMmWindow] contributing paths
    float* s_scale      = reinterpret_cast<float*>(s_cnt + kMmWindow);         // [kMmWindow] 2^e undoing the operand scale
    for (int o = lane * 16; o < cfg.off_list; o += 32 * 16) *reinterpret_cast<uint4*>(wsm + o) = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                              // the only CTA-wide barrier: warps are independent from here on

    // Work distribution: the grid is the resident CTAs; every warp draws runs of consecutive users from a device counter
    // (ticket[0] = next user) until its draw starts past the last user.  The draws shrink towards the end of the launch (8, 4, then 2
    // users: long draws fill the rounds of 32 lanes better, short ones even out the tail).  ticket[1] counts the warps that are
    // through; the last one resets both for the next launch.  A static split left the second wave of CTAs 60 % full (cfg1: 14.5 of 21
    // warps active on average).
    long long u_begin = 0, u_end = 0;
    const unsigned n_users32 = (unsigned)d.n_users;
    const unsigned ltmask = (1u << lane) - 1u;
    const int K = d.K, M = d.M, P0 = d.P0;
    const bool need_angles = kPlain ? false : prologue_needs_angles(d);
    const bool in_f64 = kPlain ? false : (d.in_f64 != 0);
    const bool fov_any = kPlain ? false : (d.fov_any != 0);
    const bool triv0 = side_angles_trivial(d, 0), triv1 = side_angles_trivial(d, 1);
    const int g = lane >> 2, t = lane & 3;
    const double kstep = (double)d.subc_step;

    // power rows of the users [from, from + kMmWindow): loaded one pass ahead (right after a pass knows how many users it takes), so
    // that their latency runs under the operand and MMA phases
    float pw[kMmWindow];                                                       // only the NaN-ness of the power is used
    auto load_window = [&](long long from) {
        const long long prow = from * (long long)d.ld + lane;
        #pragma unroll
        for (int ul = 0; ul < kMmWindow; ++ul) {
            pw[ul] = __int_as_float(0x7fc00000);
            if (from + ul < u_end && lane < P0)
                pw[ul] = in_f64 ? (float)__ldg(reinterpret_cast<const double*>(d.power) + prow + ul * d.ld) : __ldg(d.power + prow + ul * d.ld);
        }
    };
    for (;;) {
    unsigned first = 0, cnt = 0;
    if (lane == 0) {
        const unsigned seen = *reinterpret_cast<volatile unsigned int*>(ticket);      // may be stale: it only sizes the draw
        const unsigned left = seen < n_users32 ? n_users32 - seen : 0u;
        cnt = left > cfg.draw8_above ? 8u : (left > cfg.draw4_above ? 4u : 2u);
        if (cfg.users_per_warp > 0) cnt = (unsigned)cfg.users_per_warp;                 // pinned draw size (A/B timing)
        first = atomicAdd(ticket, cnt);
    }
    first = __shfl_sync(0xffffffffu, first, 0);
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    if (first >= n_users32) break;
    u_begin = (long long)first;
    u_end = min(u_begin + (long long)cnt, d.n_users);
    load_window(u_begin);
    for (long long cur = u_begin; cur < u_end; ) {
        // ---- 1. window: which columns of the next users run their chain (every lane holds the four masks)
        const int n_in = (int)min((long long)kMmWindow, u_end - cur);
        unsigned vbs[kMmWindow], nbs[kMmWindow];
        #pragma unroll
        for (int ul = 0; ul < kMmWindow; ++ul) {
            const bool in = ul < n_in && lane < P0;
            const bool valid = in && lane < d.P && !(pw[ul] != pw[ul]);           // channel.py:260, dataset.py:258-261
            vbs[ul] = __ballot_sync(0xffffffffu, valid);
            nbs[ul] = fov_any ? __ballot_sync(0xffffffffu, in) : vbs[ul];
        }
        // ---- 2. whole users, in order, while their pairs fit one round of lanes and the pool (a user's first slot is even: a lane
        //         loads the operands of its two paths of a k-step with one aligned access).  Measured and left out: first fit over a
        //         window of 8 pending users fills 27 instead of 23 lanes but costs more in the window than it saves in the chains.
        int cum = 0, slots = 0, n_take = 0;
        bool open = true;
        #pragma unroll
        for (int ul = 0; ul < kMmWindow; ++ul) {
            const unsigned nb = nbs[ul], vb = vbs[ul];
            const int c = __popc(nb);
            const int base = (slots + 1) & ~1;
            open = open && ul < n_in && (ul == 0 || (cum + c <= 32 && base + c <= kMmSlots));
            if (open) {                                                         // warp-uniform
                if (lane == 0) { s_base[ul] = base; s_cnt[ul] = 0; s_scale[ul] = 0.f; }
                const bool run = (nb >> lane) & 1u;
                if (run) list[cum + __popc(nb & ltmask)] = (unsigned char)((ul << 5) | lane);
                if (lane < P0) {
                    const long long o = (cur + ul) * (long long)P0 + lane;
                    if (d.valid_mask) d.valid_mask[o] = (vb >> lane) & 1u;
                    if (!run) {                                                 // no FoV mask is built and the column has no power
                        if (d.fov_mask)  d.fov_mask[o] = 1;
                        if (d.clip_mask) d.clip_mask[o] = 0;
                    }
                }
                cum += c;
                slots = base + c;
                ++n_take;
            }
        }
        __syncwarp();
        const long long next_cur = cur + n_take;
        load_window(next_cur);
        if (lane < 7 * 4) {                                                     // rows of the users after this pass towards L2
            const int arr = lane >> 2;
            const long long u = next_cur + (lane & 3);
            if (u < u_end) {
                const float* base = (arr == 0) ? d.power : (arr == 1) ? d.phase : (arr == 2) ? d.delay : (arr == 3) ? d.az[0] : (arr == 4) ? d.el[0]
                                  : (arr == 5) ? d.az[1] : d.el[1];
                asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(base) + u * (long long)d.ld * (in_f64 ? 8 : 4)));
            }
        }
        // ---- 3. the chain round: lane = dense pair (cum <= 32)
        const bool act = lane < cum;
        int ul = kMmWindow, col = 0;
        if (act) { const int e = list[lane]; ul = e >> 5; col = e & 31; }
        PathState st;
        st.contrib = false; st.valid = false; st.fov = true; st.over = false;
        st.c = make_float2(0.f, 0.f); st.wcyc = 0.0; st.u[0] = st.u[1] = st.v[0] = st.v[1] = 0.0;
        if (act) {
            const long long user = cur + ul;
            SideOut s0, s1; GainOut gn;
            if (!in_f64) {                                                      // the seven entries of the pair in one batch of loads
                PathIn in;
                load_path_in(d, user, col, in);
                if (need_angles) { prologue_side_in<true>(d, user, 0, in, s0, d.Mt > 1);  prologue_side_in<true>(d, user, 1, in, s1, d.Mr > 1); }
                else {
                    if (triv0) prologue_side_trivial(in.el[0], in.az[0], s0); else prologue_side_in<false>(d, user, 0, in, s0, d.Mt > 1);
                    if (triv1) prologue_side_trivial(in.el[1], in.az[1], s1); else prologue_side_in<false>(d, user, 1, in, s1, d.Mr > 1);
                }
                prologue_gain_in<true>(d, col, in, gn);
            } else {
                if (need_angles) { prologue_side<true>(d, user, col, 0, s0, d.Mt > 1);  prologue_side<true>(d, user, col, 1, s1, d.Mr > 1); }
                else             { prologue_side<false>(d, user, col, 0, s0, d.Mt > 1); prologue_side<false>(d, user, col, 1, s1, d.Mr > 1); }
                prologue_gain<true>(d, user, col, gn);
            }
            prologue_combine<true, kPlain>(d, s0, s1, gn, st);
            const long long om = user * (long long)P0 + col;
            if (d.fov_mask)  d.fov_mask[om]  = st.fov ? 1 : 0;
            if (d.clip_mask) d.clip_mask[om] = (st.valid && st.over) ? 1 : 0;
        }
        const bool contrib = act && st.contrib;
        const unsigned bc = __ballot_sync(0xffffffffu, contrib);
        const unsigned same = __match_any_sync(0xffffffffu, ul);                 // lanes of the same user (idle lanes: key kMmWindow)
        // operand scale: 2^-(e + 1), e = exponent of the user's largest |re|, |im| of a path gain
        const unsigned amp_bits = contrib ? __float_as_uint(fmaxf(fabsf(st.c.x), fabsf(st.c.y))) : 0u;
        unsigned ef = __reduce_max_sync(same, amp_bits) >> 23;
        ef = ef < 1u ? 1u : (ef > 252u ? 252u : ef);
        const float sc_dn = __uint_as_float((253u - ef) << 23);
        const int slot = act ? s_base[ul] + __popc(bc & same & ltmask) : 0;      // contributing paths of a user, in column order
        if (act && lane == __ffs(same) - 1) { s_cnt[ul] = __popc(bc & same); s_scale[ul] = __uint_as_float((ef + 1u) << 23); }
        const float2 cs = make_float2(st.c.x * sc_dn, st.c.y * sc_dn);
        // ---- 4a. F: column j = 4 b + i of a chunk -> f4[b] * f1[i], stored once as (Fr, -Fi) = rows (2p, 2p + 1) of B for the real parts
        float2 wb[SB > 0 ? SB : 1];                                              // block phasors of the L rows
        if (contrib) {
            float2 f1[4];
            f1[0] = make_float2(1.f, 0.f);
            #pragma unroll
            for (int i = 1; i < 4; ++i) f1[i] = phasor_cycles_sfu(-(st.wcyc * (kstep * (double)i)));
            #pragma unroll 1
            for (int b = 0; b < NT; ++b) {                                       // rolled: the pass has to stay inside the instruction cache
                float2 f4b = make_float2(1.f, 0.f);
                if (b) f4b = phasor_cycles_sfu(-(st.wcyc * (kstep * (double)(4 * b))));
                const int row_b = (b >> 1) * 8 + 4 * (b & 1);                       // j = 4 b + i: column group j >> 3, column j & 7
                #pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = cmul(f4b, f1[i]);
                    unsigned x, y;                                               // (lo half = Fr, hi half = -Fi): rows (2p, 2p + 1) of B for the real parts
                    split_f16x2(f.x, -f.y, x, y);
                    *reinterpret_cast<unsigned*>(sBh + mm_b_off(row_b + i, slot)) = x;
                    *reinterpret_cast<unsigned*>(sBl + mm_b_off(row_b + i, slot)) = y;
                }
            }
            if constexpr (SB > 0) {
                wb[0] = make_float2(1.f, 0.f);
                #pragma unroll
                for (int i = 1; i < SB; ++i)                                      // chunk i of a block: antenna step i >> lg along y, segment i & mask
                    wb[i] = phasor_cycles_sfu(fma((double)(i >> cfg.lg_blk_seg), st.u[0], -(st.wcyc * (kstep * (double)(J * (i & ((1 << cfg.lg_blk_seg) - 1)))))));
            }
        }
        float2* out_pass = d.out + cur * (long long)M * K;
        #pragma unroll 1
        for (int mt0 = 0; mt0 < cfg.n_mt; mt0 += cfg.G) {
            const int n_g = min(cfg.G, cfg.n_mt - mt0);
            // ---- 4b. L rows of the m-tiles mt0 .. mt0 + n_g
            if (contrib) {
                #pragma unroll 1
                for (int ml = 0; ml < n_g; ++ml) {
                    auto chunk_phasor = [&](unsigned r) {                         // gain x steering x coarse delay phasor of chunk r
                        r = min(r, (unsigned)(cfg.R - 1));
                        const unsigned m = cfg.mul_s ? __umulhi(r, cfg.mul_s) : r;
                        const unsigned seg = r - m * (unsigned)cfg.S;
                        double c0, c1, c2, c3;
                        if (wide) { const ushort4 q = s_coef16[m]; c0 = (double)q.x; c1 = (double)q.y; c2 = (double)q.z; c3 = (double)q.w; }
                        else      { c0 = s_coef[m][0]; c1 = s_coef[m][1]; c2 = s_coef[m][2]; c3 = s_coef[m][3]; }
                        const double cyc = fma(c0, st.u[0], fma(c1, st.v[0], fma(c2, st.u[1], fma(c3, st.v[1], -(st.wcyc * s_k0[seg])))));
                        return cmul(cs, phasor_cycles_sfu(cyc));
                    };
                    const unsigned r0 = (unsigned)(mt0 + ml) * 16u;
                    if constexpr (SB > 0) {
                        constexpr int SBc = SB, NB = 16 / SBc;
                        float2 base[NB];
                        #pragma unroll
                        for (int b = 0; b < NB; ++b) base[b] = chunk_phasor(r0 + b * SBc);
                        #pragma unroll
                        for (int gg = 0; gg < 8; ++gg) {
                            const float2 l0 = (gg % SBc == 0) ? base[gg / SBc] : cmul(base[gg / SBc], wb[gg % SBc]);
                            const float2 l1 = ((gg + 8) % SBc == 0) ? base[(gg + 8) / SBc] : cmul(base[(gg + 8) / SBc], wb[(gg + 8) % SBc]);
                            unsigned h0, q0, h1, q1;
                            split_f16x2(l0.x, l0.y, h0, q0);
                            split_f16x2(l1.x, l1.y, h1, q1);
                            const int off = mm_a_off(ml * 8 + gg, slot);
                            *reinterpret_cast<uint2*>(sAh + off) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(sAl + off) = make_uint2(q0, q1);
                        }
                    } else {
                        #pragma unroll 2
                        for (int gg = 0; gg < 8; ++gg) {
                            const float2 l0 = chunk_phasor(r0 + gg), l1 = chunk_phasor(r0 + gg + 8);
                            unsigned h0, q0, h1, q1;
                            split_f16x2(l0.x, l0.y, h0, q0);
                            split_f16x2(l1.x, l1.y, h1, q1);
                            const int off = mm_a_off(ml * 8 + gg, slot);
                            *reinterpret_cast<uint2*>(sAh + off) = make_uint2(h0, h1);
                            *reinterpret_cast<uint2*>(sAl + off) = make_uint2(q0, q1);
                        }
                    }
                }
            }
            __syncwarp();
            // ---- 5. per user: fragments, MMAs, stores; two m-tiles at a time share the B fragments where the accumulators fit (J = 16)
            const unsigned a_sw = (unsigned)(g & 1) << 6, b_sw = (unsigned)(g & 3) << 5;        // swizzles of this lane's fragment rows
            const MmRagged rg = {cfg.ragged, K, cfg.S, (K & 1) == 0, cfg.mul_s};
            const unsigned char* bh_row = sBh + g * 128;
            const unsigned char* bl_row = sBl + g * 128;
            #pragma unroll 1
            for (int uu = 0; uu < n_take; ++uu) {
                const int np = s_cnt[uu];
                const float sc_up = s_scale[uu];
                float2* out_u = out_pass + (size_t)uu * (size_t)(M * K) + 2 * t;
                const int qb = s_base[uu] + 2 * t;
                int ml = 0;
                if constexpr (MP == 2) {
                    #pragma unroll 1
                    for (; ml + 1 < n_g; ml += 2)
                        mm_consume<NT, 2, !kPlain>(sAh, sAl, bh_row, bl_row, ml * 8 + g, a_sw, b_sw, qb, np, t, sc_up, out_u, (mt0 + ml) * 16 + g, cfg.R, rg);
                }
                #pragma unroll 1
                for (; ml < n_g; ++ml)
                    mm_consume<NT, 1, !kPlain>(sAh, sAl, bh_row, bl_row, ml * 8 + g, a_sw, b_sw, qb, np, t, sc_up, out_u, (mt0 + ml) * 16 + g, cfg.R, rg);
            }
            __syncwarp();                                                       // the L rows are rewritten by the next group / pass
        }
        cur = next_cur;
    }
    }
    if (lane == 0 && atomicAdd(ticket + 1, 1u) == gridDim.x * (unsigned)kMmWarps - 1u) {      // the last warp of the launch
        atomicExch(ticket, 0u);
        atomicExch(ticket + 1, 0u);
    }
}

}  // namespace dmk
