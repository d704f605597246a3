// dmk_fd_ws.cuh -- FD channel kernel on tcgen05, persistent and warp-specialised (the production tensor-core kernel).
//
// Same arithmetic as fd_tc_kernel (dmk_fd_tc.cuh: transposed tile, FP16 hi/lo operand split, FP32 accumulation in
// TMEM, per-user operand scale); what changes is who does what and when.  Phase traces of the one-CTA-per-user kernel
// on the city-scale shape (64 x 1024 x 8 B = 512 KB per user; tools/tc_trace.py, profiles/README.md) showed
//   * 20 % of a CTA's life in the float64 prologue + table building (no stores),
//   * every stage serialised as  build operands -> MMA -> drain, with all eight warps in the same phase, so the SM's
//     store stream stops whenever both resident CTAs build,
//   * the MMA batch itself taking 3-8 k cycles because a lone thread under `if (tid == 256)` makes the compiler wrap
//     every tcgen05.mma in an ELECT / R2UR.BROADCAST waterfall (~30 dependent instructions per MMA).
// Roles here (320 threads, 2 persistent CTAs per SM, users drawn from a device-side ticket counter):
//   warp 9      helper   ticket -> float64 prologue of the NEXT user (lanes = path columns) -> its phasor tables
//   warps 4-7   builders A/B operand tiles of stage g+1 (K-major SWIZZLE_128B, FP16 hi/lo) as soon as MMA(g) has read its operands
//   warp 8      issuer   elect.sync lane issues the tcgen05.mma batch of stage g, commits to mma_done[g & 1]
//   warps 0-3   drain    TMEM lane quarter q = warp: tcgen05.ld -> scale -> 128-byte row-segment stores of stage g-1
// All hand-offs are mbarriers (no CTA-wide barrier in the steady state); the accumulator is double-buffered in TMEM,
// the table set and the user record are double-buffered in shared memory.  Stores therefore never wait for operand
// building or for the prologue -- only for HBM.
#pragma once
#include "dmk_fd_tc.cuh"

namespace dmk {

constexpr int kWsConsumers = 288;  // warps 0-8: threads that read a user record (and arrive on ub_empty)
constexpr int kWsDrain0   = 0;     // warps 0-3
constexpr int kWsBuild0   = 4;     // warps 4-7
constexpr int kWsIssuer   = 8;
constexpr int kWsHelper0  = 9;     // warps 9 .. 9 + H - 1
constexpr int kWsBuilders = 128;
constexpr int kWsMaxHelpers = 4;

struct WsBars {
    uint64_t ub_full[2 * kWsMaxHelpers], ub_empty[2 * kWsMaxHelpers];   // helper -> everyone (1 arrival) ; everyone -> helper (288 arrivals)
    uint64_t op_full;                   // builders -> issuer (128 arrivals): operand tiles of the next stage are in smem
    uint64_t mma_done[2];               // tcgen05.commit: accumulator g & 1 complete, operand tiles free again
    uint64_t acc_empty[2];              // drain warps -> issuer (128 arrivals): accumulator g & 1 has been read out
};

__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
// Wait for the phase with the given parity to complete.  A lost arrival must fail loudly rather than hang the device, but a
// legitimately slow neighbour (time-sliced GPU, debugger) must not be killed: the watchdog runs on the global nanosecond
// timer and only fires after 30 s without progress on this barrier.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    const uint32_t bar = smem_u32(b);
    uint32_t ok = 0;
    unsigned long long t0 = 0;
    for (unsigned spins = 0; ; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if ((spins & 4095u) == 4095u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 30000000000ULL) __trap();
        }
    }
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}

// One complex value per path slot x 4 -> FP16 hi/lo, written to the row pair (2c: Re H, 2c+1: Im H) of a B tile:
// row 2c holds (w.x, -w.y), row 2c+1 holds (w.y, w.x); the second row's halves are a sign flip / swap of the first's.
__device__ __forceinline__ void st_split8_f16_rowpair(unsigned char* hi, unsigned char* lo, int off_re, int off_im, const float2 (&w)[4])
{
    uint32_t hr[4], lr[4], hi_[4], li[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(w[i].x, w[i].y);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(w[i].x - back.x, w[i].y - back.y);
        const uint32_t h = *reinterpret_cast<const uint32_t*>(&hh), l = *reinterpret_cast<const uint32_t*>(&ll);
        hr[i] = h ^ 0x80000000u;  lr[i] = l ^ 0x80000000u;                    // (x, -y)
        hi_[i] = __byte_perm(h, 0, 0x1032); li[i] = __byte_perm(l, 0, 0x1032);  // (y, x)
    }
    *reinterpret_cast<uint4*>(hi + off_re) = make_uint4(hr[0], hr[1], hr[2], hr[3]);
    *reinterpret_cast<uint4*>(lo + off_re) = make_uint4(lr[0], lr[1], lr[2], lr[3]);
    *reinterpret_cast<uint4*>(hi + off_im) = make_uint4(hi_[0], hi_[1], hi_[2], hi_[3]);
    *reinterpret_cast<uint4*>(lo + off_im) = make_uint4(li[0], li[1], li[2], li[3]);
}

// Helper warp: ticket -> prologue -> per-user tables of one buffer.  lanes = path columns, then lanes = table entries.
// Row np of every table is zero-filled: the operand builders read it (index min(p, np)) for the padding slots.
__device__ __noinline__ void ws_helper_prepare(const DevDesc& d, const TcCfg& cfg, int ksplit, unsigned int n_items, unsigned int n_draw_last,
                                               unsigned int* ticket, TcUserBuf& ub, unsigned char* tab, int lane)
{
    unsigned int t = 0;
    if (lane == 0) {
        t = atomicAdd(ticket, 1u);
        if (t == n_draw_last) atomicExch(ticket, 0u);      // every helper warp of every CTA draws exactly one ticket >= n_items: this is the last draw
        ub.item = t;
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= n_items) return;
    const long long user = t / (unsigned)ksplit;
    const bool write_masks = (t % (unsigned)ksplit) == 0;
    FdShared& sh = ub.sh;

    PathState st;
    const bool active = lane < d.P0;
    st.contrib = false; st.valid = false; st.fov = true; st.over = false;
    st.c = make_float2(0.f, 0.f);
    if (active) {
        SideOut s0, s1;
        GainOut g;
        if (prologue_needs_angles(d)) { prologue_side<true>(d, user, lane, 0, s0);  prologue_side<true>(d, user, lane, 1, s1); }
        else                          { prologue_side<false>(d, user, lane, 0, s0); prologue_side<false>(d, user, lane, 1, s1); }
        prologue_gain<true>(d, user, lane, g);
        prologue_combine<true>(d, s0, s1, g, st);
    }
    const bool contrib = active && st.contrib;
    const unsigned ballot = __ballot_sync(0xffffffffu, contrib);
    const int np = __popc(ballot);
    if (contrib) {
        const int j = __popc(ballot & ((1u << lane) - 1u));
        sh.c[j] = st.c; sh.wcyc[j] = st.wcyc; sh.fd[j] = st.fd;
        sh.u[0][j] = st.u[0]; sh.v[0][j] = st.v[0];
        sh.u[1][j] = st.u[1]; sh.v[1][j] = st.v[1];
    }
    if (lane == 0) sh.np = np;
    if (write_masks && active) {
        const long long o = user * (long long)d.P0 + lane;
        if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
        if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
        if (d.clip_mask)  d.clip_mask[o]  = (st.valid && st.over) ? 1 : 0;
    }
    if (np == 0) return;
    // per-user operand scale: largest |c_p| component (FP16 operands live in [-1, 1])
    float mx = contrib ? fmaxf(fabsf(st.c.x), fabsf(st.c.y)) : 0.f;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) ub.scale = mx;
    const float inv_scale = 1.0f / mx;
    __syncwarp();

    float2* tY   = reinterpret_cast<float2*>(tab + cfg.off_tY);
    float2* tQ   = reinterpret_cast<float2*>(tab + cfg.off_tQ);
    float2* wB   = reinterpret_cast<float2*>(tab + cfg.off_wB);
    float2* seed = reinterpret_cast<float2*>(tab + cfg.off_seed);
    const int bs0 = d.bs0, bs1 = d.bs1, nq = d.Mr * d.bs1;
    for (int e = lane; e < np * bs0; e += 32) {
        const int p = e / bs0, y = e - p * bs0;
        tY[p * cfg.sY + y] = phasor_cycles((double)y * sh.u[0][p]);
    }
    for (int e = lane; e < np * nq; e += 32) {
        const int p = e / nq, q = e - p * nq;
        const int r = q / bs1, z = q - r * bs1;
        const int yr = r % d.ue0, zr = r / d.ue0;
        const float2 cs = make_float2(sh.c[p].x * inv_scale, sh.c[p].y * inv_scale);
        tQ[p * cfg.sQ + q] = cmul(cs, phasor_cycles((double)z * sh.v[0][p] + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
    }
    for (int e = lane; e < np * 16; e += 32) {
        const int p = e >> 4, b = e & 15;
        wB[p * cfg.sB + b] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * b)));
    }
    const int n_hi = (cfg.nA + 7) >> 3, n_sd = 8 + n_hi;
    for (int e = lane; e < np * n_sd; e += 32) {
        const int p = e / n_sd, b = e - p * n_sd;
        seed[p * cfg.sS + b] = (b < 8) ? phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * 16 * b)))                             // seed_lo[b]
                                       : phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_start + d.subc_step * 128 * (b - 8))));     // seed_hi[b - 8]
    }
    // row np of every table is the zero row: the operand builders read it for the padding slots np .. nslot-1
    const float2 zero = make_float2(0.f, 0.f);
    for (int e = lane; e < bs0; e += 32)  tY[np * cfg.sY + e] = zero;
    for (int e = lane; e < nq; e += 32)   tQ[np * cfg.sQ + e] = zero;
    for (int e = lane; e < 16; e += 32)   wB[np * cfg.sB + e] = zero;
    for (int e = lane; e < n_sd; e += 32) seed[np * cfg.sS + e] = zero;
}

// Drain one accumulator (mtile antenna rows x 128 floats of segment `seg`): the warp owns TMEM lane quarter q.
__device__ __forceinline__ void ws_drain(uint32_t tmem_base, int acc_col, float* out_u, long long pitch, int M, int mtile,
                                         int row0, int seg, int q, int lane, float scale)
{
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc_col;
    float* o = out_u + (long long)row0 * pitch + seg * kTcN + q * 32 + lane;
    if (mtile >= 32) {
        for (int r = 0; r < mtile; r += 32)
            tc_store_rows<32>(taddr + r, o + (long long)r * pitch, pitch, row0 + r, M, scale);
    } else {
        tc_store_rows<16>(taddr, o, pitch, row0, M, scale);
    }
}

// Consumers (drain, builders, issuer) walk the users in the order it = 0, 1, 2, ...: user `it` is prepared by helper it % H into
// buffer it % (2H) (every helper owns two buffers).  A helper that draws a ticket >= n_items publishes it as a sentinel and stops;
// its slots are skipped from then on, and the walk ends when every helper of the CTA has stopped.
template <int H>
__device__ __forceinline__ bool ws_next_user(WsBars& bars, const unsigned char* bufs, int buf_stride, unsigned n_items,
                                             unsigned& it, unsigned& done, int& b)
{
    constexpr unsigned R = 2 * H, kAll = (1u << H) - 1u;
    for (;;) {
        if (done == kAll) return false;
        const unsigned h = it % H;
        if (!((done >> h) & 1u)) {
            b = (int)(it % R);
            mbar_wait(&bars.ub_full[b], (it / R) & 1u);
            if (reinterpret_cast<const TcUserBuf*>(bufs + (size_t)b * buf_stride)->item < n_items) return true;
            done |= 1u << h;
        }
        ++it;
    }
}

template <int H>        // helper warps per CTA: 1 (two CTAs per SM, large per-user outputs) or 4 (one CTA per SM, helper-bound shapes)
__global__ void __launch_bounds__((9 + H) * 32, H == 1 ? 2 : 1)
fd_ws_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ TcCfg cfg, const int ksplit,
             const unsigned int n_items, unsigned int* ticket, const int pdl_wait)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ WsBars bars;
    if (pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");     // plain stream order unless the caller declared independence
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the next launch may take SMs as our CTAs retire
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 sWa[kTcSlots * 9];              // stage-local coarse delay phasors [slot][8 groups of 16 subcarriers], stride 9

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mtile = cfg.mtile, nsub = cfg.nsub;

    unsigned char* sm = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
    unsigned char* sAhi = sm + cfg.off_A;                    // [mtile rows][128 B]: 32 path slots x (re, im) fp16
    unsigned char* sAlo = sAhi + mtile * 128;
    unsigned char* sBhi = sm + cfg.off_B;                    // per sub-tile: [128 rows][128 B] hi, then lo
    unsigned char* sBlo = sBhi + kTcN * 128;
    unsigned char* bufs = sm + cfg.off_tab;                  // 2H buffers of cfg.tab_bytes: [TcUserBuf][tables]
    constexpr int kUb = (int)((sizeof(TcUserBuf) + 15) & ~size_t(15));
    auto user_buf = [&](int b) -> TcUserBuf& { return *reinterpret_cast<TcUserBuf*>(bufs + (size_t)b * cfg.tab_bytes); };
    auto user_tab = [&](int b) -> unsigned char* { return bufs + (size_t)b * cfg.tab_bytes + kUb; };

    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(2 * 128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 64) {
        for (int b = 0; b < 2 * H; ++b) { mbar_init(&bars.ub_full[b], 1); mbar_init(&bars.ub_empty[b], kWsConsumers); }
        mbar_init(&bars.op_full, kWsBuilders);
        mbar_init(&bars.mma_done[0], 1);  mbar_init(&bars.mma_done[1], 1);
        mbar_init(&bars.acc_empty[0], 128); mbar_init(&bars.acc_empty[1], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int K = d.K, M = d.M;
    const int n_seg = K / (kTcN / 2);                         // 64-subcarrier segments per row
    const int n_ct = (n_seg + nsub - 1) / nsub;               // pipeline stages (column super-tiles) per row tile
    const int n_rt = (M + mtile - 1) / mtile;
    const long long pitch = 2LL * K;                          // floats per output row
    const int acc_stride = 128 / nsub;                        // TMEM columns per accumulator; 2 * nsub accumulators in 256 columns

    if (warp >= kWsHelper0) {
        // ------------------------------------------------------------------------------------------ helpers
        const int h = warp - kWsHelper0;
        const unsigned int n_draw_last = n_items + gridDim.x * (unsigned)H - 1u;
        for (unsigned k = 0;; ++k) {
            const int b = h + (int)(k & 1u) * H;                          // this helper's two buffers, alternately
            mbar_wait(&bars.ub_empty[b], ((k >> 1) & 1u) ^ 1u);          // the readers of this buffer's previous user are done
            ws_helper_prepare(d, cfg, ksplit, n_items, n_draw_last, ticket, user_buf(b), user_tab(b), lane);
            __syncwarp();
            const unsigned int item = user_buf(b).item;
            if (lane == 0) mbar_arrive(&bars.ub_full[b]);
            if (item >= n_items) break;
        }
    } else if (warp >= kWsBuild0 && warp < kWsIssuer) {
        // ------------------------------------------------------------------------------------------ operand builders
        const int bt = tid - kWsBuild0 * 32;                  // 0..127
        const int a_row  = bt & (mtile - 1);
        const int a_ngrp = kWsBuilders / mtile > 0 ? kWsBuilders / mtile : 1;      // thread groups per antenna row: 1, 2, 4 or 8
        const int a_grp  = bt / mtile;
        const int a_off0 = (a_row >> 3) * 1024 + (a_row & 7) * 128;
        const int b_col  = bt & 63;
        const int b_grp  = bt >> 6;                           // 0..1
        const int b_row0 = 2 * b_col;                         // rows 2c (Re H) and 2c + 1 (Im H)
        const int b_off0 = (b_row0 >> 3) * 1024 + (b_row0 & 7) * 128;
        const int b_sw0 = b_row0 & 7, b_sw1 = (b_row0 + 1) & 7;
        unsigned g = 0;                                       // global stage counter (identical in every role)
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            if (np > 0) {
                const unsigned char* tab = user_tab(cur);
                const float2* tY   = reinterpret_cast<const float2*>(tab + cfg.off_tY);
                const float2* tQ   = reinterpret_cast<const float2*>(tab + cfg.off_tQ);
                const float2* wB   = reinterpret_cast<const float2*>(tab + cfg.off_wB);
                const float2* seed = reinterpret_cast<const float2*>(tab + cfg.off_seed);
                const int nslot = ((np + 7) >> 3) << 3;       // slots the tensor core reads (table row np is the zero row)
                const int nq4 = nslot >> 2;
                bool a_valid = false;
                for (int ct = ks; ct < n_ct; ct += ksplit) {
                    const int seg0 = ct * nsub;
                    const int nsub_here = min(nsub, n_seg - seg0);
                    bool b_valid = false;
                    for (int rt = 0; rt < n_rt; ++rt, ++g) {
                        const int row0 = rt * mtile;
                        if (g > 0) mbar_wait(&bars.mma_done[(g - 1) & 1], ((g - 1) >> 1) & 1u);     // MMA(g-1) has read the operand tiles
                        if (!b_valid) {
                            // stage-local coarse phasors wA[p][grp] = seed_hi * seed_lo for the <= 8 groups of 16 subcarriers
                            for (int e = bt; e < nslot * 8; e += kWsBuilders) {
                                const int p = e >> 3, grp = e & 7;
                                if (grp < 4 * nsub_here) {
                                    const int a = seg0 * 4 + grp;
                                    const float2* sd = seed + min(p, np) * cfg.sS;
                                    sWa[p * 9 + grp] = cmul(sd[8 + (a >> 3)], sd[a & 7]);
                                }
                            }
                        }
                        // ---- A_hi / A_lo: antenna rows x path slots
                        if (!(a_valid && n_rt == 1) && bt < a_ngrp * mtile) {
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
                            const float2* q0 = tQ + a_q;
                            const float2* y0 = tY + a_y;
                            #pragma unroll 1
                            for (int qd = a_grp; qd < nq4; qd += a_ngrp) {
                                float2 a[4];
                                #pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int p = min(qd * 4 + i, np);
                                    a[i] = a_ok ? cmul(q0[p * cfg.sQ], y0[p * cfg.sY]) : make_float2(0.f, 0.f);
                                }
                                st_split8_f16(sAhi, sAlo, a_off0 + (((qd ^ (a_row & 7)) & 7) << 4), a);
                            }
                        }
                        // ---- B_hi / B_lo per sub-tile (rows 2c -> Re H, 2c+1 -> Im H)
                        if (!b_valid) {
                            asm volatile("bar.sync 3, %0;" :: "n"(kWsBuilders) : "memory");        // sWa complete
                            const float2* wb0 = wB + (b_col & 15);
                            #pragma unroll 1
                            for (int sub = 0; sub < nsub_here; ++sub) {
                                unsigned char* sBh = sBhi + sub * (2 * kTcN * 128);
                                unsigned char* sBl = sBh + kTcN * 128;
                                const float2* wa0 = sWa + sub * 4 + (b_col >> 4);
                                #pragma unroll 1
                                for (int qd = b_grp; qd < nq4; qd += 2) {
                                    float2 w[4];
                                    #pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int p = qd * 4 + i;
                                        w[i] = cmul(wa0[p * 9], wb0[min(p, np) * cfg.sB]);
                                    }
                                    st_split8_f16_rowpair(sBh, sBl, b_off0 + (((qd ^ b_sw0) & 7) << 4), b_off0 + 128 + (((qd ^ b_sw1) & 7) << 4), w);
                                }
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive(&bars.op_full);
                        b_valid = true;
                        a_valid = true;
                    }
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
    } else if (warp == kWsIssuer) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        // instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = antenna rows of the tile, M = 128 floats
        const uint32_t idesc = (1u << 4) | ((uint32_t)(mtile >> 3) << 17) | ((uint32_t)(kTcN >> 4) << 24);
        const uint64_t dAhi = umma_desc_kmajor_sw128(smem_u32(sAhi)), dAlo = umma_desc_kmajor_sw128(smem_u32(sAlo));
        const uint64_t dBhi = umma_desc_kmajor_sw128(smem_u32(sBhi)), dBlo = umma_desc_kmajor_sw128(smem_u32(sBlo));
        unsigned g = 0;
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            if (np > 0) {
                const int ksteps = (np + 7) >> 3;                     // 16 fp16 (8 path slots) per MMA
                for (int ct = ks; ct < n_ct; ct += ksplit) {
                    const int nsub_here = min(nsub, n_seg - ct * nsub);
                    for (int rt = 0; rt < n_rt; ++rt, ++g) {
                        const unsigned ab = g & 1;
                        mbar_wait(&bars.op_full, g & 1u);                                 // operand tiles of stage g are in smem
                        mbar_wait(&bars.acc_empty[ab], ((g >> 1) & 1u) ^ 1u);              // accumulator ab drained (stage g-2)
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one()) {
                            #pragma unroll 1
                            for (int sub = 0; sub < nsub_here; ++sub) {
                                const uint32_t acc = tmem_base + (uint32_t)(((ab * nsub) + sub) * acc_stride);
                                const uint64_t sub_off = (uint64_t)(sub * (2 * kTcN * 128) >> 4);       // descriptor address field: 16-byte units
                                #pragma unroll
                                for (int s = 0; s < 3; ++s) {                                           // hi*hi, lo*hi, hi*lo
                                    const uint64_t da = (s == 2) ? dAlo : dAhi;
                                    const uint64_t db = ((s == 1) ? dBlo : dBhi) + sub_off;
                                    #pragma unroll
                                    for (int kk = 0; kk < 4; ++kk) {
                                        if (kk < ksteps) {
                                            const uint32_t accum = (s | kk) != 0;
                                            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                                         :: "r"(acc), "l"(db + 2 * kk), "l"(da + 2 * kk), "r"(idesc), "r"(accum) : "memory");
                                        }
                                    }
                                }
                            }
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                         :: "r"(smem_u32(&bars.mma_done[ab])) : "memory");
                        }
                        __syncwarp();
                    }
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
    } else {
        // ------------------------------------------------------------------------------------------ drain warps 0-3
        const int q = warp;                                   // TMEM lane quarter = 32 floats (128 bytes) of every row segment
        unsigned g = 0;
        unsigned it = 0, done = 0;
        int cur = 0;
        for (; ws_next_user<H>(bars, bufs, cfg.tab_bytes, n_items, it, done, cur); ++it) {
            const TcUserBuf& ub = user_buf(cur);
            const unsigned int item = ub.item;
            const long long user = item / (unsigned)ksplit;
            const int ks = (int)(item % (unsigned)ksplit);
            const int np = ub.sh.np;
            const float scale = ub.scale;
            float* out_u = reinterpret_cast<float*>(d.out + user * (long long)M * K);
            if (np == 0) {
                // users without contributing paths: zeros (channel.py:257,:269-271), one 512-byte row segment per warp store
                float4* o = reinterpret_cast<float4*>(out_u);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int ct = ks; ct < n_ct; ct += ksplit)
                    for (int sub = 0; sub < nsub && ct * nsub + sub < n_seg; ++sub) {
                        float4* ot = o + (ct * nsub + sub) * (kTcN / 4) + lane;
                        for (int m = warp; m < M; m += 4) __stcs(ot + (long long)m * (pitch / 4), z);
                    }
            } else {
                for (int ct = ks; ct < n_ct; ct += ksplit) {
                    const int seg0 = ct * nsub;
                    const int nsub_here = min(nsub, n_seg - seg0);
                    for (int rt = 0; rt < n_rt; ++rt, ++g) {
                        const unsigned ab = g & 1;
                        mbar_wait(&bars.mma_done[ab], (g >> 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        for (int sub = 0; sub < nsub_here; ++sub)
                            ws_drain(tmem_base, (int)((ab * nsub + sub) * acc_stride), out_u, pitch, M, mtile, rt * mtile, seg0 + sub, q, lane, scale);
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive(&bars.acc_empty[ab]);
                    }
                }
            }
            mbar_arrive(&bars.ub_empty[cur]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 3) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(2 * 128));
}

}  // namespace dmk
