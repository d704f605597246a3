// dmk_fd_small.cuh -- FD channel kernel for small arrays (M = M_rx * M_tx <= 16): one WARP per user.
//
// With a handful of antenna rows a user's whole output is a few KB (BASELINE config 1: 8 x 64 complex = 4 KB) and the
// cost is the per-user prologue, not the accumulation.  Giving every user a CTA (fd_fast_kernel / fd_tc_kernel) leaves
// 7 of 8 warps waiting on the prologue; here a warp does everything for its user with warp-level synchronisation only:
//   lanes = path columns: the three float64 prologue chains back to back, combine, ballot compaction;
//   lanes = table entries: TX y-steering, (RX x TX-z x gain), delay phasors wB[16] and the fine/coarse seed tables
//     (phase reduced in float64 per entry);
//   lanes = column pairs: per 64-column pass each lane owns columns 2l, 2l+1 and all M rows in registers, the W entries of its
//     two columns are formed on the fly (seed_hi * seed_lo * wB), 2 FFMA2 per complex MAC as in fd_fast_kernel;
//   16-byte streaming stores, 512 contiguous bytes per row per warp instruction.
// Several warps (users) share a CTA only to share the launch; they never synchronise with each other.
#pragma once
#include "dmk_fd.cuh"

namespace dmk {

constexpr int kSmallWarps = 3;            // users per CTA: 6 CTAs = 18 warps per SM fit the 12 KB-per-warp tables and 112 registers (4 x 4 = 16 did)

struct SmallCfg {
    int off_sh, off_tY, off_tQ, off_wB, off_seed, off_A;    // byte offsets inside a warp's shared-memory slice
    int warp_bytes;
    int sY, sQ, sS, n_hi, pcap;                             // odd table strides (float2 units); sS = seed stride >= 8 + n_hi
    unsigned mul_mt, mul_bs0;
};

template <int MT>      // rows held in registers: 4, 8 or 16 (M <= MT)
__global__ void __launch_bounds__(kSmallWarps * 32, 6)
fd_small_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ SmallCfg cfg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long user = (long long)blockIdx.x * kSmallWarps + warp;
    if (user >= d.n_users) return;                           // warp-uniform; no CTA-wide barrier is ever used

    unsigned char* wsm = smem_raw + warp * cfg.warp_bytes;
    FdShared& sh = *reinterpret_cast<FdShared*>(wsm + cfg.off_sh);
    float2* tY   = reinterpret_cast<float2*>(wsm + cfg.off_tY);      // [np][sY]
    float2* tQ   = reinterpret_cast<float2*>(wsm + cfg.off_tQ);      // [np][sQ]
    float2* wB   = reinterpret_cast<float2*>(wsm + cfg.off_wB);      // [np][17]
    float2* seed = reinterpret_cast<float2*>(wsm + cfg.off_seed);    // [np][sS]: 8 fine, then n_hi coarse
    float4* sA   = reinterpret_cast<float4*>(wsm + cfg.off_A);       // [np][MT] (re, re, im, im)

    // ---- prologue: lane = path column, the three chains back to back
    {
        PathState st;
        const bool active = lane < d.P0;
        st.contrib = false; st.valid = false; st.fov = true; st.over = false;
        if (active) {
            SideOut s0, s1; GainOut g;
            if (prologue_needs_angles(d)) { prologue_side<true>(d, user, lane, 0, s0);  prologue_side<true>(d, user, lane, 1, s1); }
            else                          { prologue_side<false>(d, user, lane, 0, s0); prologue_side<false>(d, user, lane, 1, s1); }
            prologue_gain<true>(d, user, lane, g);
            prologue_combine<true>(d, s0, s1, g, st);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, active && st.contrib);
        if (active && st.contrib) {
            const int j = __popc(ballot & ((1u << lane) - 1u));
            sh.c[j] = st.c; sh.wcyc[j] = st.wcyc;
            sh.u[0][j] = st.u[0]; sh.v[0][j] = st.v[0];
            sh.u[1][j] = st.u[1]; sh.v[1][j] = st.v[1];
        }
        if (lane == 0) sh.np = __popc(ballot);
        if (active) {
            const long long o = user * (long long)d.P0 + lane;
            if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
            if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
            if (d.clip_mask)  d.clip_mask[o]  = (st.valid && st.over) ? 1 : 0;
        }
    }
    __syncwarp();
    const int np = sh.np;
    const int K = d.K, M = d.M;
    float2* out_u = d.out + user * (long long)M * K;
    const bool vec_ok = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(d.out) & 15) == 0);

    if (np == 0) {                                           // zeros (channel.py:257,:269-271)
        const long long total = (long long)M * K;
        if (vec_ok) { float4* o = reinterpret_cast<float4*>(out_u); for (long long e = lane; e < total / 2; e += 32) __stcs(o + e, make_float4(0.f, 0.f, 0.f, 0.f)); }
        else        { for (long long e = lane; e < total; e += 32) __stcs(out_u + e, make_float2(0.f, 0.f)); }
        return;
    }

    // ---- tables (phase reduced in float64 for every entry)
    const int nq = d.Mr * d.bs1;
    for (int e = lane; e < np * d.bs0; e += 32) {
        const int p = e / d.bs0, y = e - p * d.bs0;
        tY[p * cfg.sY + y] = phasor_cycles((double)y * sh.u[0][p]);
    }
    for (int e = lane; e < np * nq; e += 32) {
        const int p = e / nq, q = e - p * nq;
        const int r = q / d.bs1, z = q - r * d.bs1;
        const int yr = r % d.ue0, zr = r / d.ue0;
        tQ[p * cfg.sQ + q] = cmul(sh.c[p], phasor_cycles((double)z * sh.v[0][p] + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
    }
    for (int e = lane; e < np * 16; e += 32) {
        const int p = e >> 4, b = e & 15;
        wB[p * 17 + b] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * b)));
    }
    for (int e = lane; e < np * (8 + cfg.n_hi); e += 32) {
        const int p = e / (8 + cfg.n_hi), b = e - p * (8 + cfg.n_hi);
        const double k0 = (b < 8) ? (double)(d.subc_step * 16 * b) : (double)(d.subc_start + d.subc_step * 128 * (b - 8));
        seed[p * cfg.sS + b] = phasor_cycles(-(sh.wcyc[p] * k0));
    }
    __syncwarp();
    // ---- A[m][p] = gain * steering, stored (re, re, im, im) for the FFMA2 operand modifiers
    for (int e = lane; e < np * MT; e += 32) {
        const int p = e / MT, m = e - p * MT;
        float2 a = make_float2(0.f, 0.f);
        if (m < M) {
            const unsigned mm = (unsigned)m;
            const unsigned rr = cfg.mul_mt ? __umulhi(mm, cfg.mul_mt) : mm;
            const unsigned t = mm - rr * (unsigned)d.Mt;
            const unsigned zt = cfg.mul_bs0 ? __umulhi(t, cfg.mul_bs0) : t;
            const unsigned yt = t - zt * (unsigned)d.bs0;
            a = cmul(tQ[p * cfg.sQ + rr * d.bs1 + zt], tY[p * cfg.sY + yt]);
        }
        sA[p * MT + m] = make_float4(a.x, a.x, a.y, a.y);
    }
    __syncwarp();

    // ---- accumulation: 64 columns per pass, lane owns columns 2l and 2l+1
    for (int col0 = 0; col0 < K; col0 += 64) {
        const int c0 = col0 + 2 * lane;
        const int cc = min(c0, K - 1);                       // clamp for the table indices of idle lanes
        const int a_idx = cc >> 4, b0 = cc & 15, b1 = (cc + 1) & 15;     // c0 is even: both columns share the coarse index
        float2 acc[MT][2];
        #pragma unroll
        for (int m = 0; m < MT; ++m) { acc[m][0] = make_float2(0.f, 0.f); acc[m][1] = make_float2(0.f, 0.f); }
        #pragma unroll 1
        for (int p = 0; p < np; ++p) {
            const float2* sd = seed + p * cfg.sS;
            const float2 wa = cmul(sd[8 + (a_idx >> 3)], sd[a_idx & 7]);
            const float2 w0 = cmul(wa, wB[p * 17 + b0]);
            const float2 w1 = cmul(wa, wB[p * 17 + b1]);
            const float2 w0s = make_float2(w0.y, w0.x), w1s = make_float2(w1.y, w1.x);
            #pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float4 a = sA[p * MT + m];
                const float2 a1 = make_float2(a.x, a.y), a2 = make_float2(-a.z, a.w);
                acc[m][0] = __ffma2_rn(a1, w0, acc[m][0]);
                acc[m][0] = __ffma2_rn(a2, w0s, acc[m][0]);
                acc[m][1] = __ffma2_rn(a1, w1, acc[m][1]);
                acc[m][1] = __ffma2_rn(a2, w1s, acc[m][1]);
            }
        }
        #pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m >= M) break;
            float2* o = out_u + (long long)m * K + c0;
            if (vec_ok && c0 + 1 < K) __stcs(reinterpret_cast<float4*>(o), make_float4(acc[m][0].x, acc[m][0].y, acc[m][1].x, acc[m][1].y));
            else {
                if (c0 < K)     __stcs(o, acc[m][0]);
                if (c0 + 1 < K) __stcs(o + 1, acc[m][1]);
            }
        }
    }
}


// =================================================================================================
// fd_small2_kernel -- small arrays (M <= 16), densely packed (round 2).
//
// ncu on fd_small_kernel (profiles/r01_ncu_fd_small_cfg1.txt) counted ~3 000 warp-instructions per user for 4 KB of output: the
// float64 chains ran with lanes = path COLUMNS (25 of 32 lanes, of which only the ~12 valid ones do useful work), and the per-user
// tables were filled by loops whose lanes spent a third of their instructions on index arithmetic.  Here a warp walks a contiguous
// range of users in passes:
//   1. window: the power rows of the next 8 users (one coalesced load each) -> per-user bit masks of the columns whose chain must
//      run (valid power; every column when an FoV mask is built, because the mask is defined for NaN-power columns too);
//      valid_mask and the default fov/clip entries of the other columns are written here;
//   2. as many whole users as fit the warp's table pool (`cap` paths) are taken, their (user, column) pairs listed densely;
//   3. chain rounds, lane = dense pair: the three float64 chains + combine, fov/clip masks, then -- still in that lane -- the path's
//      steering column A[0..M) (gain folded in, stored (re, re, im, im) for the FFMA2 operand modifiers) and its delay-phasor
//      seed levels go straight into the pool slot (user's first slot + rank among the user's contributing paths, in column order);
//   4. per user, lanes = column pairs: W entries from the seed levels (2-3 complex multiplies), 2 FFMA2 per complex MAC, rows in
//      registers, 512 contiguous bytes per row store.
// Every phasor that goes into a table has its phase reduced in float64 (phasor_cycles); W is a product of <= 3 of them.
// Pool rows are padded to an odd multiple of the store width so that the 32 lanes of a round never share a bank.
// =================================================================================================
constexpr int kS2Warps  = 4;
constexpr int kS2Window = 4;          // users examined per pass
constexpr int kS2MaxSeeds = 48;       // 16 + 16 + 16 (K <= 4096)

struct Small2Cfg {
    int warp_bytes;                   // bytes of a warp's shared-memory slice
    int off_A, off_W, off_list, off_meta;
    int cap;                          // table pool capacity in paths (>= n_cols)
    int window;                       // users examined per pass (<= kS2Window)
    int strideA;                      // bytes between pool rows of A: MT * 16 + 16
    int strideW;                      // float2 between pool rows of the seed levels: n_seed | 1
    int n0, log0, n1, n2;             // seed levels: column c -> L0[c & (n0-1)], L1[(c >> log0) & 15], L2[c >> 8] (n2 == 0: two levels)
    int users_per_warp;
    unsigned mul_mt, mul_bs0;
};

template <int MT, bool kL3>      // rows held in registers; three seed levels (K > 256) or two
__global__ void __launch_bounds__(kS2Warps * 32, MT <= 8 ? 5 : 4)      // <= 8 rows: 96 registers, 20 warps per SM (measured +10 % on cfg1)
fd_small2_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ Small2Cfg cfg)
{   // (occupancy is set by the host through the shared-memory slice per warp; the bound only caps registers at 128)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float4 s_coef[16];                 // element m -> (y_t, z_t, y_r, z_r) panel coordinates
    __shared__ double s_kseed[kS2MaxSeeds];       // seed entry e -> subcarrier offset whose delay phasor it holds
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_seed = cfg.n0 + cfg.n1 + cfg.n2;
    if (tid < 16) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < d.M) {
            const int r = tid / d.Mt, t = tid - r * d.Mt;
            c = make_float4((float)(t % d.bs0), (float)(t / d.bs0), (float)(r % d.ue0), (float)(r / d.ue0));
        }
        s_coef[tid] = c;
    }
    if (tid >= 32 && tid < 32 + n_seed) {
        const int e = tid - 32;
        double k;
        if (e < cfg.n0)               k = (double)d.subc_step * e;
        else if (e < cfg.n0 + cfg.n1) k = (double)d.subc_step * cfg.n0 * (e - cfg.n0) + (cfg.n2 == 0 ? (double)d.subc_start : 0.0);
        else                          k = (double)d.subc_step * 256.0 * (e - cfg.n0 - cfg.n1) + (double)d.subc_start;
        s_kseed[e] = k;
    }
    __syncthreads();                              // the only CTA-wide barrier: warps are independent from here on

    unsigned char* wsm = smem_raw + warp * cfg.warp_bytes;
    unsigned char* sA   = wsm + cfg.off_A;                                     // [cap][strideA]: MT x (re, re, im, im)
    float2* sW          = reinterpret_cast<float2*>(wsm + cfg.off_W);          // [cap][strideW]
    unsigned char* list = wsm + cfg.off_list;                                  // [cap] (user in window << 5) | column
    int* s_base         = reinterpret_cast<int*>(wsm + cfg.off_meta);          // [kS2Window + 1]
    int* s_cnt          = s_base + kS2Window + 1;                              // [kS2Window]
    unsigned* s_need    = reinterpret_cast<unsigned*>(s_cnt + kS2Window);      // [kS2Window] columns whose chain runs
    unsigned* s_valid   = s_need + kS2Window;                                  // [kS2Window] columns with a path (valid_mask)

    const long long u_begin = ((long long)blockIdx.x * kS2Warps + warp) * cfg.users_per_warp;
    const long long u_end = min(u_begin + (long long)cfg.users_per_warp, d.n_users);
    const unsigned ltmask = (1u << lane) - 1u;
    const int K = d.K, M = d.M, P0 = d.P0;
    const bool need_angles = prologue_needs_angles(d);
    const bool vec_ok = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(d.out) & 15) == 0);

    for (long long cur = u_begin; cur < u_end; ) {
        // ---- 1. window: which columns of the next users need their chain.  All loads first (independent, one latency), then the
        //         ballots; the per-user bit masks go to shared memory so that everything after this is a short rolled loop.
        const int n_in = (int)min((long long)cfg.window, u_end - cur);
        const long long prow = cur * (long long)d.ld + lane;
        float pw[kS2Window];                                                   // only the NaN-ness of the power is used here
        #pragma unroll
        for (int ul = 0; ul < kS2Window; ++ul) {
            pw[ul] = __int_as_float(0x7fc00000);
            if (ul < n_in && lane < P0)
                pw[ul] = d.in_f64 ? (float)__ldg(reinterpret_cast<const double*>(d.power) + prow + ul * d.ld) : __ldg(d.power + prow + ul * d.ld);
        }
        #pragma unroll
        for (int ul = 0; ul < kS2Window; ++ul) {
            const bool in = ul < n_in && lane < P0;
            const bool valid = in && lane < d.P && !(pw[ul] != pw[ul]);           // channel.py:260, dataset.py:258-261
            const unsigned vb = __ballot_sync(0xffffffffu, valid);
            const unsigned nb = d.fov_any ? __ballot_sync(0xffffffffu, in) : vb;
            if (lane == ul) { s_valid[ul] = vb; s_need[ul] = nb; }
        }
        __syncwarp();
        // ---- 2. take whole users while the pool has room; dense list of their (user, column) pairs; masks of the columns that
        //         run no chain (valid_mask of every column)
        int cum = 0, n_take = 0;
        long long o = cur * (long long)P0 + lane;
        #pragma unroll 1
        for (int ul = 0; ul < n_in; ++ul, o += P0) {
            const unsigned nb = s_need[ul], vb = s_valid[ul];
            const int c = __popc(nb);
            if (ul > 0 && cum + c > cfg.cap) break;
            if (lane == 0) { s_base[ul] = cum; s_cnt[ul] = 0; }
            const bool run = (nb >> lane) & 1u;
            if (run) list[cum + __popc(nb & ltmask)] = (unsigned char)((ul << 5) | lane);
            if (lane < P0) {
                if (d.valid_mask) d.valid_mask[o] = (vb >> lane) & 1u;
                if (!run) {                                                     // no FoV mask is built and the column has no power
                    if (d.fov_mask)  d.fov_mask[o] = 1;
                    if (d.clip_mask) d.clip_mask[o] = 0;
                }
            }
            cum += c;
            ++n_take;
        }
        __syncwarp();
        // rows of the users after this pass: pull them towards L2/L1 while the chains of this pass run (lane = (array, user))
        if (lane < 7 * 4) {
            const int arr = lane >> 2;
            const long long u = cur + n_take + (lane & 3);
            if (u < u_end) {
                const float* base = (arr == 0) ? d.power : (arr == 1) ? d.phase : (arr == 2) ? d.delay : (arr == 3) ? d.az[0] : (arr == 4) ? d.el[0]
                                  : (arr == 5) ? d.az[1] : d.el[1];
                asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(base) + u * (long long)d.ld * (d.in_f64 ? 8 : 4)));
            }
        }
        // ---- 3. chain rounds: lane = dense pair
        #pragma unroll 1
        for (int i0 = 0; i0 < cum; i0 += 32) {
            const int i = i0 + lane;
            const bool act = i < cum;
            int ul = kS2Window, col = 0;
            if (act) { const int e = list[i]; ul = e >> 5; col = e & 31; }
            const long long user = cur + ul;
            PathState st;
            st.contrib = false; st.valid = false; st.fov = true; st.over = false;
            if (act) {
                SideOut s0, s1; GainOut g;
                if (need_angles) { prologue_side<true>(d, user, col, 0, s0, d.Mt > 1);  prologue_side<true>(d, user, col, 1, s1, d.Mr > 1); }
                else             { prologue_side<false>(d, user, col, 0, s0, d.Mt > 1); prologue_side<false>(d, user, col, 1, s1, d.Mr > 1); }
                prologue_gain<true>(d, user, col, g);
                prologue_combine<true>(d, s0, s1, g, st);
                const long long o = user * (long long)P0 + col;
                if (d.fov_mask)  d.fov_mask[o]  = st.fov ? 1 : 0;
                if (d.clip_mask) d.clip_mask[o] = (st.valid && st.over) ? 1 : 0;
            }
            const bool contrib = act && st.contrib;
            const unsigned bc = __ballot_sync(0xffffffffu, contrib);
            const unsigned same = __match_any_sync(0xffffffffu, ul);             // lanes of the same user (idle lanes: key kS2Window)
            const int before = act ? s_cnt[ul] : 0;
            __syncwarp();
            if (act && lane == __ffs(same) - 1) s_cnt[ul] = before + __popc(bc & same);
            __syncwarp();
            if (contrib) {
                const int q = s_base[ul] + before + __popc(bc & same & ltmask);
                float4* a = reinterpret_cast<float4*>(sA + (size_t)q * cfg.strideA);
                #pragma unroll 4
                for (int m = 0; m < MT; ++m) {
                    float2 v = make_float2(0.f, 0.f);
                    if (m < M) {
                        const float4 cf = s_coef[m];
                        const double cyc = fma((double)cf.x, st.u[0], fma((double)cf.y, st.v[0], fma((double)cf.z, st.u[1], (double)cf.w * st.v[1])));
                        v = cmul(st.c, phasor_cycles_sfu(cyc));
                    }
                    a[m] = make_float4(v.x, v.x, v.y, v.y);
                }
                float2* w = sW + (size_t)q * cfg.strideW;
                #pragma unroll 4
                for (int e = 0; e < n_seed; ++e) w[e] = phasor_cycles_sfu(-(st.wcyc * s_kseed[e]));
            }
        }
        __syncwarp();
        // ---- 4. accumulate and store, one user at a time: lane owns columns 2l, 2l+1 of a 64-column pass and all rows
        #pragma unroll 1
        for (int ul = 0; ul < n_take; ++ul) {
            const long long user = cur + ul;
            const int np = s_cnt[ul];
            const unsigned char* aU = sA + (size_t)s_base[ul] * cfg.strideA;
            const float2* wU = sW + (size_t)s_base[ul] * cfg.strideW;
            float2* out_u = d.out + user * (long long)M * K;
            if (np == 0) {                                                      // zeros (channel.py:257,:269-271)
                const long long total = (long long)M * K;
                if (vec_ok) { float4* o = reinterpret_cast<float4*>(out_u); for (long long e = lane; e < total / 2; e += 32) __stcs(o + e, make_float4(0.f, 0.f, 0.f, 0.f)); }
                else        { for (long long e = lane; e < total; e += 32) __stcs(out_u + e, make_float2(0.f, 0.f)); }
                continue;
            }
            for (int col0 = 0; col0 < K; col0 += 64) {
                const int c0 = col0 + 2 * lane;
                const int cc = min(c0, K - 1);                                  // idle lanes read valid table entries
                const float2* w_lo = wU + (cc & (cfg.n0 - 1));
                const float2* w_m1 = wU + cfg.n0 + (kL3 ? ((cc >> 4) & 15) : (cc >> cfg.log0));
                const float2* w_m2 = wU + cfg.n0 + cfg.n1 + (cc >> 8);
                const unsigned char* ar_b = aU;
                float2 acc[MT][2];
                #pragma unroll
                for (int m = 0; m < MT; ++m) { acc[m][0] = make_float2(0.f, 0.f); acc[m][1] = make_float2(0.f, 0.f); }
                #pragma unroll 1
                for (int p = 0; p < np; ++p) {
                    float2 hi = *w_m1;
                    if (kL3) hi = cmul(hi, *w_m2);
                    const float2 w0 = cmul(hi, w_lo[0]);
                    const float2 w1 = cmul(hi, w_lo[1]);
                    const float2 w0s = make_float2(w0.y, w0.x), w1s = make_float2(w1.y, w1.x);
                    const float4* ar = reinterpret_cast<const float4*>(ar_b);
                    #pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const float4 a = ar[m];
                        const float2 a1 = make_float2(a.x, a.y), a2 = make_float2(-a.z, a.w);
                        acc[m][0] = __ffma2_rn(a1, w0, acc[m][0]);
                        acc[m][0] = __ffma2_rn(a2, w0s, acc[m][0]);
                        acc[m][1] = __ffma2_rn(a1, w1, acc[m][1]);
                        acc[m][1] = __ffma2_rn(a2, w1s, acc[m][1]);
                    }
                    w_lo += cfg.strideW; w_m1 += cfg.strideW; ar_b += cfg.strideA;
                    if (kL3) w_m2 += cfg.strideW;
                }
                float2* o = out_u + c0;
                if (vec_ok && col0 + 64 <= K && M == MT) {                      // whole pass inside the row, every register row is real
                    #pragma unroll
                    for (int m = 0; m < MT; ++m)
                        __stcs(reinterpret_cast<float4*>(o + (long long)m * K), make_float4(acc[m][0].x, acc[m][0].y, acc[m][1].x, acc[m][1].y));
                } else {
                    #pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        if (m >= M) break;
                        float2* om = o + (long long)m * K;
                        if (vec_ok && c0 + 1 < K) __stcs(reinterpret_cast<float4*>(om), make_float4(acc[m][0].x, acc[m][0].y, acc[m][1].x, acc[m][1].y));
                        else {
                            if (c0 < K)     __stcs(om, acc[m][0]);
                            if (c0 + 1 < K) __stcs(om + 1, acc[m][1]);
                        }
                    }
                }
            }
        }
        __syncwarp();                                                           // the pool is rewritten by the next pass
        cur += n_take;
    }
}

}  // namespace dmk
