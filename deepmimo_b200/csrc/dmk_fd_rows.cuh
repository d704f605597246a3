// dmk_fd_rows.cuh -- FD channel kernel for a handful of selected subcarriers (K <= 8): one WARP per user, lanes = antenna rows.
//
// The reference's DEFAULT parameters select ONE subcarrier (ofdm.selected_subcarriers = [0], channel.py:61): a user's output is M
// complex values.  Every other FD kernel spreads the K columns over lanes (fd_fast_kernel: 256 column lanes per CTA, one of them
// busy: 54 ms for 200 k users of a 32x8 x 2x2 panel) or over MMA columns (fd_mma_kernel: 15 of 16 wasted).  Here
//   lanes = path columns: the float64 chains + combine, masks; the contributing paths are compacted into a per-warp table
//     (gain, steering cycles, delay phasors of the K subcarriers -- any subcarrier list, affine or not);
//   lanes = antenna rows, 32 at a time: per path one float64-reduced SFU phasor of the row's steering phase and a complex multiply
//     by the gain, then K complex multiply-adds with the path's delay phasors; the warp stores 32 K contiguous complex values.
// Arithmetic per (32 rows, path): ~16 + 5 K instructions; no tensor cores -- there is no contraction worth a tile.
#pragma once
#include "dmk_fd.cuh"

namespace dmk {

constexpr int kRowsWarps = 4;

struct RowsCfg { unsigned mul_mt, mul_bs0, mul_ue0; };       // ceil(2^32 / d) for d > 1, else 0: exact division while m * d < 2^32

struct RowsPath { float2 c; double u0, v0, u1, v1; };       // 40 bytes

template <int KM>      // accumulators per row: the smallest of 1, 2, 4, 8 that holds K
__global__ void __launch_bounds__(kRowsWarps * 32)
fd_rows_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ RowsCfg cfg)
{
    __shared__ RowsPath s_path[kRowsWarps][kMaxPaths];
    __shared__ float2 s_w[kRowsWarps][kMaxPaths][KM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long user = (long long)blockIdx.x * kRowsWarps + warp;
    if (user >= d.n_users) return;                          // warp-uniform; warps never synchronise with each other
    RowsPath* tp = s_path[warp];
    float2 (*tw)[KM] = s_w[warp];
    const int K = d.K;

    // ---- chains, lane = path column (a single-element unrotated side -- the reference's default UE -- skips its chain)
    const bool triv0 = side_angles_trivial(d, 0), triv1 = side_angles_trivial(d, 1);
    PathState st;
    const bool active = lane < d.P0;
    st.contrib = false; st.valid = false; st.fov = true; st.over = false;
    st.c = make_float2(0.f, 0.f); st.wcyc = 0.0; st.u[0] = st.u[1] = st.v[0] = st.v[1] = 0.0;
    if (active) {
        SideOut s0, s1; GainOut g;
        if (prologue_needs_angles(d)) { prologue_side<true>(d, user, lane, 0, s0, d.Mt > 1);  prologue_side<true>(d, user, lane, 1, s1, d.Mr > 1); }
        else                          { prologue_side_auto<false>(d, user, lane, 0, s0, d.Mt > 1, triv0); prologue_side_auto<false>(d, user, lane, 1, s1, d.Mr > 1, triv1); }
        prologue_gain<true>(d, user, lane, g);
        prologue_combine<true>(d, s0, s1, g, st);
        const long long o = user * (long long)d.P0 + lane;
        if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
        if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
        if (d.clip_mask)  d.clip_mask[o]  = (st.valid && st.over) ? 1 : 0;
    }
    const bool contrib = active && st.contrib;
    const unsigned cb = __ballot_sync(0xffffffffu, contrib);
    const int np = __popc(cb);
    if (contrib) {
        const int j = __popc(cb & ((1u << lane) - 1u));
        tp[j].c = st.c; tp[j].u0 = st.u[0]; tp[j].v0 = st.v[0]; tp[j].u1 = st.u[1]; tp[j].v1 = st.v[1];
        #pragma unroll
        for (int k = 0; k < KM; ++k)                         // delay phasors of the selected subcarriers (channel.py:184-186)
            tw[j][k] = (k < K) ? phasor_cycles(-(st.wcyc * (double)subcarrier_at(d, k))) : make_float2(0.f, 0.f);
    }
    __syncwarp();

    // ---- rows, lane = antenna row
    float2* out_u = d.out + user * (long long)d.M * K;
    for (int m0 = 0; m0 < d.M; m0 += 32) {
        const unsigned m = (unsigned)min(m0 + lane, d.M - 1);
        const unsigned rr = cfg.mul_mt ? __umulhi(m, cfg.mul_mt) : (d.Mt > 1 ? 0u : m);
        const unsigned t = m - rr * (unsigned)d.Mt;
        const unsigned zt = cfg.mul_bs0 ? __umulhi(t, cfg.mul_bs0) : (d.bs0 > 1 ? 0u : t);
        const unsigned yt = t - zt * (unsigned)d.bs0;
        const unsigned zr = cfg.mul_ue0 ? __umulhi(rr, cfg.mul_ue0) : (d.ue0 > 1 ? 0u : rr);
        const unsigned yr = rr - zr * (unsigned)d.ue0;
        const double fyt = (double)yt, fzt = (double)zt, fyr = (double)yr, fzr = (double)zr;
        float2 acc[KM];
        #pragma unroll
        for (int k = 0; k < KM; ++k) acc[k] = make_float2(0.f, 0.f);
        #pragma unroll 2
        for (int p = 0; p < np; ++p) {
            const RowsPath q = tp[p];                         // broadcast reads
            const double cyc = fma(fyt, q.u0, fma(fzt, q.v0, fma(fyr, q.u1, fzr * q.v1)));
            const float2 a = cmul(q.c, phasor_cycles_sfu(cyc));
            #pragma unroll
            for (int k = 0; k < KM; ++k) {
                const float2 w = tw[p][k];
                acc[k].x = fmaf(a.x, w.x, fmaf(-a.y, w.y, acc[k].x));
                acc[k].y = fmaf(a.x, w.y, fmaf(a.y, w.x, acc[k].y));
            }
        }
        if (m0 + lane < d.M) {
            float2* o = out_u + (long long)(m0 + lane) * K;
            #pragma unroll
            for (int k = 0; k < KM; ++k)
                if (k < K) __stcs(o + k, acc[k]);
        }
    }
}

}  // namespace dmk
