// dmk_bf.cuh -- beam amplitude maps without ever writing H.
//
// The first consumer of H in the reference's manual (docs/manual.ipynb cell 105) applies a beamforming codebook
// F [n_beams, M_t] (dm.steering_vec, geometry.py:322-339) and averages the amplitude over RX antennas and subcarriers:
//     amp[u, b] = mean_{r, k} | sum_t F[b, t] H[u, r, t, k] |           (np.abs(F1 @ channel).mean(axis=1).mean(axis=-1))
// Because H is a rank-P sum, the codebook folds into the TX side of every path,
//     sum_t F[b,t] H[u,r,t,k] = sum_p (c_p a_rx[r,p] G[b,p]) W[p,k],     G[b,p] = sum_t F[b,t] a_tx[t,p],
// which is the same contraction as the channel itself with the M_t antenna rows replaced by n_beams beam rows -- and
// the [n, M_r, n_beams, K] result is reduced to [n, n_beams] in registers, so the kernel writes 4 bytes per (user, beam)
// instead of 8 M_r M_t K bytes per user: the HBM-write bound of the channel kernels disappears.
// One CTA per user; same float64 prologue, same float64-reduced phasors, FP32 accumulation.
#pragma once
#include "dmk_fd.cuh"

namespace dmk {

struct BfCfg {
    int n_beams;
    int off_G, off_rows, off_tY, off_tZ, off_aR;      // byte offsets into dynamic shared memory (after the W and A tiles)
    const float2* F;                                   // DEVICE [n_beams, Mt] complex64, row-major
    float* out;                                        // DEVICE [n_users, n_beams] float32
};

__global__ void __launch_bounds__(kFdThreads, 2)
bf_kernel(const __grid_constant__ DevDesc d, const __grid_constant__ BfCfg c)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sW = reinterpret_cast<float2*>(smem_raw);                    // [kMaxPaths][kTK]
    float2* sA = sW + kMaxPaths * kTK;                                   // [kMaxPaths][kTM]
    unsigned char* ext = smem_raw + (size_t)kMaxPaths * (kTK + kTM) * sizeof(float2);
    float2* G   = reinterpret_cast<float2*>(ext + c.off_G);             // [n_beams][kMaxPaths]
    float*  rows = reinterpret_cast<float*>(ext + c.off_rows);          // [Mr * n_beams] sum over columns of |Y|
    float2* tY  = reinterpret_cast<float2*>(ext + c.off_tY);            // [np][bs0]
    float2* tZ  = reinterpret_cast<float2*>(ext + c.off_tZ);            // [np][bs1]
    float2* aR  = reinterpret_cast<float2*>(ext + c.off_aR);            // [np][Mr], path gain folded in
    __shared__ FdShared sh;
    __shared__ PrologueScratch psc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long user = blockIdx.x;
    const int B = c.n_beams;
    float* out_u = c.out + user * (long long)B;

    fd_cta_prologue(d, user, sh, psc, true);
    const int np = sh.np;
    if (np == 0) {                                                       // no contributing path: H == 0 -> amplitude 0
        for (int b = tid; b < B; b += kFdThreads) out_u[b] = 0.f;
        return;
    }
    const int bs0 = d.bs0, bs1 = d.bs1, Mr = d.Mr, Mt = d.Mt;
    for (int e = tid; e < np * bs0; e += kFdThreads) { const int p = e / bs0, y = e - p * bs0; tY[e] = phasor_cycles((double)y * sh.u[0][p]); }
    for (int e = tid; e < np * bs1; e += kFdThreads) { const int p = e / bs1, z = e - p * bs1; tZ[e] = phasor_cycles((double)z * sh.v[0][p]); }
    for (int e = tid; e < np * Mr; e += kFdThreads) {
        const int p = e / Mr, r = e - p * Mr;
        const int yr = r % d.ue0, zr = r / d.ue0;
        aR[e] = cmul(sh.c[p], phasor_cycles((double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
    }
    const int n_rows = Mr * B;
    for (int e = tid; e < n_rows; e += kFdThreads) rows[e] = 0.f;
    __syncthreads();
    // ---- beam response of every path: G[b][p] = sum_t F[b,t] a_tx[t,p]   (a_tx[t,p] = tY[p][t % bs0] tZ[p][t / bs0])
    for (int e = tid; e < B * np; e += kFdThreads) {
        const int b = e / np, p = e - b * np;
        const float2* f = c.F + (long long)b * Mt;
        float2 g = make_float2(0.f, 0.f);
        for (int z = 0; z < bs1; ++z) {
            float2 gz = make_float2(0.f, 0.f);
            for (int y = 0; y < bs0; ++y) {
                const float2 fv = __ldg(f + z * bs0 + y), ty = tY[p * bs0 + y];
                gz.x = fmaf(fv.x, ty.x, gz.x); gz.x = fmaf(-fv.y, ty.y, gz.x);
                gz.y = fmaf(fv.x, ty.y, gz.y); gz.y = fmaf(fv.y, ty.x, gz.y);
            }
            const float2 tz = tZ[p * bs1 + z];
            g.x = fmaf(gz.x, tz.x, g.x); g.x = fmaf(-gz.y, tz.y, g.x);
            g.y = fmaf(gz.x, tz.y, g.y); g.y = fmaf(gz.y, tz.x, g.y);
        }
        G[b * kMaxPaths + p] = g;
    }

    const int ncols = d.K;
    const int n_ct = (ncols + kTK - 1) / kTK;
    const int n_rt = (n_rows + kTM - 1) / kTM;
    for (int ct = 0; ct < n_ct; ++ct) {
        const int col0 = ct * kTK;
        __syncthreads();                                   // G complete (first pass) / previous tile's readers are done with sW
        for (int e = tid; e < np * kTK; e += kFdThreads) {
            const int p = e / kTK, cidx = e % kTK;
            const int col = col0 + cidx;
            float2 w = make_float2(0.f, 0.f);
            if (col < ncols) w = phasor_cycles(-(sh.wcyc[p] * (double)subcarrier_at(d, col)));
            sW[p * kTK + cidx] = w;
        }
        for (int rt = 0; rt < n_rt; ++rt) {
            const int row0 = rt * kTM;
            if (rt > 0) __syncthreads();                   // previous row tile's readers are done with sA
            // ---- A tile: rows are (rx element r, beam b): c_p a_rx[r,p] G[b,p]
            for (int e = tid; e < np * kTM; e += kFdThreads) {
                const int p = e / kTM, r = e % kTM;
                const int m = row0 + r;
                float2 a = make_float2(0.f, 0.f);
                if (m < n_rows) {
                    const int rr = m / B, b = m - rr * B;
                    a = cmul(aR[p * Mr + rr], G[b * kMaxPaths + p]);
                }
                sA[p * kTM + r] = a;
            }
            __syncthreads();

            float2 acc[8][4];
            #pragma unroll
            for (int i = 0; i < 8; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
            tile_rank_update(sW, sA, np, lane, warp, acc);
            // ---- |Y| summed over this tile's columns; each warp owns its 8 rows, so no atomics
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                float s = 0.f;
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = col0 + (j >> 1) * 64 + 2 * lane + (j & 1);
                    if (col < ncols) s += sqrtf(fmaf(acc[i][j].x, acc[i][j].x, acc[i][j].y * acc[i][j].y));
                }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const int m = row0 + warp * 8 + i;
                if (lane == 0 && m < n_rows) rows[m] += s;
            }
        }
    }
    __syncthreads();
    const float inv = 1.0f / ((float)Mr * (float)ncols);
    for (int b = tid; b < B; b += kFdThreads) {
        float s = 0.f;
        for (int r = 0; r < Mr; ++r) s += rows[r * B + b];
        out_u[b] = s * inv;
    }
}

}  // namespace dmk
