// dmk_fd.cuh -- fused frequency-domain channel kernel.
//
//   H[u, m, q] = sum_p A[m,p] * W[p,q]          m = r*Mt + t  (M rows), q = k_idx*T + it (K*T columns)
//   A[m,p] = c_p * exp(j 2 pi (y_t u_t + z_t v_t + y_r u_r + z_r v_r))      (steering, gain folded in)
//   W[p,q] = exp(j 2 pi (f_D[p] t_it - k delay_n[p] / N))                   (channel.py:196-197)
//
// One CTA owns one user (or a slice of its column tiles when there are few users).  Warp 0 runs
// the per-path prologue and compacts the contributing paths into shared memory; the CTA then
// walks 64 x 128 output tiles: W tile and A tile are built in shared memory with the phase
// argument reduced in float64, and each thread accumulates an 8 x 4 complex register tile over
// the (compacted) paths with FP32 FMAs.  Output is written once, coalesced, streaming.
#pragma once
#include "dmk_prologue.cuh"

namespace dmk {

constexpr int kFdThreads = 256;
constexpr int kTM = 64;      // rows per CTA tile  (8 warps x 8 rows)
constexpr int kTK = 128;     // columns per CTA tile (32 lanes x 4 columns)

#ifdef DMK_TC_TRACE
__device__ long long g_pro_trace[16];
#endif

struct FdShared {
    float2 c[kMaxPaths];
    double wcyc[kMaxPaths];
    double fd[kMaxPaths];
    double u[2][kMaxPaths];
    double v[2][kMaxPaths];
    int    np;
};

// Prologue of one user, all threads of the CTA call this: warps 0-2 run the three float64 chains of every path
// column (cta_prologue_chains), then warp 0 combines them, writes the masks and compacts the contributing paths.
// Ends with a CTA barrier; sh.np is valid afterwards.
__device__ __forceinline__ void fd_cta_prologue(const DevDesc& d, long long user, FdShared& sh, PrologueScratch& sc, bool write_masks)
{
#ifdef DMK_TC_TRACE
    long long pt0 = clock64();
#endif
    cta_prologue_chains<true>(d, user, sc);
    if (threadIdx.x >= 224) prefetch_user_rows_shifted(d, user);
#ifdef DMK_TC_TRACE
    long long pt1 = clock64();
    if (blockIdx.x == gridDim.x / 2 && (threadIdx.x & 31) == 3 && threadIdx.x < 96) { g_pro_trace[(threadIdx.x >> 5) * 2] = pt0; g_pro_trace[(threadIdx.x >> 5) * 2 + 1] = pt1; }
#endif
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        PathState st;
        const bool active = lane < d.P0;
        st.contrib = false; st.valid = false; st.fov = true; st.over = false;
#ifdef DMK_TC_TRACE
        if (blockIdx.x == gridDim.x / 2 && lane == 3) g_pro_trace[6] = clock64();
#endif
        if (active) prologue_combine<true>(d, sc.side[0][lane], sc.side[1][lane], sc.gain[lane], st);
        const unsigned ballot = __ballot_sync(0xffffffffu, active && st.contrib);
        if (active && st.contrib) {
            const int j = __popc(ballot & ((1u << lane) - 1u));
            sh.c[j] = st.c; sh.wcyc[j] = st.wcyc; sh.fd[j] = st.fd;
            sh.u[0][j] = st.u[0]; sh.v[0][j] = st.v[0];
            sh.u[1][j] = st.u[1]; sh.v[1][j] = st.v[1];
        }
        if (lane == 0) sh.np = __popc(ballot);
        if (write_masks && active) {
            const long long o = user * (long long)d.P0 + lane;
            if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
            if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
            if (d.clip_mask)  d.clip_mask[o]  = (st.valid && st.over) ? 1 : 0;
        }
#ifdef DMK_TC_TRACE
        if (blockIdx.x == gridDim.x / 2 && lane == 3) g_pro_trace[7] = clock64();
#endif
    }
    __syncthreads();
}

// Out-of-line instance for the tensor-core kernel: inlining the three float64 chains there raises its register
// allocation (93 -> 112) and costs ~10 % in its main loop (measured A/B on the same B200), whereas the packed-FP32
// kernel is ~6 % faster with the inlined version.
__device__ __noinline__ void fd_cta_prologue_outlined(const DevDesc& d, long long user, FdShared& sh, PrologueScratch& sc, bool write_masks)
{
    fd_cta_prologue(d, user, sh, sc, write_masks);
}

// Rank-np update of a thread's 8 x 4 register tile from the W tile [np][kTK] and the A tile [np][kTM] in shared memory (generic
// tile kernels): four scalar FFMA per complex MAC.  The packed form (two FFMA2 with (re, re) / (-im, im) operands built in registers
// from this A layout) was measured 20 % slower here (1.71 -> 2.08 ms on cfg2 x 1024 users): the operand shuffles cost more than the
// FFMAs they save; fd_fast_kernel pays for a pre-duplicated A strip in shared memory instead.
__device__ __forceinline__ void tile_rank_update(const float2* sW, const float2* sA, int np, int lane, int warp, float2 (&acc)[8][4])
{
    const float4* wrow = reinterpret_cast<const float4*>(sW) + lane;          // columns 2*lane, 2*lane+1 (and +64)
    const float4* arow = reinterpret_cast<const float4*>(sA) + warp * 4;      // rows warp*8 .. +7
    #pragma unroll 1
    for (int p = 0; p < np; ++p) {
        const float4 w01 = wrow[p * (kTK / 2)];
        const float4 w23 = wrow[p * (kTK / 2) + 32];
        const float2 w[4] = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w), make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 t = arow[p * (kTM / 2) + i];                          // rows 2i, 2i+1: (re, im, re, im)
            const float2 a[2] = {make_float2(t.x, t.y), make_float2(t.z, t.w)};
            #pragma unroll
            for (int h = 0; h < 2; ++h)
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2& c = acc[2 * i + h][j];
                    c.x = fmaf(a[h].x, w[j].x, c.x); c.x = fmaf(-a[h].y, w[j].y, c.x);
                    c.y = fmaf(a[h].x, w[j].y, c.y); c.y = fmaf(a[h].y, w[j].x, c.y);
                }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Receive low-pass filter (ofdm.rx_filter = 1; channel.py:166-168, :193-194).  The per-path frequency response is no
// longer a pure phasor but the N-point DFT of the sampled sinc,
//     W[p, k] = sum_{d=0}^{N-1} sinc(d - tau_p) exp(-j 2 pi d k / N),        tau_p = delay_n[p]
// (the reference forms the [P, N] sinc matrix and multiplies it with the [N, K] DFT matrix in complex128).
// sinc(d - tau) = sin(pi tau) (-1)^d / (pi (tau - d)): sin(pi tau) is evaluated once per path in float64 from the
// fractional part of tau, so every sample costs one float64 subtraction, one conversion and one float32 division.
// Per column tile and batch of `lpf_batch` paths the CTA fills x[b][d] in shared memory and transforms it:
//   N a power of two -> in-place radix-2 FFT (bit-reversed fill, log2 N barrier-separated stages, exact twiddle table);
//   otherwise        -> direct DFT of the selected columns only (twiddle index walks d*k mod N).
// The [np][128] tile of W then feeds the same rank-np accumulation as the unfiltered path.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void lpf_w_tile(const DevDesc& d, const FdShared& sh, int np, int col0, int ncols, float2* sW,
                                           float2* lpf_mem, bool first_tile)
{
    // Users whose transformed paths fit (np * K <= d.lpf_cache complex values, no time axis) are transformed once, on the first
    // tile, into Wc[p][K]; later tiles only copy their 128 columns.  Users with more paths redo the transforms per tile.
    const bool cached = !d.has_time_axis && (long long)np * d.K <= (long long)d.lpf_cache;
    float2* Wc = lpf_mem + (size_t)(1 + d.lpf_batch) * d.N;
    if (cached && !first_tile) {
        for (int e = threadIdx.x; e < np * kTK; e += kFdThreads) {
            const int p = e / kTK, cidx = e % kTK;
            const int col = col0 + cidx;
            sW[p * kTK + cidx] = col < ncols ? Wc[p * d.K + col] : make_float2(0.f, 0.f);
        }
        return;
    }
    __shared__ float  lpf_a[kMaxPaths];      // (-1)^k sin(pi r) / pi,  tau = k + r
    __shared__ float  lpf_sk[kMaxPaths];     // sinc(r): the sample d == k
    __shared__ int    lpf_k[kMaxPaths];
    __shared__ double lpf_tau[kMaxPaths];
    const int tid = threadIdx.x;
    const int N = d.N, B = d.lpf_batch, lg = d.lpf_log2n;
    float2* tw = lpf_mem;                    // [N]   exp(-j 2 pi i / N)
    float2* x  = lpf_mem + N;                // [B][N]
    if (first_tile) {
        for (int i = tid; i < N; i += kFdThreads) tw[i] = phasor_cycles(-((double)i * d.inv_n));
        if (tid < np) {
            const double tau = sh.wcyc[tid] * (double)N;           // delay_n (wcyc = delay_n / N)
            const double k = rint(tau), r = tau - k;
            const double sp = sinpi(r);
            const long long ki = (long long)k;
            lpf_tau[tid] = tau;
            lpf_k[tid] = (k >= 0.0 && k < (double)N) ? (int)ki : -1;
            lpf_a[tid] = (float)(((ki & 1) ? -sp : sp) / kPi);
            lpf_sk[tid] = (r == 0.0) ? 1.0f : (float)(sp / (kPi * r));      // np.sinc: sin(pi x)/(pi x), 1 at x == 0
        }
    }
    __syncthreads();
    for (int p0 = 0; p0 < np; p0 += B) {
        const int nb = min(B, np - p0);
        if (p0 > 0) __syncthreads();                               // the previous batch's gather is done with x
        // ---- samples of the shifted sinc (bit-reversed order for the FFT route)
        for (int e = tid; e < nb * N; e += kFdThreads) {
            const int b = lg >= 0 ? (e >> lg) : e / N;
            const int dd = e - b * N;
            const int p = p0 + b;
            float s;
            if (dd == lpf_k[p]) s = lpf_sk[p];
            else {
                const float den = (float)(lpf_tau[p] - (double)dd);
                s = __fdiv_rn((dd & 1) ? -lpf_a[p] : lpf_a[p], den);
            }
            const int pos = lg > 0 ? (int)(__brev((unsigned)dd) >> (32 - lg)) : dd;
            x[b * N + pos] = make_float2(s, 0.f);
        }
        __syncthreads();
        if (lg >= 0) {
            // ---- in-place radix-2 decimation-in-time FFT of nb sequences
            const int half_n = N >> 1;
            for (int lh = 0; lh < lg; ++lh) {
                const int half = 1 << lh;
                for (int t = tid; t < nb * half_n; t += kFdThreads) {
                    const int b = t >> (lg - 1), j = t & (half_n - 1);
                    const int pos = j & (half - 1);
                    const int i0 = ((j >> lh) << (lh + 1)) + pos;
                    float2* xb = x + b * N;
                    const float2 w = tw[pos << (lg - 1 - lh)];
                    const float2 a = xb[i0], c = cmul(xb[i0 + half], w);
                    xb[i0] = make_float2(a.x + c.x, a.y + c.y);
                    xb[i0 + half] = make_float2(a.x - c.x, a.y - c.y);
                }
                __syncthreads();
            }
        }
        if (cached) {
            // ---- every selected column of the batch's paths goes to the per-user cache
            const int K = d.K;
            for (int e = tid; e < nb * K; e += kFdThreads) {
                const int b = e / K, ki = e - b * K;
                int k = subcarrier_at(d, ki) % N;
                if (k < 0) k += N;
                float2 w;
                if (lg >= 0) w = x[b * N + k];
                else {
                    const float2* xb = x + b * N;
                    float wr = 0.f, wi = 0.f;
                    int idx = 0;
                    for (int dd = 0; dd < N; ++dd) {
                        const float sv = xb[dd].x;
                        const float2 t = tw[idx];
                        wr = fmaf(sv, t.x, wr); wi = fmaf(sv, t.y, wi);
                        idx += k; if (idx >= N) idx -= N;
                    }
                    w = make_float2(wr, wi);
                }
                Wc[(p0 + b) * K + ki] = w;
            }
            continue;
        }
        // ---- the tile's columns
        for (int e = tid; e < nb * kTK; e += kFdThreads) {
            const int b = e / kTK, cidx = e % kTK;
            const int col = col0 + cidx;
            const int p = p0 + b;
            float2 w = make_float2(0.f, 0.f);
            if (col < ncols) {
                const int ki = col / d.T, it = col - ki * d.T;
                int k = subcarrier_at(d, ki) % N;
                if (k < 0) k += N;
                if (lg >= 0) {
                    w = x[b * N + k];
                } else {
                    const float2* xb = x + b * N;
                    float wr = 0.f, wi = 0.f;
                    int idx = 0;
                    for (int dd = 0; dd < N; ++dd) {
                        const float s = xb[dd].x;
                        const float2 t = tw[idx];
                        wr = fmaf(s, t.x, wr); wi = fmaf(s, t.y, wi);
                        idx += k; if (idx >= N) idx -= N;
                    }
                    w = make_float2(wr, wi);
                }
                if (d.has_time_axis) w = cmul(w, phasor_cycles(sh.fd[p] * d.times[it]));
            }
            sW[p * kTK + cidx] = w;
        }
    }
    if (cached) {
        __syncthreads();
        for (int e = tid; e < np * kTK; e += kFdThreads) {
            const int p = e / kTK, cidx = e % kTK;
            const int col = col0 + cidx;
            sW[p * kTK + cidx] = col < ncols ? Wc[p * d.K + col] : make_float2(0.f, 0.f);
        }
    }
}

__global__ void __launch_bounds__(kFdThreads, 2)
fd_tile_kernel(const __grid_constant__ DevDesc d, const int ksplit)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sW = reinterpret_cast<float2*>(smem_raw);                    // [kMaxPaths][kTK]
    float2* sA = sW + kMaxPaths * kTK;                                   // [kMaxPaths][kTM]
    __shared__ FdShared sh;
    __shared__ PrologueScratch psc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long user = blockIdx.x / ksplit;
    const int ks = blockIdx.x % ksplit;

    fd_cta_prologue(d, user, sh, psc, ks == 0);
    const int np = sh.np;

    const int ncols = d.K * d.T;
    const int n_ct = (ncols + kTK - 1) / kTK;
    const int n_rt = (d.M + kTM - 1) / kTM;
    float2* out_u = d.out + user * (long long)d.M * ncols;
    const bool vec_ok = ((ncols & 1) == 0) && ((reinterpret_cast<uintptr_t>(d.out) & 15) == 0);

    for (int ct = ks; ct < n_ct; ct += ksplit) {
        const int col0 = ct * kTK;
        __syncthreads();                                   // previous tile's readers are done with sW
        if (d.rx_filter) {
            lpf_w_tile(d, sh, np, col0, ncols, sW, reinterpret_cast<float2*>(smem_raw + (size_t)kMaxPaths * (kTK + kTM) * sizeof(float2)), ct == ks);
        } else
        // ---- W tile: per-(path, column) phasor, phase reduced in float64
        for (int e = tid; e < np * kTK; e += kFdThreads) {
            const int p = e / kTK, cidx = e % kTK;
            const int col = col0 + cidx;
            float2 w = make_float2(0.f, 0.f);
            if (col < ncols) {
                const int ki = col / d.T, it = col - ki * d.T;
                const double kval = (double)subcarrier_at(d, ki);
                double cyc = -(sh.wcyc[p] * kval);
                if (d.has_time_axis) cyc += sh.fd[p] * d.times[it];
                w = phasor_cycles(cyc);
            }
            sW[p * kTK + cidx] = w;
        }
        for (int rt = 0; rt < n_rt; ++rt) {
            const int row0 = rt * kTM;
            if (rt > 0) __syncthreads();                   // previous row tile's readers are done with sA
            // ---- A tile: gain * TX steering * RX steering, one phasor per (row, path)
            for (int e = tid; e < np * kTM; e += kFdThreads) {
                const int p = e / kTM, r = e % kTM;
                const int m = row0 + r;
                float2 a = make_float2(0.f, 0.f);
                if (m < d.M) {
                    const int rr = m / d.Mt, t = m - rr * d.Mt;
                    const int yt = t % d.bs0, zt = t / d.bs0;
                    const int yr = rr % d.ue0, zr = rr / d.ue0;
                    const double cyc = (double)yt * sh.u[0][p] + (double)zt * sh.v[0][p]
                                     + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p];
                    a = cmul(sh.c[p], phasor_cycles(cyc));
                }
                sA[p * kTM + r] = a;
            }
            __syncthreads();

            // ---- rank-np accumulation, 8 rows x 4 columns per thread
            float2 acc[8][4];
            #pragma unroll
            for (int i = 0; i < 8; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

            tile_rank_update(sW, sA, np, lane, warp, acc);

            // ---- store: each lane owns columns {2l, 2l+1} and {64+2l, 64+2l+1} of the tile
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = row0 + warp * 8 + i;
                if (m >= d.M) break;
                float2* orow = out_u + (long long)m * ncols + col0;
                #pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int c = h * 64 + 2 * lane;
                    if (vec_ok && col0 + c + 1 < ncols) {
                        __stcs(reinterpret_cast<float4*>(orow + c),
                               make_float4(acc[i][2 * h].x, acc[i][2 * h].y, acc[i][2 * h + 1].x, acc[i][2 * h + 1].y));
                    } else {
                        if (col0 + c < ncols)     __stcs(orow + c, acc[i][2 * h]);
                        if (col0 + c + 1 < ncols) __stcs(orow + c + 1, acc[i][2 * h + 1]);
                    }
                }
            }
        }
    }
}


// =================================================================================================
// fd_fast_kernel -- the production FD kernel (affine subcarrier selection, no time axis).
//
// Same mathematics as fd_tile_kernel; what changes is how often transcendental work is redone and
// how the FP32 pipe is fed:
//   * per user, once: separable phasor tables in shared memory
//        tY[p][y]            TX steering along the panel's y axis
//        tQ[p][r*bs1 + z]    RX steering of element r  x  TX steering along z  x  path gain c_p
//        wA[p][a], wB[p][b]  delay phasors: column q = 16a + b  ->  W[p,q] = wA[p][a] * wB[p][b]
//     every table entry has its phase reduced in float64 (phasor_cycles); tile entries are products
//     of 2-3 unit-modulus float32 phasors (relative error ~1e-7 each).
//   * W tile [np][256] rebuilt per column tile by the whole CTA (one complex multiply per entry);
//     A strip [np][8 rows] rebuilt per row tile by each warp for its own rows and used for two
//     128-column passes -> no CTA barrier inside the row loop, only __syncwarp.
//   * inner loop on packed FP32, 2 FFMA2 per complex MAC:
//        acc(re,im) += (a.re, a.re) * (w.re, w.im);   acc += (-a.im, a.im) * (w.im, w.re)
//     A entries are stored as (re, re, im, im) so that the second operand pair is a plain register
//     pair with the per-half negate modifier and W needs only the free .LO_HI swap: per path
//     64 FFMA2 + 10 LDS.128 issue slots for 128 FMA-pipe cycles, no fix-up instructions.
// =================================================================================================
constexpr int kTKW = 256;    // W tile width of the fast kernel (two passes of kTK columns)

struct FastCfg {
    int off_W, off_A, off_tY, off_tQ, off_wA, off_wB;   // byte offsets into dynamic smem
    int nA;                                              // ceil(K / 16)
    int pcap;                                            // path capacity the layout was sized for
    unsigned mul_mt, mul_bs0;                            // ceil(2^32 / Mt), ceil(2^32 / bs0): exact division for m < 2^32 / d
};

__device__ __forceinline__ float4 ldsA(const float4* p) { return *p; }

// Beam mode (row f3, dmk_beam_amplitude_fd): the same contraction with the M_t antenna rows replaced by n_beams beam rows.
// The host hands the kernel a descriptor whose TX panel is the "virtual array" 1 x n_beams (bs0 = 1, bs1 = Mt = n_beams), the
// table tQ[p][r * n_beams + b] = c_p a_rx[r,p] G[b,p] with G[b,p] = sum_t F[b,t] a_tx[t,p] is built from the true panel, and
// the epilogue reduces |acc| over the columns instead of storing it: 4 bytes per (user, beam) leave the SM.
struct BeamCfg {
    int n_beams, bs0, bs1;            // codebook rows; the true TX panel
    int off_rows;                     // byte offset of rows[Mr * n_beams] (float) in dynamic shared memory
    const float2* F;                  // DEVICE [n_beams, bs0 * bs1] complex64
    float* out;                       // DEVICE [n_users, n_beams] float32
};

template <bool kBeams>
__device__ __forceinline__ void fd_fast_body(const DevDesc& d, const FastCfg& cfg, const int ksplit, const BeamCfg& bf)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sW = reinterpret_cast<float2*>(smem_raw + cfg.off_W);    // [pcap][kTKW]
    float4* sA = reinterpret_cast<float4*>(smem_raw + cfg.off_A);    // [8 warps][pcap][8] (re,re,im,im)
    float2* tY = reinterpret_cast<float2*>(smem_raw + cfg.off_tY);   // [pcap][bs0]
    float2* tQ = reinterpret_cast<float2*>(smem_raw + cfg.off_tQ);   // [pcap][Mr*bs1]
    float2* wA = reinterpret_cast<float2*>(smem_raw + cfg.off_wA);   // [pcap][nA]
    float2* wB = reinterpret_cast<float2*>(smem_raw + cfg.off_wB);   // [pcap][16]
    __shared__ FdShared sh;
    __shared__ PrologueScratch psc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long user = blockIdx.x / ksplit;
    const int ks = blockIdx.x % ksplit;

    fd_cta_prologue(d, user, sh, psc, ks == 0);
    const int np = sh.np;
    const int ncols = d.K;
    const int n_ct = (ncols + kTKW - 1) / kTKW;
    const int n_rt = (d.M + kTM - 1) / kTM;
    float2* out_u = kBeams ? nullptr : d.out + user * (long long)d.M * ncols;
    const bool vec_ok = ((ncols & 1) == 0) && ((reinterpret_cast<uintptr_t>(d.out) & 15) == 0);
    const int nq = d.Mr * d.bs1;
    float* rows = reinterpret_cast<float*>(smem_raw + bf.off_rows);     // beam mode: sum over columns of |Y| per (rx element, beam)
    if (kBeams && np == 0) {                                            // no contributing path: H == 0 -> amplitude 0
        for (int b = tid; b < bf.n_beams; b += kFdThreads) bf.out[user * (long long)bf.n_beams + b] = 0.f;
        return;
    }

    // ---- per-user tables (phase reduced in float64 for every entry)
    {
        const int bs0 = d.bs0, bs1 = d.bs1, nA = cfg.nA;
        if (kBeams) {
            // TX steering of the true panel and gain * RX steering go to scratch tables in the (still unused) W tile area;
            // the codebook is staged one panel row (tb0 entries of every beam) at a time in the (still unused) A strip area,
            // so its loads are coalesced and off the dependent path.  Odd table strides: lanes that differ in p hit different banks.
            const int tb0 = bf.bs0, tb1 = bf.bs1, Mr = d.Mr, B = bf.n_beams;
            const int sy = tb0 | 1, sz = tb1 | 1;
            float2* xY = sW;                              // [np][sy]
            float2* xZ = xY + np * sy;                    // [np][sz]
            float2* xR = xZ + np * sz;                    // [np][Mr]
            float2* fS = reinterpret_cast<float2*>(sA);   // [B][tb0]
            for (int e = tid; e < np * tb0; e += kFdThreads) { const int p = e / tb0, y = e - p * tb0; xY[p * sy + y] = phasor_cycles((double)y * sh.u[0][p]); }
            for (int e = tid; e < np * tb1; e += kFdThreads) { const int p = e / tb1, z = e - p * tb1; xZ[p * sz + z] = phasor_cycles((double)z * sh.v[0][p]); }
            for (int e = tid; e < np * Mr; e += kFdThreads) {
                const int p = e / Mr, r = e - p * Mr;
                const int yr = r % d.ue0, zr = r / d.ue0;
                xR[e] = cmul(sh.c[p], phasor_cycles((double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
            }
            for (int e = tid; e < Mr * B; e += kFdThreads) rows[e] = 0.f;
            for (int e = tid; e < np; e += kFdThreads) tY[e] = make_float2(1.f, 0.f);          // virtual panel: bs0 == 1
            // G[b][p] = sum_z xZ[p][z] sum_y F[b, z*tb0 + y] xY[p][y]; a thread owns up to kBeamAcc (beam, path) pairs
            constexpr int kBeamAcc = 8;
            float2 g[kBeamAcc];
            #pragma unroll
            for (int k = 0; k < kBeamAcc; ++k) g[k] = make_float2(0.f, 0.f);
            const int n_bp = B * np;
            for (int z = 0; z < tb1; ++z) {
                __syncthreads();                          // tables written (z == 0) / the previous row's readers are done with fS
                for (int e = tid; e < B * tb0; e += kFdThreads) {
                    const int b = e / tb0, y = e - b * tb0;
                    fS[e] = __ldg(bf.F + (long long)b * (tb0 * tb1) + z * tb0 + y);
                }
                __syncthreads();
                #pragma unroll
                for (int k = 0; k < kBeamAcc; ++k) {
                    const int e = tid + k * kFdThreads;
                    if (e < n_bp) {
                        const int b = e / np, p = e - b * np;
                        const float2* f = fS + b * tb0;
                        const float2* ty = xY + p * sy;
                        float2 gz = make_float2(0.f, 0.f);
                        for (int y = 0; y < tb0; ++y) {
                            const float2 fv = f[y], t = ty[y];
                            gz.x = fmaf(fv.x, t.x, gz.x); gz.x = fmaf(-fv.y, t.y, gz.x);
                            gz.y = fmaf(fv.x, t.y, gz.y); gz.y = fmaf(fv.y, t.x, gz.y);
                        }
                        const float2 tz = xZ[p * sz + z];
                        g[k].x = fmaf(gz.x, tz.x, g[k].x); g[k].x = fmaf(-gz.y, tz.y, g[k].x);
                        g[k].y = fmaf(gz.x, tz.y, g[k].y); g[k].y = fmaf(gz.y, tz.x, g[k].y);
                    }
                }
            }
            // tQ[p][r * B + b] = c_p a_rx[r,p] G[b,p]
            #pragma unroll
            for (int k = 0; k < kBeamAcc; ++k) {
                const int e = tid + k * kFdThreads;
                if (e < n_bp) {
                    const int b = e / np, p = e - b * np;
                    for (int r = 0; r < Mr; ++r) tQ[p * nq + r * B + b] = cmul(xR[p * Mr + r], g[k]);
                }
            }
        } else {
        for (int e = tid; e < np * bs0; e += kFdThreads) {
            const int p = e / bs0, y = e - p * bs0;
            tY[p * bs0 + y] = phasor_cycles((double)y * sh.u[0][p]);
        }
        for (int e = tid; e < np * nq; e += kFdThreads) {
            const int p = e / nq, q = e - p * nq;
            const int r = q / bs1, z = q - r * bs1;
            const int yr = r % d.ue0, zr = r / d.ue0;
            tQ[p * nq + q] = cmul(sh.c[p], phasor_cycles((double)z * sh.v[0][p] + (double)yr * sh.u[1][p] + (double)zr * sh.v[1][p]));
        }
        }
        for (int e = tid; e < np * nA; e += kFdThreads) {
            const int p = e / nA, a = e - p * nA;
            wA[p * nA + a] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_start + d.subc_step * 16 * a)));
        }
        for (int e = tid; e < np * 16; e += kFdThreads) {
            const int p = e >> 4, b = e & 15;
            wB[e] = phasor_cycles(-(sh.wcyc[p] * (double)(d.subc_step * b)));
        }
    }

    float4* sAw = sA + warp * cfg.pcap * 8;
    const int my_i = lane & 7;                   // row of the warp's 8-row strip this lane builds
    const int p_lane = lane >> 3;                // first path this lane builds (then +4, +8, ...)

    for (int ct = ks; ct < n_ct; ct += ksplit) {
        const int col0 = ct * kTKW;
        __syncthreads();                          // tables ready (first pass) / W tile readers done
        for (int e = tid; e < np * kTKW; e += kFdThreads) {
            const int p = e >> 8, c = e & (kTKW - 1);
            const int col = col0 + c;
            float2 w = make_float2(0.f, 0.f);
            if (col < ncols) w = cmul(wA[p * cfg.nA + (col >> 4)], wB[p * 16 + (col & 15)]);
            sW[e] = w;
        }
        __syncthreads();
        const int n_pass = (ncols - col0 > kTK) ? 2 : 1;

        for (int rt = 0; rt < n_rt; ++rt) {
            const int row0 = rt * kTM + warp * 8;
            if (row0 >= d.M) break;              // warp-uniform
            // ---- this warp's A strip: [np][8 rows], one complex multiply per entry
            __syncwarp();
            {
                const int m = row0 + my_i;
                const bool ok = m < d.M;
                const unsigned mm = ok ? (unsigned)m : 0u;
                const unsigned r = cfg.mul_mt ? __umulhi(mm, cfg.mul_mt) : mm;        // Mt == 1 -> r = m
                const unsigned t = mm - r * (unsigned)d.Mt;
                const unsigned zt = cfg.mul_bs0 ? __umulhi(t, cfg.mul_bs0) : t;       // bs0 == 1 -> z = t
                const unsigned yt = t - zt * (unsigned)d.bs0;
                const float2* q_ptr = tQ + (r * (unsigned)d.bs1 + zt);
                const float2* y_ptr = tY + yt;
                for (int p = p_lane; p < np; p += 4) {
                    float2 a = cmul(q_ptr[p * nq], y_ptr[p * d.bs0]);
                    if (!ok) a = make_float2(0.f, 0.f);
                    sAw[p * 8 + my_i] = make_float4(a.x, a.x, a.y, a.y);
                }
            }
            __syncwarp();

            for (int pass = 0; pass < n_pass; ++pass) {
                float2 acc[8][4];
                #pragma unroll
                for (int i = 0; i < 8; ++i)
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

                const float4* wrow = reinterpret_cast<const float4*>(sW) + pass * (kTK / 2) + lane;
                #pragma unroll 1
                for (int p = 0; p < np; ++p) {
                    const float4 w01 = wrow[p * (kTKW / 2)];
                    const float4 w23 = wrow[p * (kTKW / 2) + 32];
                    const float2 w[4]  = {make_float2(w01.x, w01.y), make_float2(w01.z, w01.w),
                                          make_float2(w23.x, w23.y), make_float2(w23.z, w23.w)};
                    const float2 ws[4] = {make_float2(w01.y, w01.x), make_float2(w01.w, w01.z),
                                          make_float2(w23.y, w23.x), make_float2(w23.w, w23.z)};
                    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a = sAw[p * 8 + i];
                        const float2 a1 = make_float2(a.x, a.y), a2 = make_float2(-a.z, a.w);
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[i][j] = __ffma2_rn(a1, w[j], acc[i][j]);
                            acc[i][j] = __ffma2_rn(a2, ws[j], acc[i][j]);
                        }
                    }
                }

                if (kBeams) {
                    // ---- |Y| summed over the pass's columns (W is zero beyond ncols); the warp owns its 8 rows: no atomics
                    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float sa = 0.f;
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float q = fmaf(acc[i][j].x, acc[i][j].x, acc[i][j].y * acc[i][j].y);
                            sa = fmaf(q, rsqrtf(fmaxf(q, 1e-37f)), sa);          // |y| = q * rsqrt(q) (2 ulp; exact 0 for q == 0)
                        }
                        #pragma unroll
                        for (int o = 16; o > 0; o >>= 1) sa += __shfl_xor_sync(0xffffffffu, sa, o);
                        if (lane == 0 && row0 + i < d.M) rows[row0 + i] += sa;
                    }
                    continue;
                }
                // ---- store: lane owns columns {2l, 2l+1} and {64+2l, 64+2l+1} of this 128-column pass
                const int colp = col0 + pass * kTK;
                float2* obase = out_u + (long long)row0 * ncols + colp + 2 * lane;
                if (vec_ok && row0 + 8 <= d.M && colp + kTK <= ncols) {
                    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4* o = reinterpret_cast<float4*>(obase + (long long)i * ncols);
                        __stcs(o,      make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y));
                        __stcs(o + 32, make_float4(acc[i][2].x, acc[i][2].y, acc[i][3].x, acc[i][3].y));
                    }
                } else {
                    #pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (row0 + i >= d.M) break;
                        float2* orow = obase + (long long)i * ncols;
                        #pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = colp + h * 64 + 2 * lane;
                            if (c < ncols)     __stcs(orow + h * 64, acc[i][2 * h]);
                            if (c + 1 < ncols) __stcs(orow + h * 64 + 1, acc[i][2 * h + 1]);
                        }
                    }
                }
            }
        }
    }
    if (kBeams) {
        __syncthreads();
        const int B = bf.n_beams;
        const float inv = 1.0f / ((float)d.Mr * (float)ncols);
        for (int b = tid; b < B; b += kFdThreads) {
            float sa = 0.f;
            for (int r = 0; r < d.Mr; ++r) sa += rows[r * B + b];
            bf.out[user * (long long)B + b] = sa * inv;
        }
    }
}

__global__ void __launch_bounds__(kFdThreads, 2)
fd_fast_kernel(const __grid_constant__ DevDesc d, const FastCfg cfg, const int ksplit)
{
    fd_fast_body<false>(d, cfg, ksplit, BeamCfg{});
}

__global__ void __launch_bounds__(kFdThreads, 2)
bf_fast_kernel(const __grid_constant__ DevDesc d, const FastCfg cfg, const __grid_constant__ BeamCfg bf)
{
    fd_fast_body<true>(d, cfg, 1, bf);
}

}  // namespace dmk
