// dmk_td.cuh -- fused time-domain channel kernel (freq_domain = 0) and per-path by-product kernels.
//
//   H[u, m, j, it] = c_{p_j} * exp(j 2 pi (steer(m, p_j) + f_D[p_j] t_it))     j < n_valid(u)
//                  = 0                                                        otherwise
// p_j is the j-th column of user u whose power is not NaN (channel.py:260,:274-287); a path outside
// the FoV keeps its slot and is zero (dataset.py:508-511 + geometry.py:65-80).  The kernel is bound
// by the output write: every element is one complex multiply of two shared-memory table entries.
#pragma once
#include "dmk_prologue.cuh"

namespace dmk {

constexpr int kTdThreads = 256;
constexpr int kTdRows = 64;               // rows of the steering table built per pass

struct TdShared {
    float2 c[kMaxPaths];
    double fd[kMaxPaths];
    double u[2][kMaxPaths];
    double v[2][kMaxPaths];
    unsigned char contrib[kMaxPaths];
    int nv;
};

__device__ __forceinline__ void td_cta_prologue(const DevDesc& d, long long user, TdShared& sh, PrologueScratch& sc)
{
    cta_prologue_chains<false>(d, user, sc);
    if (threadIdx.x >= 224) prefetch_user_rows_shifted(d, user);
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        PathState st;
        const bool active = lane < d.P0;
        st.contrib = false; st.valid = false; st.fov = true; st.over = false;
        if (active) prologue_combine<false>(d, sc.side[0][lane], sc.side[1][lane], sc.gain[lane], st);
        const unsigned ballot = __ballot_sync(0xffffffffu, active && st.valid);
        const int j = __popc(ballot & ((1u << lane) - 1u));
        if (active && st.valid) {
            sh.c[j] = st.c; sh.fd[j] = st.fd; sh.contrib[j] = st.contrib ? 1 : 0;
            sh.u[0][j] = st.u[0]; sh.v[0][j] = st.v[0];
            sh.u[1][j] = st.u[1]; sh.v[1][j] = st.v[1];
        }
        if (lane == 0) sh.nv = __popc(ballot);
        if (active) {
            const long long o = user * (long long)d.P0 + lane;
            if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
            if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
            if (d.path_slot)  d.path_slot[o]  = st.valid ? j : -1;
        }
        if (d.tau_out) {
            // slot j <- ToA of the j-th valid column (sionna_adapter.py:196-198: tau[..., :num_paths] = ToA), zeros behind
            float* tau_u = d.tau_out + user * (long long)d.P;
            if (active && st.valid) {
                const long long o = user * (long long)d.ld + lane;
                tau_u[j] = d.in_f64 ? (float)reinterpret_cast<const double*>(d.delay)[o] : d.delay[o];      // tau is np.single (:178)
            }
            const int nv = __popc(ballot);
            if (lane >= nv && lane < d.P) tau_u[lane] = 0.f;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kTdThreads, 4)
td_kernel(const __grid_constant__ DevDesc d)
{
    __shared__ TdShared sh;
    __shared__ PrologueScratch psc;
    __shared__ __align__(16) float2 sA[kTdRows * kMaxPaths]; // [row][slot] gain * steering
    __shared__ __align__(16) float2 sD[kMaxPaths * 64]; // [slot][it] Doppler phasors, 64 snapshots per pass

    const int tid = threadIdx.x;
    const long long user = blockIdx.x;
    td_cta_prologue(d, user, sh, psc);
    const int nv = sh.nv;
    const int P = d.P, T = d.T;
    float2* out_u = d.out + user * (long long)d.M * P * T;

    for (int it0 = 0; it0 < T; it0 += 64) {
        const int tn = min(64, T - it0);
        __syncthreads();
        if (d.has_time_axis) {
            for (int e = tid; e < nv * tn; e += kTdThreads) {
                const int j = e / tn, it = e - j * tn;
                sD[j * 64 + it] = phasor_cycles(sh.fd[j] * d.times[it0 + it]);
            }
        }
        for (int row0 = 0; row0 < d.M; row0 += kTdRows) {
            const int rn = min(kTdRows, d.M - row0);
            __syncthreads();
            for (int e = tid; e < rn * P; e += kTdThreads) {
                const int r = e / P, j = e - r * P;
                float2 a = make_float2(0.f, 0.f);
                if (j < nv && sh.contrib[j]) {
                    const int m = row0 + r;
                    const int rr = m / d.Mt, t = m - rr * d.Mt;
                    const int yt = t % d.bs0, zt = t / d.bs0;
                    const int yr = rr % d.ue0, zr = rr / d.ue0;
                    const double cyc = (double)yt * sh.u[0][j] + (double)zt * sh.v[0][j]
                                     + (double)yr * sh.u[1][j] + (double)zr * sh.v[1][j];
                    a = cmul(sh.c[j], phasor_cycles(cyc));
                }
                sA[r * P + j] = a;
            }
            __syncthreads();
            if (!d.has_time_axis) {
                // [rn, P] block is contiguous in the output
                float2* o = out_u + (long long)row0 * P;
                for (int e = tid; e < rn * P; e += kTdThreads) __stcs(o + e, sA[e]);
            } else if ((tn & 1) == 0 && (T & 1) == 0 && (kTdThreads % (tn >> 1)) == 0) {
                // fast path: a thread owns two consecutive snapshots (one 16-byte store) and walks the (row, slot) pairs with
                // incremental indices -- no integer division per element
                const int tpr = tn >> 1;                         // threads per (row, slot) pair
                const int step = kTdThreads / tpr;               // (row, slot) pairs per pass
                const int it = (tid % tpr) * 2;
                int rj = tid / tpr;
                int j = rj % P;
                const int jstep = step % P;
                float4* o = reinterpret_cast<float4*>(out_u + ((long long)row0 * P + rj) * T + it0 + it);
                const long long ostep = (long long)step * T / 2;  // float4 units
                for (; rj < rn * P; rj += step) {
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < nv) {
                        const float2 a = sA[rj];
                        const float4 dd = *reinterpret_cast<const float4*>(&sD[j * 64 + it]);
                        val = make_float4(fmaf(a.x, dd.x, -a.y * dd.y), fmaf(a.x, dd.y, a.y * dd.x),
                                          fmaf(a.x, dd.z, -a.y * dd.w), fmaf(a.x, dd.w, a.y * dd.z));
                    }
                    __stcs(o, val);
                    o += ostep;
                    j += jstep;
                    if (j >= P) j -= P;
                }
            } else {
                for (int e = tid; e < rn * P * tn; e += kTdThreads) {
                    const int rj = e / tn, it = e - rj * tn;         // rj = r*P + j
                    const int j = rj % P;
                    float2 val = make_float2(0.f, 0.f);
                    if (j < nv) val = cmul(sA[rj], sD[j * 64 + it]);
                    __stcs(out_u + ((long long)row0 * P + rj) * T + it0 + it, val);
                }
            }
        }
    }
}

// =================================================================================================
// td_warp_kernel -- the reference's own time-domain mode (freq_domain = 0, no time axis): one WARP per user.
//
// Without a time axis a table entry of td_kernel is used once, so that kernel evaluates a float64-reduced polynomial phasor
// (plus the integer divisions of its index arithmetic) for every 8 bytes it writes and gives a whole CTA to every user:
// 0.3 TB/s on 8x1 panels (1.6 KB per user), 1.0-1.3 TB/s on larger ones.  Here
//   lanes = path columns: the three float64 chains + combine (as everywhere), masks, path slots, tau;
//   compaction: lane j < P becomes output slot j and fetches the state of the j-th valid column by shuffles
//     (channel.py:274-287: the valid paths lead, FoV-masked ones keep their slot with a zero, the rest is zero);
//   the warp walks the antenna rows (y fastest, then z, then the RX element): a float64 multiply-add on the running phase, an SFU
//     phasor (float64-reduced argument, <= 3.6e-7), a complex multiply and ONE store instruction per row -- 8 P contiguous bytes;
//     rows are contiguous, so a user is one linear stream.
// No shared memory, no barriers; warps of a CTA only share the launch.
// =================================================================================================
constexpr int kTdwWarps = 4;

__global__ void __launch_bounds__(kTdwWarps * 32)
td_warp_kernel(const __grid_constant__ DevDesc d)
{
    const int lane = threadIdx.x & 31;
    const long long user = (long long)blockIdx.x * kTdwWarps + (threadIdx.x >> 5);
    if (user >= d.n_users) return;                          // warp-uniform
    const bool triv0 = side_angles_trivial(d, 0), triv1 = side_angles_trivial(d, 1);
    PathState st;
    const bool active = lane < d.P0;
    st.contrib = false; st.valid = false; st.fov = true; st.over = false;
    st.c = make_float2(0.f, 0.f); st.u[0] = st.u[1] = st.v[0] = st.v[1] = 0.0;
    if (active) {
        SideOut s0, s1; GainOut g;
        if (prologue_needs_angles(d)) { prologue_side<true>(d, user, lane, 0, s0, d.Mt > 1);  prologue_side<true>(d, user, lane, 1, s1, d.Mr > 1); }
        else                          { prologue_side_auto<false>(d, user, lane, 0, s0, d.Mt > 1, triv0); prologue_side_auto<false>(d, user, lane, 1, s1, d.Mr > 1, triv1); }
        prologue_gain<false>(d, user, lane, g);
        prologue_combine<false>(d, s0, s1, g, st);
    }
    const unsigned vb = __ballot_sync(0xffffffffu, active && st.valid);
    const int nv = __popc(vb);
    const int rank = __popc(vb & ((1u << lane) - 1u));
    if (active) {
        const long long o = user * (long long)d.P0 + lane;
        if (d.fov_mask)   d.fov_mask[o]   = st.fov ? 1 : 0;
        if (d.valid_mask) d.valid_mask[o] = st.valid ? 1 : 0;
        if (d.path_slot)  d.path_slot[o]  = st.valid ? rank : -1;
    }
    if (d.tau_out) {                                        // slot j <- ToA of the j-th valid column (sionna_adapter.py:196-198), zeros behind
        float* tau_u = d.tau_out + user * (long long)d.P;
        if (active && st.valid) {
            const long long o = user * (long long)d.ld + lane;
            tau_u[rank] = d.in_f64 ? (float)reinterpret_cast<const double*>(d.delay)[o] : d.delay[o];
        }
        if (lane >= nv && lane < d.P) tau_u[lane] = 0.f;
    }
    // slot `lane` takes the state of the lane-th valid column (a non-contributing path -- outside the FoV, NaN angle -- is a zero)
    const int src = (lane < nv) ? (int)__fns(vb, 0, lane + 1) : 0;
    const bool on = __shfl_sync(0xffffffffu, (int)st.contrib, src) != 0 && lane < nv;
    float2 c;
    c.x = __shfl_sync(0xffffffffu, st.c.x, src); c.y = __shfl_sync(0xffffffffu, st.c.y, src);
    double u0 = __shfl_sync(0xffffffffu, st.u[0], src), v0 = __shfl_sync(0xffffffffu, st.v[0], src);
    double u1 = __shfl_sync(0xffffffffu, st.u[1], src), v1 = __shfl_sync(0xffffffffu, st.v[1], src);
    if (!on) { c = make_float2(0.f, 0.f); u0 = v0 = u1 = v1 = 0.0; }      // empty slots and zeroed paths: (0, 0) x a finite phasor

    const int P = d.P;
    float2* o = d.out + user * (long long)d.M * P + lane;
    const bool wr = lane < P;
    double cyc_r = 0.0;                                      // RX element (y_r, z_r): y_r u_rx + z_r v_rx
    for (int zr = 0; zr < d.ue1; ++zr) {
        double cyc_ry = cyc_r;
        for (int yr = 0; yr < d.ue0; ++yr) {
            double cyc_z = cyc_ry;                           // + z_t v_tx
            for (int zt = 0; zt < d.bs1; ++zt) {
                double cyc = cyc_z;                          // + y_t u_tx
                #pragma unroll 4
                for (int yt = 0; yt < d.bs0; ++yt) {
                    const float2 val = cmul(c, phasor_cycles_sfu(cyc));
                    if (wr) __stcs(o, val);
                    o += P;
                    cyc += u0;
                }
                cyc_z += v0;
            }
            cyc_ry += u1;
        }
        cyc_r += v1;
    }
}

// Per-path by-products (Dataset caches): rotated angles, power with antenna gain, FoV mask.
__global__ void __launch_bounds__(256)
prologue_kernel(const __grid_constant__ DevDesc d, double* __restrict__ angles_rot, double* __restrict__ power_gain)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = d.n_users * d.P0;
    if (idx >= total) return;
    const long long user = idx / d.P0;
    const int p = (int)(idx - user * d.P0);
    PathState st;
    path_prologue_angles(d, user, p, st);
    if (angles_rot) {
        angles_rot[0 * total + idx] = st.th[0];
        angles_rot[1 * total + idx] = st.ph[0];
        angles_rot[2 * total + idx] = st.th[1];
        angles_rot[3 * total + idx] = st.ph[1];
    }
    if (power_gain) power_gain[idx] = st.pw;
    if (d.fov_mask) d.fov_mask[idx] = st.fov ? 1 : 0;
}

// Per-user by-products, one warp per user, lanes = path columns:
//   num_paths  = count of non-NaN FoV-filtered AoA azimuths over ALL columns            (dataset.py:613-619)
//   los        = 1 / 0 / -1 from the interaction code of the first in-FoV path           (dataset.py:569-611)
//   pathloss   = -10 log10 |sum_p sqrt(p_lin) e^{j phase}|^2 (coherent) / (sum_p sqrt(p_lin))^2 (non-coherent), NaN where 0
//                over the raw power / phase matrices, no FoV, no element pattern         (dataset.py:541-566)
__global__ void __launch_bounds__(256)
user_byproducts_kernel(const __grid_constant__ DevDesc d, const float* __restrict__ inter, int* __restrict__ num_paths,
                       int* __restrict__ los, float* __restrict__ pl_coh, float* __restrict__ pl_noncoh)
{
    const long long user = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (user >= d.n_users) return;
    const bool active = lane < d.P0;
    PathState st;
    st.fov = false; st.ph[1] = __longlong_as_double(0x7ff8000000000000LL);
    if (active) path_prologue_angles(d, user, lane, st);
    const bool counted = active && !(st.ph[1] != st.ph[1]) && (!d.fov_any || st.fov);     // where(mask, aoa_az_rot, NaN) is not NaN
    const int np = __popc(__ballot_sync(0xffffffffu, counted));
    const unsigned fovb = __ballot_sync(0xffffffffu, active && st.fov);
    const long long row = user * (long long)d.ld;
    if (lane == 0) {
        if (num_paths) num_paths[user] = np;
        if (los) {
            const bool has = d.fov_any ? (fovb != 0u) : (np > 0);
            const int first = d.fov_any ? (fovb ? __ffs(fovb) - 1 : 0) : 0;
            const float code = inter ? inter[row + first] : __int_as_float(0x7fc00000);
            los[user] = has ? ((d.fov_any && !fovb) ? 0 : (code == 0.0f ? 1 : 0)) : -1;
        }
    }
    if (pl_coh || pl_noncoh) {
        double re = 0.0, im = 0.0, amp_sum = 0.0;
        if (active && d.in_f64) {
            // float64 matrices: sqrt(p_lin).astype(complex64) * exp(1j * deg2rad(phase)) with the phasor in complex128
            const float amp = (float)sqrt(exp10(reinterpret_cast<const double*>(d.power)[row + lane] / 10.0));
            double sn, cs;
            sincos(reinterpret_cast<const double*>(d.phase)[row + lane] * (kPi / 180.0), &sn, &cs);
            const double gr = (double)amp * cs, gi = (double)amp * sn;
            if (!(amp != amp)) amp_sum = (double)amp;
            if (!(gr != gr) && !(gi != gi)) { re = gr; im = gi; }
        } else if (active) {
            const float pw_db = d.power[row + lane];
            const float amp = __fsqrt_rn(exp10f(__fdiv_rn(pw_db, 10.0f)));                 // generator_utils.py:35, sqrt in float32
            float sn, cs;
            sincosf(__fmul_rn(d.phase[row + lane], 0x1.1df46ap-6f), &sn, &cs);             // np.deg2rad on float32, exp(1j x)
            const float gr = __fmul_rn(amp, cs), gi = __fmul_rn(amp, sn);
            if (!(amp != amp)) amp_sum = (double)amp;                                       // nansum drops NaN entries
            if (!(gr != gr) && !(gi != gi)) { re = (double)gr; im = (double)gi; }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, o);
            im += __shfl_xor_sync(0xffffffffu, im, o);
            amp_sum += __shfl_xor_sync(0xffffffffu, amp_sum, o);
        }
        if (lane == 0) {
            const float nanf_ = __int_as_float(0x7fc00000);
            const double tc = re * re + im * im, tn = amp_sum * amp_sum;
            if (pl_coh)    pl_coh[user]    = tc > 0.0 ? (float)(-10.0 * log10(tc)) : nanf_;
            if (pl_noncoh) pl_noncoh[user] = tn > 0.0 ? (float)(-10.0 * log10(tn)) : nanf_;
        }
    }
}

__global__ void np_sincosf_kernel(const float* __restrict__ x, float* __restrict__ s, float* __restrict__ c, long long n)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        float ss, cc;
        np_sincosf(x[i], ss, cc);
        s[i] = ss; c[i] = cc;
    }
}

}  // namespace dmk
