// dmk_common.cuh -- shared device-side definitions for libdmk (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dmk.h"

namespace dmk {

constexpr int kMaxPaths = DMK_MAX_PATHS;
constexpr double kPi    = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

// Device-side view of dmk_desc + launch arguments (kernel parameter, by value).
struct DevDesc {
    // geometry
    int bs0, bs1, ue0, ue1;       // elements along y, z of each panel (geometry.py:117-120)
    int Mt, Mr, M;                // M = Mr*Mt rows of the per-user [M, K] output matrix
    int P, P0;                    // P = min(num_paths, n_cols), P0 = n_cols
    int N, K, T;                  // OFDM size, selected subcarriers, time snapshots (T >= 1)
    int has_time_axis;            // 1 when n_times > 0
    int rx_filter;                // 1: receive low-pass filter (channel.py:166-168, :193-194), FD only
    int lpf_batch, lpf_log2n;     // paths per FFT batch; log2(N) when N is a power of two (FFT route), else -1 (direct DFT)
    int lpf_cache;                // complex values of the per-user cache of transformed paths (0 = none)
    int fov_any, fov_side[2];     // [0] = BS (AoD), [1] = UE (AoA)
    int pat[2];
    int subc_start, subc_step;    // affine selection if subc_step != 0 or K == 1
    const int32_t* subc;
    const double*  times;
    double sp[2];                 // spacing in wavelengths, [bs, ue]
    // rotation about x, y: sin/cos computed on the host with libm (== NumPy's float64 sin/cos,
    // SURVEY.md Appendix A R3); about z: radians.  [0] = BS, [1] = UE (uniform case).
    double sx[2], cx[2], sy[2], cy[2], rz[2];
    // FoV thresholds exactly as geometry.py:184-190 computes them in float64
    double h_lo[2], h_hi[2], v_lo[2], v_hi[2];
    int    in_f64;                // 1: the seven path matrices are float64 (DMK_FLAG_F64_INPUTS): all-float64 prologue
    double ts_f64;                // 1/bandwidth            (channel.py:223)
    float  ts_f32;                // float32(1/bandwidth)   (channel.py:183, R11)
    float  n_f32;                 // float32(N)
    double inv_n;                 // 1/N
    int    ld;
    int    n_sms;                 // SMs of the device the launch runs on (prefetch distance of the one-CTA-per-user kernels)
    long long n_users;
    const float *power, *phase, *delay, *az[2], *el[2], *doppler;   // az/el: [0] = AoD, [1] = AoA
    const double* ue_rot;         // per-user [n,3] degrees or nullptr
    uint8_t *fov_mask, *valid_mask, *clip_mask;
    int32_t* path_slot;
    float*   tau_out;             // time domain only: [n, P] delay of the path in each output slot, 0 in empty slots (Sionna layout), or nullptr
    float2*  out;
};

// ---------------------------------------------------------------------------------------------
// NumPy float32 sin/cos, bit-exact (rounding point R2; geometry.py:301-302 via NumPy 2.3.5's SIMD
// kernel).  CPU twin: oracle/np_trig_emul.c, verified there against np.sin/np.cos on every float32
// in [-2pi, 2pi].  Every operation uses an explicit-rounding intrinsic so nvcc cannot re-associate
// or contract differently from the CPU sequence.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void np_sincosf(float x, float& s_out, float& c_out)
{
    const float magic = 0x1.800000p+23f;
    float q = __fmaf_rn(x, 0x1.45f306p-1f, magic);
    q = __fsub_rn(q, magic);
    float r = __fmaf_rn(q, -0x1.921fb0p+00f, x);
    r = __fmaf_rn(q, -0x1.5110b4p-22f, r);
    r = __fmaf_rn(q, -0x1.846988p-48f, r);
    float r2 = __fmul_rn(r, r);
    float c = __fmaf_rn(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f);
    c = __fmaf_rn(c, r2, 0x1.55553cp-05f);
    c = __fmaf_rn(c, r2, -0x1.000000p-01f);
    c = __fmaf_rn(c, r2, 0x1.000000p+00f);
    float s = __fmaf_rn(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f);
    s = __fmaf_rn(s, r2, 0x1.11119ap-07f);
    s = __fmaf_rn(s, r2, -0x1.555556p-03f);
    s = __fmaf_rn(s, r2, 0.0f);
    s = __fmaf_rn(s, r, r);
    int iq = __float2int_rn(q);
    float rs = ((iq & 1) == 0) ? s : c;           // sine: quadrant iq
    if (iq & 2) rs = __fsub_rn(0.0f, rs);
    int ic = iq + 1;                              // cosine: quadrant iq + 1
    float rc = ((ic & 1) == 0) ? s : c;
    if (ic & 2) rc = __fsub_rn(0.0f, rc);
    if (x != x) { rs = __int_as_float(0x7fc00000); rc = rs; }
    s_out = rs; c_out = rc;
}

// Branch-free float64 sin/cos for the rotation chains (geometry.py:294-299: sin/cos of phi - gamma and of the rotation angles).
// The CUDA library's sincos(double) is just as accurate but carries a slow-path branch (|x| > 105615) that splits the
// prologue into basic blocks, so ptxas cannot interleave the independent TX-side / RX-side chains; the prologue is a
// latency-bound float64 dependency chain, and instruction-level parallelism between the two sides is what shortens it.
// Cody-Waite reduction by pi/2 with a 33 + 53-bit split (exact product for |q| < 2^20, i.e. |x| < 1.6e6; larger arguments
// lose accuracy gracefully, NaN/Inf -> NaN), fdlibm kernel polynomials on [-pi/4, pi/4]: <= 1 ulp away from multiples of
// pi/2, absolute error <= 2e-16 everywhere -- the same class as libm / NumPy (f64-ulp noise is
// irrelevant at the 1e-5 bar and flips a FoV compare with probability ~1e-16 per path).
__device__ __forceinline__ void dsincos_bf(double x, double& s_out, double& c_out)
{
    const double q = rint(x * 6.36619772367581382433e-01);                 // x * 2/pi
    double r = fma(-q, 1.57079632673412561417e+00, x);                     // pio2_1 (33 bits)
    r = fma(-q, 6.07710050650619224932e-11, r);                            // pio2_1t
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sn = fma(ps * z, r, r);                                    // r + r^3 * poly
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));                    // 1 - z/2 + z^2 * poly
    const int iq = __double2int_rn(q);
    double rs = (iq & 1) ? cs : sn;
    double rc = (iq & 1) ? sn : cs;
    rs = (iq & 2) ? -rs : rs;
    rc = ((iq + 1) & 2) ? -rc : rc;
    s_out = rs; c_out = rc;
}

// exp(j 2 pi cyc) as float2, cyc in cycles (float64): reduce in double, evaluate in float32
// (tau*f reaches thousands of cycles, steering phases ~100 cycles).
__device__ __forceinline__ float2 phasor_cycles(double cyc)
{
    // Fraction of a cycle in float64, then a short branch-free float32 evaluation: t = 2 fr in [-1, 1] (angle pi t),
    // quadrant q = rint(2 t), r = t - q/2 exact in [-1/4, 1/4], Taylor polynomials of sin/cos(pi r) (|pi r| <= pi/4:
    // truncation < 2e-9).  The float32 rounding of the fraction bounds the error at 1.9e-7 rad, the same as
    // sincospif(2.0f * (float)fr) which this replaces at about half the instructions (verified on 3e7 random arguments).
    const double fr = cyc - rint(cyc);            // [-0.5, 0.5]
    const float t = 2.0f * (float)fr;
    const float q = rintf(2.0f * t);
    const float r = fmaf(q, -0.5f, t);
    const float x = r * 3.14159265358979323846f;
    const float z = x * x;
    float ps = fmaf(z, 2.7557319e-6f, -1.9841270e-4f);
    ps = fmaf(ps, z, 8.3333333e-3f);
    ps = fmaf(ps, z, -1.6666667e-1f);
    const float sn = fmaf(ps * z, x, x);
    float pc = fmaf(z, -2.7557319e-7f, 2.4801587e-5f);
    pc = fmaf(pc, z, -1.3888889e-3f);
    pc = fmaf(pc, z, 4.1666667e-2f);
    pc = fmaf(pc, z, -0.5f);
    const float cs = fmaf(pc, z, 1.0f);
    const int iq = __float2int_rn(q);
    float rs = (iq & 1) ? cs : sn, rc = (iq & 1) ? sn : cs;
    rs = (iq & 2) ? -rs : rs;
    rc = ((iq + 1) & 2) ? -rc : rc;
    return make_float2(rc, rs);
}

// exp(j 2 pi cyc) with the SFU: the fraction of a cycle is still taken in float64 (tau * f reaches thousands of cycles), the
// evaluation is sin.approx / cos.approx of 2 pi fr, |2 pi fr| <= pi.  MUFU.SIN/COS are accurate to 2^-21.4 absolute on [-pi, pi]
// (CUDA programming guide, intrinsic table), i.e. <= 3.6e-7 on a unit phasor -- three times the polynomial version above, still
// 28x inside the 1e-5 parity bar -- at 9 instead of ~35 instructions.  Used where a kernel is bound by the number of phasors it
// has to evaluate per user (small arrays: a few KB of output per user).
__device__ __forceinline__ float2 phasor_cycles_sfu(double cyc)
{
    const double fr = cyc - rint(cyc);            // [-0.5, 0.5]
    const float x = (float)fr * 6.28318530717958647692f;
    return make_float2(__cosf(x), __sinf(x));
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

__device__ __forceinline__ int subcarrier_at(const DevDesc& d, int i)
{
    return d.subc ? d.subc[i] : d.subc_start + d.subc_step * i;
}

}  // namespace dmk
