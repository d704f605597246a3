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
    float  ts_f32;                // float32(1/bandwidth)   (channel.py:183, R11)
    float  n_f32;                 // float32(N)
    double inv_n;                 // 1/N
    int    ld;
    long long n_users;
    const float *power, *phase, *delay, *az[2], *el[2], *doppler;   // az/el: [0] = AoD, [1] = AoA
    const double* ue_rot;         // per-user [n,3] degrees or nullptr
    uint8_t *fov_mask, *valid_mask, *clip_mask;
    int32_t* path_slot;
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

// exp(j 2 pi cyc) as float2, cyc in cycles (float64): reduce in double, evaluate in float32
// (tau*f reaches thousands of cycles, steering phases ~100 cycles).
__device__ __forceinline__ float2 phasor_cycles(double cyc)
{
    double fr = cyc - rint(cyc);                  // [-0.5, 0.5]
    float s, c;
    sincospif(2.0f * (float)fr, &s, &c);
    return make_float2(c, s);
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
