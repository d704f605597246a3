// dmk_prologue.cuh -- per-(user, path) prologue: rotation, FoV, element pattern, path gain.
//
// One thread evaluates one path column of one user.  dtype flow follows the reference for
// float32 inputs; comments cite /root/reference files.
#pragma once
#include "dmk_common.cuh"

namespace dmk {

struct PathState {
    bool   valid;        // ~isnan(power) and column < num_paths            (channel.py:260, dataset.py:258-261)
    bool   fov;          // Dataset._fov_mask (true when no mask is built)   (dataset.py:493-504)
    bool   over;         // delay_n >= N                                     (channel.py:187)
    bool   contrib;      // path adds a non-zero term to H
    float2 c;            // complex path gain: FD sqrt(p/N) e^{j phase}, TD sqrt(p) e^{j phase}
    double wcyc;         // delay_n / N: cycles per unit subcarrier index
    double u[2], v[2];   // steering cycles per element step along y, z; [0] = TX, [1] = RX
    double fd;           // Doppler shift, Hz
    double th[2], ph[2]; // rotated angles, radians (before FoV NaN-ing)
    double pw;           // power_linear_ant_gain (float64 view)
};

// Global -> local direction of one side.  geometry.py:284-310.
//
// The reference goes through the angles: theta' = arccos(x), phi' = angle(re + j im), and the array response then
// takes sin/cos of them again (geometry.py:99-101).  The steering needs only
//     sin(theta') sin(phi') = sqrt((1-x)(1+x)) * im / hypot(re, im),      cos(theta') = x,
// which are the same functions of (x, re, im) evaluated to a few float64 ulps, so the angles themselves (acos, atan2:
// ~half of the prologue's dependent latency) are computed only when something consumes them -- the FoV compares
// (geometry.py:180-193), the dipole pattern (ant_patterns.py:57-69) or the by-product kernel.
template <bool kNeedAngles>
__device__ __forceinline__ void rotate_side(float el_deg, float az_deg,
                                            double sx, double cx, double sy, double cy, double rz,
                                            double& th, double& ph, double& sin_th_sin_ph, double& cos_th)
{
    const float d2r = 0x1.1df46ap-6f;                       // float32(pi/180): np.deg2rad on float32 (R1)
    float th32 = __fmul_rn(el_deg, d2r);                    // :284
    float ph32 = __fmul_rn(az_deg, d2r);                    // :285
    float st32, ct32;
    np_sincosf(th32, st32, ct32);                           // :301-302 float32 SIMD sin/cos (R2)
    double st = (double)st32, ct = (double)ct32;
    double dphi = __dsub_rn((double)ph32, rz);              // :294 float32 - float64 -> float64 (R3)
    double sd, cd;
    sincos(dphi, &sd, &cd);
    // :305-306  arccos(cy*cx*ct + st*(sy*cx*cd - sx*sd)), same association, no contraction
    double x = __dadd_rn(__dmul_rn(__dmul_rn(cy, cx), ct),
                         __dmul_rn(st, __dsub_rn(__dmul_rn(__dmul_rn(sy, cx), cd), __dmul_rn(sx, sd))));
    // :308-310  angle(cy*st*cd - sy*ct + 1j*(cy*sx*ct + st*(sy*sx*cd + cx*sd)))
    double re = __dsub_rn(__dmul_rn(__dmul_rn(cy, st), cd), __dmul_rn(sy, ct));
    double im = __dadd_rn(__dmul_rn(__dmul_rn(cy, sx), ct),
                          __dmul_rn(st, __dadd_rn(__dmul_rn(__dmul_rn(sy, sx), cd), __dmul_rn(cx, sd))));
    re = __dadd_rn(re, 0.0);                                // real + real(1j*im) = re + 0
    if (kNeedAngles) {
        th = acos(x);
        ph = atan2(im, re);
    } else {
        // NaN exactly where the reference's angles are NaN (|x| > 1, NaN inputs); finite otherwise
        th = (fabs(x) <= 1.0) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
        ph = (re == re && im == im) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
    }
    const double sin_th = sqrt(__dmul_rn(__dsub_rn(1.0, x), __dadd_rn(1.0, x)));     // sin(arccos(x)); NaN for |x| > 1
    const double h = sqrt(__dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im)));
    const double sin_ph = (h > 0.0) ? __ddiv_rn(im, h) : 0.0;                        // sin(atan2(im, re)); atan2(0, +0) = 0
    sin_th_sin_ph = sin_th * sin_ph;
    cos_th = x;
}

// geometry.py:180-193 on one side.  theta in [0, pi] so mod(theta, 2pi) == theta; phi in (-pi, pi].
__device__ __forceinline__ bool in_fov(const DevDesc& d, int side, double th, double ph)
{
    double phm = (ph < 0.0) ? __dadd_rn(ph, kTwoPi) : ph;  // np.mod(phi, 2*pi)
    bool in_h = (phm <= d.h_lo[side]) || (phm >= d.h_hi[side]);
    bool in_v = (th <= d.v_hi[side]) && (th >= d.v_lo[side]);
    return in_h && in_v;
}

// ant_patterns.py:51-69: 1.643 * cos(pi/2 cos th)^2 / sin th where |sin th| > 1e-10, else 0.
__device__ __forceinline__ double dipole_gain(double th)
{
    double s = sin(th);
    if (!(fabs(s) > 1e-10)) return 0.0;                     // NaN -> 0
    double ct = cos(__dmul_rn(kPi / 2, cos(th)));
    return __dmul_rn(1.643, __ddiv_rn(__dmul_rn(ct, ct), s));
}

template <bool kFreqDomain, bool kNeedAngles>
__device__ __forceinline__ void path_prologue_impl(const DevDesc& d, long long user, int p, PathState& o)
{
    const long long off = user * (long long)d.ld + p;
    const float pw_db = d.power[off];
    o.valid = (p < d.P) && !(pw_db != pw_db);

    // ---- rotation (dataset.py:341-349)
    double sxu = d.sx[1], cxu = d.cx[1], syu = d.sy[1], cyu = d.cy[1], rzu = d.rz[1];
    if (d.ue_rot) {                                         // per-user UE rotation (dataset.py:328-338)
        const double k = kPi / 180.0;                       // np.deg2rad float64: x * (pi/180)
        const double* r = d.ue_rot + user * 3;
        sincos(__dmul_rn(r[0], k), &sxu, &cxu);
        sincos(__dmul_rn(r[1], k), &syu, &cyu);
        rzu = __dmul_rn(r[2], k);
    }
    double ss[2], cc[2];
    rotate_side<kNeedAngles>(d.el[0][off], d.az[0][off], d.sx[0], d.cx[0], d.sy[0], d.cy[0], d.rz[0], o.th[0], o.ph[0], ss[0], cc[0]);
    rotate_side<kNeedAngles>(d.el[1][off], d.az[1][off], sxu, cxu, syu, cyu, rzu, o.th[1], o.ph[1], ss[1], cc[1]);

    // ---- FoV (dataset.py:493-512)
    bool fov = true;
    if (d.fov_any) {
        if (d.fov_side[0]) fov = fov && in_fov(d, 0, o.th[0], o.ph[0]);
        if (d.fov_side[1]) fov = fov && in_fov(d, 1, o.th[1], o.ph[1]);
    }
    o.fov = fov;
    const bool ang_ok = !(o.th[0] != o.th[0]) && !(o.th[1] != o.th[1]) &&
                        !(o.ph[0] != o.ph[0]) && !(o.ph[1] != o.ph[1]);

    // ---- steering cycles per element step (geometry.py:99-101 with x == 0; dataset.py:393)
    #pragma unroll
    for (int s = 0; s < 2; ++s) {
        o.u[s] = d.sp[s] * ss[s];
        o.v[s] = d.sp[s] * cc[s];
    }

    // ---- power: generator_utils.py:35 (float32 divide, float32 pow; R7), ant_patterns.py:167-168
    float  p_lin = (float)exp10((double)__fdiv_rn(pw_db, 10.0f));
    const bool f64_power = (d.pat[0] != DMK_PATTERN_ISOTROPIC) || (d.pat[1] != DMK_PATTERN_ISOTROPIC);
    double pw64 = (double)p_lin;
    if (f64_power) {
        const double nan64 = __longlong_as_double(0x7ff8000000000000LL);
        double gt = (d.pat[0] == DMK_PATTERN_HALFWAVE_DIPOLE) ? dipole_gain(fov ? o.th[0] : nan64) : 1.0;
        double gr = (d.pat[1] == DMK_PATTERN_HALFWAVE_DIPOLE) ? dipole_gain(fov ? o.th[1] : nan64) : 1.0;
        pw64 = __dmul_rn(pw64, __dmul_rn(gt, gr));
    }
    o.pw = pw64;

    // ---- path gain
    const float d2r = 0x1.1df46ap-6f;
    const float ph32 = __fmul_rn(d.phase[off], d2r);        // np.deg2rad(phase) float32
    const float ec = (float)cos((double)ph32);              // complex64 exp(1j*x): cosf/sinf (R10, <= 1 ulp)
    const float es = (float)sin((double)ph32);
    o.over = false;
    if (kFreqDomain) {
        float dn = __fdiv_rn(d.delay[off], d.ts_f32);       // channel.py:183 (R11)
        o.over = dn >= d.n_f32;                             // :187 (R12)
        if (o.over) dn = d.n_f32;                           // :189
        o.wcyc = (double)dn * d.inv_n;
        if (f64_power) {
            double amp = o.over ? 0.0 : sqrt(__ddiv_rn(pw64, (double)d.N));   // :188, :192 (float64 branch)
            o.c = make_float2((float)(amp * (double)ec), (float)(amp * (double)es));
        } else {
            float amp = o.over ? 0.0f : __fsqrt_rn(__fdiv_rn(p_lin, d.n_f32));  // :188, :192 (R9)
            o.c = make_float2(__fmul_rn(amp, ec), __fmul_rn(amp, es));
        }
    } else {
        o.wcyc = 0.0;
        if (f64_power) {
            double amp = sqrt(pw64);                        // channel.py:286
            o.c = make_float2((float)(amp * (double)ec), (float)(amp * (double)es));
        } else {
            float amp = __fsqrt_rn(p_lin);
            o.c = make_float2(__fmul_rn(amp, ec), __fmul_rn(amp, es));
        }
    }
    o.fd = d.doppler ? (double)d.doppler[off] : 0.0;
    const bool c_ok = (o.c.x == o.c.x) && (o.c.y == o.c.y) && (o.wcyc == o.wcyc) && (o.fd == o.fd);
    // exact zeros where theta is NaN (geometry.py:65-80) or outside the FoV (dataset.py:508-511);
    // NaN gains are dropped by nansum (channel.py:283).
    o.contrib = o.valid && fov && ang_ok && c_ok && !(o.c.x == 0.0f && o.c.y == 0.0f);
}

// The angles are consumed by the FoV compares and the dipole pattern only.
template <bool kFreqDomain>
__device__ __forceinline__ void path_prologue(const DevDesc& d, long long user, int p, PathState& o)
{
    if (d.fov_any || d.pat[0] != DMK_PATTERN_ISOTROPIC || d.pat[1] != DMK_PATTERN_ISOTROPIC)
        path_prologue_impl<kFreqDomain, true>(d, user, p, o);
    else
        path_prologue_impl<kFreqDomain, false>(d, user, p, o);
}

// By-product kernel: always materialise the angles.
__device__ __forceinline__ void path_prologue_angles(const DevDesc& d, long long user, int p, PathState& o)
{
    path_prologue_impl<false, true>(d, user, p, o);
}

}  // namespace dmk
