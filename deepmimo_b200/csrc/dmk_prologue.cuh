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
// 1 / sqrt(v) for v in (0, 4]: MUFU.RSQ on the float32 value (22 good bits), two Newton steps in float64 (relative error
// ~1e-16 after the second; the first alone would leave 1e-13).  v < 0 -> NaN, NaN -> NaN.  Callers handle v == 0.
__device__ __forceinline__ double drsqrt_nr(double v)
{
    if (v < 1e-30) return __ddiv_rn(1.0, sqrt(v));          // below float32's normal range (a direction within 1e-15 rad of the pole)
    double y = (double)rsqrtf((float)v);
    const double hv = 0.5 * v;
    y = y * fma(-hv * y, y, 1.5);
    y = y * fma(-hv * y, y, 1.5);
    return y;
}

template <bool kNeedAngles>
__device__ __forceinline__ void rotate_core(double st, double ct, double dphi,
                                            double sx, double cx, double sy, double cy,
                                            double& th, double& ph, double& sin_th_sin_ph, double& cos_th, bool steer)
{
    double sd, cd;
    dsincos_bf(dphi, sd, cd);
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
    cos_th = x;
    sin_th_sin_ph = 0.0;
    if (!steer) return;        // single-element panel (e.g. one UE antenna): only the NaN-ness of the angles is consumed
    // sin(arccos(x)) sin(atan2(im, re)) = sqrt(a) * im / sqrt(b),  a = (1 - x)(1 + x) (negative, hence NaN, for |x| > 1),
    // b = re^2 + im^2 (atan2(0, +0) = 0: the product is 0 when b == 0).  Both roots come from drsqrt_nr: two library square roots
    // and a division were a third of this chain's instructions and the longest part of its dependent latency.
    const double a = __dmul_rn(__dsub_rn(1.0, x), __dadd_rn(1.0, x));
    const double b = __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
    const double sin_th = (a == 0.0) ? 0.0 : a * drsqrt_nr(a);
    sin_th_sin_ph = (b > 0.0) ? sin_th * (im * drsqrt_nr(b)) : ((b == b) ? 0.0 : b);
}

// float32 angles (the reference's storage type): deg2rad and sin/cos of theta in float32 exactly as NumPy does them (R1, R2).
template <bool kNeedAngles>
__device__ __forceinline__ void rotate_side(float el_deg, float az_deg,
                                            double sx, double cx, double sy, double cy, double rz,
                                            double& th, double& ph, double& sin_th_sin_ph, double& cos_th, bool steer = true)
{
    const float d2r = 0x1.1df46ap-6f;                       // float32(pi/180): np.deg2rad on float32 (R1)
    float th32 = __fmul_rn(el_deg, d2r);                    // :284
    float ph32 = __fmul_rn(az_deg, d2r);                    // :285
    float st32, ct32;
    np_sincosf(th32, st32, ct32);                           // :301-302 float32 SIMD sin/cos (R2)
    const double dphi = __dsub_rn((double)ph32, rz);        // :294 float32 - float64 -> float64 (R3)
    rotate_core<kNeedAngles>((double)st32, (double)ct32, dphi, sx, cx, sy, cy, th, ph, sin_th_sin_ph, cos_th, steer);
}

// float64 angles (last paragraph): every step of geometry.py:284-310 is float64 in NumPy.
template <bool kNeedAngles>
__device__ __forceinline__ void rotate_side_f64(double el_deg, double az_deg,
                                                double sx, double cx, double sy, double cy, double rz,
                                                double& th, double& ph, double& sin_th_sin_ph, double& cos_th, bool steer = true)
{
    const double k = kPi / 180.0;                           // np.deg2rad on float64: x * (pi/180)
    double st, ct;
    dsincos_bf(__dmul_rn(el_deg, k), st, ct);               // :301-302
    const double dphi = __dsub_rn(__dmul_rn(az_deg, k), rz);
    rotate_core<kNeedAngles>(st, ct, dphi, sx, cx, sy, cy, th, ph, sin_th_sin_ph, cos_th, steer);
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

// ---------------------------------------------------------------------------------------------------------
// The prologue of one path is three independent float64 dependency chains (TX-side rotation, RX-side rotation,
// gain/delay/phase) followed by a short combine.  One thread can run them back to back (by-product kernel); the
// channel kernels give each chain its own warp (lanes = path columns) so the CTA's critical path is one chain.
// ---------------------------------------------------------------------------------------------------------
struct SideOut { double th, ph, ss, cc, gain; };           // rotated angles, sin(th)sin(ph), cos(th), element power gain
struct GainOut { float p_lin, ec, es; double p_lin64, ec64, es64; double wcyc, fd; unsigned char valid, over; };   // *64: float64 inputs

// The float32 entries of one (user, path column), loaded in one batch so that a kernel exposes a single memory latency per path
// instead of one per chain (the compiler does not hoist the loads out of the chains' branches).
struct PathIn { float power, phase, delay, az[2], el[2], doppler; };

__device__ __forceinline__ void load_path_in(const DevDesc& d, long long user, int p, PathIn& in)
{
    const long long off = user * (long long)d.ld + p;
    in.power = __ldg(d.power + off); in.phase = __ldg(d.phase + off); in.delay = __ldg(d.delay + off);
    in.az[0] = __ldg(d.az[0] + off); in.el[0] = __ldg(d.el[0] + off);
    in.az[1] = __ldg(d.az[1] + off); in.el[1] = __ldg(d.el[1] + off);
    in.doppler = d.doppler ? __ldg(d.doppler + off) : 0.0f;
}

__device__ __forceinline__ void side_rotation(const DevDesc& d, long long user, int side, double& sx, double& cx, double& sy, double& cy, double& rz)
{
    sx = d.sx[side]; cx = d.cx[side]; sy = d.sy[side]; cy = d.cy[side]; rz = d.rz[side];
    if (side == 1 && d.ue_rot) {                            // per-user UE rotation (dataset.py:328-338)
        const double k = kPi / 180.0;                       // np.deg2rad float64: x * (pi/180)
        const double* r = d.ue_rot + user * 3;
        dsincos_bf(__dmul_rn(r[0], k), sx, cx);
        dsincos_bf(__dmul_rn(r[1], k), sy, cy);
        rz = __dmul_rn(r[2], k);
    }
}

// A side whose rotated angles enter the result only through their NaN-ness: a single element (no steering), no FoV filter, no
// element pattern, and the identity rotation about x and y (the reference's default UE: shape [1, 1], rotation [0, 0, 0]).  Then
// geometry.py:305-310 reduce to x = cos(theta) EXACTLY (cy cx = 1, the other products are exact zeros) and re, im = st cd, st sd,
// so theta', phi' are NaN exactly when an input angle is not finite: NumPy's float32 cos stays inside [-1, 1] for every finite
// float32 (checked exhaustively, tests/test_np_trig_emul.py::test_float32_sin_cos_stay_inside_unit_interval).  The rotation about
// z only shifts phi.  Saves the whole float64 chain of that side.
__device__ __forceinline__ bool side_angles_trivial(const DevDesc& d, int side)
{
    const int m = side == 0 ? d.Mt : d.Mr;
    return m == 1 && !d.fov_any && d.pat[0] == DMK_PATTERN_ISOTROPIC && d.pat[1] == DMK_PATTERN_ISOTROPIC && !(side == 1 && d.ue_rot) &&
           d.sx[side] == 0.0 && d.cx[side] == 1.0 && d.sy[side] == 0.0 && d.cy[side] == 1.0 && !d.in_f64;
}

__device__ __forceinline__ void prologue_side_trivial(float el_deg, float az_deg, SideOut& o)
{
    const float z = (el_deg - el_deg) + (az_deg - az_deg);           // 0 for finite angles, NaN otherwise
    o.th = (double)z; o.ph = (double)z; o.ss = 0.0; o.cc = 0.0; o.gain = 1.0;
}

// float32 inputs, values already loaded
template <bool kNeedAngles>
__device__ __forceinline__ void prologue_side_in(const DevDesc& d, long long user, int side, const PathIn& in, SideOut& o, bool steer = true)
{
    double sx, cx, sy, cy, rz;
    side_rotation(d, user, side, sx, cx, sy, cy, rz);
    rotate_side<kNeedAngles>(in.el[side], in.az[side], sx, cx, sy, cy, rz, o.th, o.ph, o.ss, o.cc, steer);
    o.gain = 1.0;
}

template <bool kNeedAngles>
__device__ __forceinline__ void prologue_side(const DevDesc& d, long long user, int p, int side, SideOut& o, bool steer = true)
{
    const long long off = user * (long long)d.ld + p;
    double sx, cx, sy, cy, rz;
    side_rotation(d, user, side, sx, cx, sy, cy, rz);
    if (d.in_f64)
        rotate_side_f64<kNeedAngles>(reinterpret_cast<const double*>(d.el[side])[off], reinterpret_cast<const double*>(d.az[side])[off],
                                     sx, cx, sy, cy, rz, o.th, o.ph, o.ss, o.cc, steer);
    else
        rotate_side<kNeedAngles>(d.el[side][off], d.az[side][off], sx, cx, sy, cy, rz, o.th, o.ph, o.ss, o.cc, steer);
    o.gain = 1.0;
}

// prologue_side with the short chain where the side's angles are trivial (side_angles_trivial, evaluated once per kernel)
template <bool kNeedAngles>
__device__ __forceinline__ void prologue_side_auto(const DevDesc& d, long long user, int p, int side, SideOut& o, bool steer, bool triv)
{
    if (triv) {
        const long long off = user * (long long)d.ld + p;
        prologue_side_trivial(d.el[side][off], d.az[side][off], o);
    } else {
        prologue_side<kNeedAngles>(d, user, p, side, o, steer);
    }
}

// float32 inputs: the gain chain from loaded values
template <bool kFreqDomain>
__device__ __forceinline__ void prologue_gain_f32(const DevDesc& d, int p, float pw_db, float phase_deg, float delay_s, float doppler, GainOut& g)
{
    g.fd = (double)doppler;
    g.valid = (p < d.P) && !(pw_db != pw_db);               // channel.py:260, dataset.py:258-261
    // generator_utils.py:35: float32 divide by 10, float32 pow.  The reference's value is itself a float32 libm/SVML result
    // (R7: within 1 ulp of correctly rounded); exp10f is within 2 ulp, i.e. <= 1.2e-7 on the amplitude.
    g.p_lin = exp10f(__fdiv_rn(pw_db, 10.0f));
    const float d2r = 0x1.1df46ap-6f;
    const float ph32 = __fmul_rn(phase_deg, d2r);           // np.deg2rad(phase) float32
    // complex64 exp(1j*x) = (cosf(x), sinf(x)) in the reference (R10, <= 1 ulp); sincosf here is within 2 ulp for |x| <= pi
    sincosf(ph32, &g.es, &g.ec);
    g.over = 0; g.wcyc = 0.0;
    if (kFreqDomain) {
        float dn = __fdiv_rn(delay_s, d.ts_f32);            // channel.py:183 (R11)
        g.over = dn >= d.n_f32;                             // :187 (R12)
        if (g.over) dn = d.n_f32;                           // :189
        g.wcyc = (double)dn * d.inv_n;
    }
    g.p_lin64 = (double)g.p_lin; g.ec64 = (double)g.ec; g.es64 = (double)g.es;
}

template <bool kFreqDomain>
__device__ __forceinline__ void prologue_gain_in(const DevDesc& d, int p, const PathIn& in, GainOut& g)
{
    prologue_gain_f32<kFreqDomain>(d, p, in.power, in.phase, in.delay, in.doppler, g);
}

template <bool kFreqDomain>
__device__ __forceinline__ void prologue_gain(const DevDesc& d, long long user, int p, GainOut& g)
{
    const long long off = user * (long long)d.ld + p;
    if (d.in_f64) {
        g.fd = d.doppler ? (double)d.doppler[off] : 0.0;
        // float64 path matrices: the whole gain chain is float64 in NumPy (generator_utils.py:35, channel.py:183-192)
        const double pw_db = reinterpret_cast<const double*>(d.power)[off];
        g.valid = (p < d.P) && !(pw_db != pw_db);
        g.p_lin64 = exp10(__ddiv_rn(pw_db, 10.0));
        g.p_lin = (float)g.p_lin64;
        sincos(__dmul_rn(reinterpret_cast<const double*>(d.phase)[off], kPi / 180.0), &g.es64, &g.ec64);
        g.es = (float)g.es64; g.ec = (float)g.ec64;
        g.over = 0; g.wcyc = 0.0;
        if (kFreqDomain) {
            double dn = __ddiv_rn(reinterpret_cast<const double*>(d.delay)[off], d.ts_f64);     // channel.py:183
            g.over = dn >= (double)d.N;                                                          // :187
            if (g.over) dn = (double)d.N;                                                        // :189
            g.wcyc = dn * d.inv_n;
        }
        return;
    }
    prologue_gain_f32<kFreqDomain>(d, p, d.power[off], d.phase[off], d.delay[off], d.doppler ? d.doppler[off] : 0.0f, g);
}

// kPlain: the caller knows that no FoV filter is set, both patterns are isotropic and the inputs are float32 (the common case gets a
// kernel instantiation without the dipole / float64-power code).
template <bool kFreqDomain, bool kPlain = false>
__device__ __forceinline__ void prologue_combine(const DevDesc& d, const SideOut& s0, const SideOut& s1, const GainOut& g, PathState& o)
{
    o.valid = g.valid; o.over = g.over; o.wcyc = g.wcyc; o.fd = g.fd;
    o.th[0] = s0.th; o.ph[0] = s0.ph; o.th[1] = s1.th; o.ph[1] = s1.ph;
    // ---- FoV (dataset.py:493-512)
    bool fov = true;
    if (!kPlain && d.fov_any) {
        if (d.fov_side[0]) fov = fov && in_fov(d, 0, s0.th, s0.ph);
        if (d.fov_side[1]) fov = fov && in_fov(d, 1, s1.th, s1.ph);
    }
    o.fov = fov;
    const bool ang_ok = !(s0.th != s0.th) && !(s1.th != s1.th) && !(s0.ph != s0.ph) && !(s1.ph != s1.ph);
    // ---- steering cycles per element step (geometry.py:99-101 with x == 0; dataset.py:393)
    o.u[0] = d.sp[0] * s0.ss; o.v[0] = d.sp[0] * s0.cc;
    o.u[1] = d.sp[1] * s1.ss; o.v[1] = d.sp[1] * s1.cc;
    // ---- power with element patterns (ant_patterns.py:167-168): float64 as soon as one side is a dipole
    const bool f64_power = !kPlain && (d.in_f64 || (d.pat[0] != DMK_PATTERN_ISOTROPIC) || (d.pat[1] != DMK_PATTERN_ISOTROPIC));
    double pw64 = g.p_lin64;
    if (!kPlain && (d.pat[0] != DMK_PATTERN_ISOTROPIC || d.pat[1] != DMK_PATTERN_ISOTROPIC)) {
        const double nan64 = __longlong_as_double(0x7ff8000000000000LL);
        const double gt = (d.pat[0] == DMK_PATTERN_HALFWAVE_DIPOLE) ? dipole_gain(fov ? s0.th : nan64) : 1.0;
        const double gr = (d.pat[1] == DMK_PATTERN_HALFWAVE_DIPOLE) ? dipole_gain(fov ? s1.th : nan64) : 1.0;
        pw64 = __dmul_rn(pw64, __dmul_rn(gt, gr));
    }
    o.pw = pw64;
    // ---- path gain
    if (kFreqDomain) {
        if (f64_power) {
            const double amp = g.over ? 0.0 : sqrt(__ddiv_rn(pw64, (double)d.N));   // channel.py:188,:192 (float64 branch)
            o.c = make_float2((float)(amp * g.ec64), (float)(amp * g.es64));
        } else {
            const float amp = g.over ? 0.0f : __fsqrt_rn(__fdiv_rn(g.p_lin, d.n_f32));  // :188, :192 (R9)
            o.c = make_float2(__fmul_rn(amp, g.ec), __fmul_rn(amp, g.es));
        }
    } else {
        if (f64_power) {
            const double amp = sqrt(pw64);                  // channel.py:286
            o.c = make_float2((float)(amp * g.ec64), (float)(amp * g.es64));
        } else {
            const float amp = __fsqrt_rn(g.p_lin);
            o.c = make_float2(__fmul_rn(amp, g.ec), __fmul_rn(amp, g.es));
        }
    }
    const bool c_ok = (o.c.x == o.c.x) && (o.c.y == o.c.y) && (o.wcyc == o.wcyc) && (o.fd == o.fd);
    // exact zeros where theta is NaN (geometry.py:65-80) or outside the FoV (dataset.py:508-511);
    // NaN gains are dropped by nansum (channel.py:283).
    o.contrib = o.valid && fov && ang_ok && c_ok && !(o.c.x == 0.0f && o.c.y == 0.0f);
}

__device__ __forceinline__ bool prologue_needs_angles(const DevDesc& d)
{
    return d.fov_any || d.pat[0] != DMK_PATTERN_ISOTROPIC || d.pat[1] != DMK_PATTERN_ISOTROPIC;
}

// Called by warp 7 (threads 224..255, not part of the chains): prefetch for the user 4 waves of CTAs ahead.
__device__ __forceinline__ void prefetch_user_rows_shifted(const DevDesc& d, long long user)
{
    const int t = threadIdx.x - 224;
    const long long ahead = user + 4LL * 2 * (d.n_sms > 0 ? d.n_sms : 148);
    if (t < 0 || t >= 7 || ahead >= d.n_users) return;
    const float* base = (t == 0) ? d.power : (t == 1) ? d.phase : (t == 2) ? d.delay : (t == 3) ? d.az[0] : (t == 4) ? d.el[0]
                      : (t == 5) ? d.az[1] : d.el[1];
    asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(base) + ahead * (long long)d.ld * (d.in_f64 ? 8 : 4)));
}

// Cooperative phase 1: warps 0, 1, 2 of the CTA run the three chains for path column `lane` of `user`.
struct PrologueScratch { SideOut side[2][kMaxPaths]; GainOut gain[kMaxPaths]; };

template <bool kFreqDomain>
__device__ __forceinline__ void cta_prologue_chains(const DevDesc& d, long long user, PrologueScratch& sc)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= d.P0 || warp > 2) return;
    if (warp < 2) {
        if (prologue_needs_angles(d)) prologue_side<true>(d, user, lane, warp, sc.side[warp][lane]);
        else                          prologue_side<false>(d, user, lane, warp, sc.side[warp][lane]);
    } else {
        prologue_gain<kFreqDomain>(d, user, lane, sc.gain[lane]);
    }
}

// Single-thread version (by-product kernel): always materialises the angles.
__device__ __forceinline__ void path_prologue_angles(const DevDesc& d, long long user, int p, PathState& o)
{
    SideOut s0, s1; GainOut g;
    prologue_side<true>(d, user, p, 0, s0);
    prologue_side<true>(d, user, p, 1, s1);
    prologue_gain<false>(d, user, p, g);
    prologue_combine<false>(d, s0, s1, g, o);
}

}  // namespace dmk
