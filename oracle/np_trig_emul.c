/*
 * oracle/np_trig_emul.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Scalar C restatement of NumPy's float32 SIMD sin/cos (NumPy 2.3.5,
 * numpy/_core/src/umath/loops_trigonometric.dispatch.*; NumPy is a third-party
 * dependency of the reference -- pyproject.toml:24 "numpy>=1.19.5" -- and is not
 * under /root/reference).  The reference calls it at
 *   deepmimo/generator/geometry.py:301-302   sin_theta = np.sin(theta); cos_theta = np.cos(theta)
 * on float32 arrays, and the rounding of those two values is amplified by up to
 * kd*(M-1) ~ 198 in the steering phase, so the
 * CUDA path reproduces it bit-for-bit; this file is the CPU twin used to
 *   (1) check the restatement exhaustively against np.sin/np.cos on the oracle
 *       host (tests/test_np_trig_emul.py), and
 *   (2) check the device function on the GPU box (tests/test_gpu_trig.py).
 *
 * Published algorithm (NumPy source comments): Cody-Waite 3-constant reduction
 * x* = x - q*pi/2, q = rint(x*2/pi) via fma(x, 2/pi, 1.5*2^23) - 1.5*2^23 (fused -- the unfused
 * form differs from NumPy 2.3.5 at x = 0x1.f6a7a4p+1f, found by exhaustive search); degree-8
 * cosine / degree-9 sine minimax polynomials in x*^2 evaluated with FMA; quadrant
 * select and sign flip.  Valid for |x| <= 71476.0625f (cos) -- larger arguments
 * go to libm in NumPy and are outside this path's domain (angles in radians).
 *
 * Build: gcc -O2 -mfma -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: every rounding below is deliberate.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static inline float np_trig_core(float x, int want_cos)
{
    if (x != x) return NAN;
    const float two_over_pi = 0x1.45f306p-1f;
    const float magic       = 0x1.800000p+23f;
    float q = fmaf(x, two_over_pi, magic);   /* fused: NumPy 2.3.5 (Highway MulAdd) rounds once here */
    q = q - magic;
    float r = fmaf(q, -0x1.921fb0p+00f, x);
    r = fmaf(q, -0x1.5110b4p-22f, r);
    r = fmaf(q, -0x1.846988p-48f, r);
    float r2 = r * r;
    /* cosine polynomial */
    float c = fmaf(0x1.98e616p-16f, r2, -0x1.6c06dcp-10f);
    c = fmaf(c, r2, 0x1.55553cp-05f);
    c = fmaf(c, r2, -0x1.000000p-01f);
    c = fmaf(c, r2, 0x1.000000p+00f);
    /* sine polynomial */
    float s = fmaf(0x1.7d3bbcp-19f, r2, -0x1.a06bbap-13f);
    s = fmaf(s, r2, 0x1.11119ap-07f);
    s = fmaf(s, r2, -0x1.555556p-03f);
    s = fmaf(s, r2, 0.0f);
    s = fmaf(s, r, r);
    int iq = (int)q;            /* q is already integral */
    if (want_cos) iq += 1;
    float res = ((iq & 1) == 0) ? s : c;
    if ((iq & 2) == 2) res = 0.0f - res;
    return res;
}

float np_sinf_emul(float x) { return np_trig_core(x, 0); }
float np_cosf_emul(float x) { return np_trig_core(x, 1); }

void np_sinf_emul_array(const float *x, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) out[i] = np_trig_core(x[i], 0);
}
void np_cosf_emul_array(const float *x, float *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) out[i] = np_trig_core(x[i], 1);
}
