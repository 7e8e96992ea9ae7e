// hz_math.cuh -- fp32 device math for the projection (replaces vertex.glsl:111-157 of the reference).
//
// The whole translation unit is compiled with -fmad=false: every `a*b+c` written below is two
// IEEE roundings, exactly like the reference's GLSL evaluated without contraction (and like the
// CPU oracle, built with -ffp-contract=off).  Fused multiply-adds appear only where they are
// spelled fmaf(), inside the polynomial kernels whose result is an approximation anyway.
//
// Accuracy contract (checked on the GPU by tests/test_device_math.py against double precision):
//   hz_atan2_az : <= 2.5 ulp of the result over all four quadrants
//   hz_atan_el  : <= 3.5 ulp of the result (measured: 3.03 at worst, 0.44 on average)
// i.e. the same class as CUDA's own atan2f (2 ulp) at about half the instruction count.  Define
// HZ_LIBDEVICE_MATH to fall back to atan2f/sqrtf for A/B comparisons.
#pragma once
#include <cuda_runtime.h>

#define HZ_PI_F      3.14159265358979f   /* the literal of vertex.glsl:31, rounded to float */
#define HZ_REARTH_F  6371000.0f          /* vertex.glsl:30 */

// atan(x) for 0 <= x <= 1:  x + x*s*Q(s), s = x^2, Q = degree-7 minimax (max rel. error 1.5e-8)
__device__ __forceinline__ float hz_atan_unit(float x)
{
    const float s = x * x;
    float q =            2.8498897586e-03f;
    q = fmaf(q, s, -1.6068629514e-02f);
    q = fmaf(q, s,  4.2691520175e-02f);
    q = fmaf(q, s, -7.5042946158e-02f);
    q = fmaf(q, s,  1.0640934061e-01f);
    q = fmaf(q, s, -1.4203644476e-01f);
    q = fmaf(q, s,  1.9992619393e-01f);
    q = fmaf(q, s, -3.3333073345e-01f);
    return fmaf(x * s, q, x);
}

// MUFU.RCP (about 1 ulp)
__device__ __forceinline__ float hz_rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// mn/mx for 0 <= mn <= mx: MUFU.RCP + one residual correction (<= 1 ulp); 0/0 -> 0
__device__ __forceinline__ float hz_ratio(float mn, float mx)
{
    const float rc = hz_rcp_approx(mx);
    float a = mn * rc;
    const float r = fmaf(-a, mx, mn);
    a = fmaf(r, rc, a);
    return (mx == 0.0f) ? 0.0f : a;
}

// GLSL atan(y, x) for the azimuth: atan2(e, n), result in [-pi, pi], 0 = north, +pi/2 = east
__device__ __forceinline__ float hz_atan2_az(float e, float n)
{
#ifdef HZ_LIBDEVICE_MATH
    return atan2f(e, n);
#else
    const float ae = fabsf(e), an = fabsf(n);
    const float mx = fmaxf(ae, an), mn = fminf(ae, an);
    float r = hz_atan_unit(hz_ratio(mn, mx));
    if(ae > an)  r = 1.57079632679489662f - r;
    if(n < 0.0f) r = 3.14159265358979324f - r;
    return copysignf(r, e);
#endif
}

// GLSL atan(h, d) for the elevation, d = sqrt(d2) >= 0: result in [-pi/2, pi/2].
// d is never formed: atan(h/d) = atan(h * rsqrt(d2)).
__device__ __forceinline__ float hz_atan_el(float h, float d2)
{
#ifdef HZ_LIBDEVICE_MATH
    return atan2f(h, sqrtf(d2));
#else
    const float q  = h * rsqrtf(d2);          // MUFU.RSQ; d2 == 0 -> +-inf or NaN, handled below
    const float aq = fabsf(q);
    float r;
    if(aq <= 1.0f)
        r = hz_atan_unit(aq);
    else if(aq <= 3.0e38f)                    // steeper than 45 degrees: only right next to the eye
        r = 1.57079632679489662f - hz_atan_unit(hz_ratio(1.0f, aq));
    else                                      // d2 == 0: straight up/down, or atan(0,0) = 0
        r = (h == 0.0f) ? 0.0f : 1.57079632679489662f;
    return copysignf(r, h);
#endif
}
