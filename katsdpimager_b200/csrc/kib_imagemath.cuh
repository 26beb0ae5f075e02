// Image-plane arithmetic shared by kib_image.cu and kib_gridfft.cu: separately rounded
// direction cosines and the W-term rotation (see the header comment of kib_image.cu).
#pragma once
#include "kib_common.cuh"

namespace kib {

template <typename Real> __device__ __forceinline__ Real mul_rn(Real a, Real b);
template <> __device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <typename Real> __device__ __forceinline__ Real add_rn(Real a, Real b);
template <> __device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// exp(2 pi i * sign * w * (n - 1)) with the argument reduced in double precision
template <typename Real>
__device__ __forceinline__ void w_rotation(Real n, double w, Real *c, Real *s);

// Single precision: the product w * (n - 1) is formed in float-float arithmetic (w split
// on the host into hi + lo, product error recovered with an FMA) and reduced to
// [-0.5, 0.5] turns with the 1.5 * 2^23 rounding constant; sin/cos come from degree-3
// minimax polynomials in f^2 on the reduced quadrant (|f| <= 1/4 half-turn, abs error
// < 1e-7).  Nothing here touches the XU (MUFU / conversion) pipe, which an earlier
// version with double-precision reduction and sincospif() saturated (67 % busy).
// Valid for |w (n - 1)| < 2^22 turns.
template <>
__device__ __forceinline__ void w_rotation<float>(float n, double w, float *c, float *s)
{
    const float MAGIC = 12582912.0f;                    // 1.5 * 2^23
    const float w_hi = (float) w;                       // loop-invariant, hoisted by the compiler
    const float w_lo = (float) (w - (double) w_hi);
    const float nm1 = __fadd_rn(n, -1.0f);              // exact (n in [0.5, 1])
    const float p = __fmul_rn(nm1, w_hi);
    const float e = __fmaf_rn(nm1, w_hi, -p);           // exact rounding error of p
    const float lo = __fmaf_rn(nm1, w_lo, e);
    const float k = __fadd_rn(__fadd_rn(p, MAGIC), -MAGIC);     // rint(p)
    const float t = 2.0f * __fadd_rn(__fadd_rn(p, -k), lo);     // phase in half-turns, |t| <= 1
    // quadrant: t = j / 2 + f, |f| <= 1/4
    const float jm = __fadd_rn(2.0f * t, MAGIC);
    const int j = __float_as_int(jm);
    const float f = __fmaf_rn(__fadd_rn(jm, -MAGIC), -0.5f, t);
    const float z = f * f;
    float sn = __fmaf_rn(z, -0.5893668532371521f, 2.5497970581054688f);
    sn = __fmaf_rn(z, sn, -5.167708873748779f);
    sn = __fmaf_rn(z, sn, 3.1415927410125732f) * f;
    float cs = __fmaf_rn(z, -1.3069566488265991f, 4.0576629638671875f);
    cs = __fmaf_rn(z, cs, -4.93479061126709f);
    cs = __fmaf_rn(z, cs, 1.0f);
    // rotate by j quarter turns: j & 1 swaps, signs from j & 2 and (j + 1) & 2
    const float a = (j & 1) ? sn : cs;
    const float b = (j & 1) ? cs : sn;
    *c = __int_as_float(__float_as_int(a) ^ (((j + 1) & 2) << 30));
    *s = __int_as_float(__float_as_int(b) ^ ((j & 2) << 30));
}

template <>
__device__ __forceinline__ void w_rotation<double>(double n, double w, double *c, double *s)
{
    const double phase = w * (n - 1.0);
    const double r = phase - rint(phase);
    sincospi(2.0 * r, s, c);
}

}  // namespace kib
