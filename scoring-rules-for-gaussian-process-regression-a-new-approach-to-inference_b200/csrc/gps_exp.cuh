// exp(x) for x <= 0 as a short DFMA sequence (every exponential on the hot path is exp(-r^2/2) or
// exp(-z^2/2)): round-to-nearest range reduction x = n ln2 + r with the 1.5 * 2^52 shift, degree-13 Horner
// polynomial on |r| <= ln2 / 2 (truncation 4e-18 relative), scaling by 2^n through the exponent field.
// Relative error <= 2 ulp against the correctly rounded value (tests/test_gpu_stages.py compares the Gram
// kernels that use it with numpy at 1e-15); ~21 instructions instead of the ~35 of the library exp, no
// special cases.  Arguments below -700 return 0 (exp(-700) = 1e-304).
#pragma once

// polynomial coefficients 1/13! .. 1/2! and the reduction constants live in constant memory: DFMA takes a
// 64-bit constant-bank operand directly, whereas a literal double with a non-zero low word costs two moves per use
// (a profile of the fused FITC pass 1 showed 19 % of its issue slots going to those moves)
__constant__ double gps_exp_c[16] = {
    1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
    2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
    8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
    1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 6755399441055744.0};

__device__ __forceinline__ double exp_neg(double x) {
  const double SHIFT = gps_exp_c[15];                   // 1.5 * 2^52
  const double tt = fma(x, gps_exp_c[12], SHIFT);       // x log2(e)
  const int n = __double2loint(tt);
  const double fn = tt - SHIFT;
  double r = fma(fn, gps_exp_c[13], x);                 // -ln2 high part (trailing zeros: fn * hi is exact)
  r = fma(fn, gps_exp_c[14], r);                        // -ln2 low part
  double p = gps_exp_c[0];                              // 1/13!
#pragma unroll
  for (int k = 1; k < 12; ++k) p = fma(p, r, gps_exp_c[k]);   // 1/12! .. 1/2!
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return x < -700.0 ? 0.0 : res;
}
