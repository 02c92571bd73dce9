// exp(x) for x <= 0 as a short DFMA sequence (every exponential on the hot path is exp(-r^2/2) or
// exp(-z^2/2)): round-to-nearest range reduction x = n ln2 + r with the 1.5 * 2^52 shift, degree-13 Horner
// polynomial on |r| <= ln2 / 2 (truncation 4e-18 relative), scaling by 2^n through the exponent field.
// Relative error <= 2 ulp against the correctly rounded value (tests/test_gpu_stages.py compares the Gram
// kernels that use it with numpy at 1e-15); ~21 instructions instead of the ~35 of the library exp, no
// special cases.  Arguments below -700 return 0 (exp(-700) = 1e-304).
#pragma once

__device__ __forceinline__ double exp_neg(double x) {
  const double SHIFT = 6755399441055744.0;              // 1.5 * 2^52
  const double tt = fma(x, 1.4426950408889634074, SHIFT);
  const int n = __double2loint(tt);
  const double fn = tt - SHIFT;
  double r = fma(fn, -6.93147180369123816490e-01, x);   // ln2 high part (trailing zeros: fn * hi is exact)
  r = fma(fn, -1.90821492927058770002e-10, r);          // ln2 low part
  double p = 1.6059043836821613e-10;                    // 1/13!
  p = fma(p, r, 2.08767569878681e-09);                  // 1/12!
  p = fma(p, r, 2.505210838544172e-08);                 // 1/11!
  p = fma(p, r, 2.755731922398589e-07);                 // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);                // 1/9!
  p = fma(p, r, 2.48015873015873e-05);                  // 1/8!
  p = fma(p, r, 1.984126984126984e-04);                 // 1/7!
  p = fma(p, r, 1.388888888888889e-03);                 // 1/6!
  p = fma(p, r, 8.333333333333333e-03);                 // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);                // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);                // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double res = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return x < -700.0 ? 0.0 : res;
}
