// fastmath64.cuh -- fp64 helpers for the fused stepper.
//
// exp_pair(L): exp(L) and exp(-L) from ONE range reduction and ONE even/odd polynomial evaluation
//   L = k ln2 + r, |r| <= ln2/2;   exp(r) = E(r^2) + r O(r^2),  exp(-r) = E(r^2) - r O(r^2)
// with the Taylor coefficients up to r^13 (truncation < 2e-17 relative on |r| <= 0.3466), scaled by 2^(+-k) through the
// exponent field.  ~20 fp64-pipe instructions for both values, against ~2 x 25 for two libdevice exp() calls or
// exp() + an IEEE division.  Max observed error 1 ulp (tests/test_gpu_mc.py::test_exp_pair).  k is clamped to
// [-1000, 1000]: exp saturates at 2^+-1000 instead of overflowing -- |log sigma| > 693 has no meaning for a volatility.
#pragma once

namespace b200sv {

__device__ __forceinline__ void exp_pair(double L, double& ep, double& em) {
  const double SHIFT = 6755399441055744.0;                     // 2^52 + 2^51: round-to-nearest-integer trick
  const double t = fma(L, 1.4426950408889634074, SHIFT);       // L * log2(e)
  int k = __double2loint(t);
  const double kf = t - SHIFT;
  double r = fma(kf, -6.93147180369123816490e-01, L);          // ln2 hi
  r = fma(kf, -1.90821492927058770002e-10, r);                 // ln2 lo
  const double r2 = r * r;
  // even part: 1 + r2/2! + r2^2/4! + ... + r2^6/12!
  double E = fma(r2, 2.08767569878680989792e-09, 2.75573192239858906526e-07);   // 1/12!, 1/10!
  E = fma(E, r2, 2.48015873015873015873e-05);                                   // 1/8!
  E = fma(E, r2, 1.38888888888888888889e-03);                                   // 1/6!
  E = fma(E, r2, 4.16666666666666666667e-02);                                   // 1/4!
  E = fma(E, r2, 0.5);
  E = fma(E, r2, 1.0);
  // odd part / r: 1 + r2/3! + ... + r2^6/13!
  double O = fma(r2, 1.60590438368216145994e-10, 2.50521083854417187751e-08);   // 1/13!, 1/11!
  O = fma(O, r2, 2.75573192239858906526e-06);                                   // 1/9!
  O = fma(O, r2, 1.98412698412698412698e-04);                                   // 1/7!
  O = fma(O, r2, 8.33333333333333333333e-03);                                   // 1/5!
  O = fma(O, r2, 1.66666666666666666667e-01);                                   // 1/3!
  O = fma(O, r2, 1.0);
  const double rO = r * O;
  const double p = E + rO, m = E - rO;                          // both in [0.70, 1.42]
  k = max(-1000, min(1000, k));
  ep = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  em = __hiloint2double(__double2hiint(m) - (k << 20), __double2loint(m));
}

}  // namespace b200sv
