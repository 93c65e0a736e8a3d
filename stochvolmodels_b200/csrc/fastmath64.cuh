// fastmath64.cuh -- fp64 helpers for the fused stepper.
//
// exp_pair(L): exp(L) and exp(-L) from ONE range reduction and ONE even/odd polynomial evaluation
//   L = k ln2 + r, |r| <= ln2/2;   exp(r) = E(r^2) + r O(r^2),  exp(-r) = E(r^2) - r O(r^2)
// with the Taylor coefficients up to r^13 (truncation < 2e-17 relative on |r| <= 0.3466), scaled by 2^(+-k) through the
// exponent field.  ~19 fp64-pipe instructions for both values, against ~2 x 25 for two libdevice exp() calls or
// exp() + an IEEE division.  Max observed error 2 ulp (tests/test_gpu_mc.py::test_exp_pair_accuracy).
// Contract: L finite.  |L| > 700 saturates at exp(+-700) (callers guard with clamp_log); k is additionally clamped to
// [-1020, 1020] so the exponent arithmetic can never wrap.  The polynomial coefficients live in constant memory so each
// DFMA takes them as a c[bank][offset] operand instead of re-materialising 64-bit immediates every iteration.
#pragma once

namespace b200sv {

__constant__ double kExpEven[6] = {2.75573192239858906526e-07, 2.48015873015873015873e-05, 1.38888888888888888889e-03,
                                   4.16666666666666666667e-02, 0.5, 1.0};                       // 1/10!, 1/8!, 1/6!, 1/4!, 1/2!, 1
__constant__ double kExpOdd[6] = {2.50521083854417187751e-08, 2.75573192239858906526e-06, 1.98412698412698412698e-04,
                                  8.33333333333333333333e-03, 1.66666666666666666667e-01, 1.0}; // 1/11!, 1/9!, 1/7!, 1/5!, 1/3!, 1

// cheap guard: keep |L| <= 700 using one integer compare on the high word in the common case
__device__ __forceinline__ double clamp_log(double L) {
  if ((__double2hiint(L) & 0x7fffffff) > 0x4085e000) L = copysign(700.0, L);   // |L| > 700 (also catches inf / NaN)
  return L;
}

__device__ __forceinline__ void exp_pair(double L, double& ep, double& em) {
  const double SHIFT = 6755399441055744.0;                     // 2^52 + 2^51: round-to-nearest-integer trick
  const double t = fma(L, 1.4426950408889634074, SHIFT);       // L * log2(e)
  int k = __double2loint(t);
  const double kf = t - SHIFT;
  double r = fma(kf, -6.93147180369123816490e-01, L);          // ln2 hi
  r = fma(kf, -1.90821492927058770002e-10, r);                 // ln2 lo
  const double r2 = r * r;
  double E = fma(r2, 2.08767569878680989792e-09, kExpEven[0]);  // 1/12!
  double O = fma(r2, 1.60590438368216145994e-10, kExpOdd[0]);   // 1/13!
#pragma unroll
  for (int i = 1; i < 6; ++i) {
    E = fma(E, r2, kExpEven[i]);
    O = fma(O, r2, kExpOdd[i]);
  }
  const double p = fma(r, O, E), m = fma(-r, O, E);            // both in [0.70, 1.42]
  k = max(-1020, min(1020, k));
  ep = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  em = __hiloint2double(__double2hiint(m) - (k << 20), __double2loint(m));
}

}  // namespace b200sv
