// black.cuh -- Black-76 implied volatility of one quote (shared by ivol_kernels.cu and the batched chain pricers of mgf_kernels.cu).
// Bracketed bisection on the call-equivalent undiscounted price: 80 halvings of [1e-8, 10], the same sequence as the checker
// oracle/bsm.py, so the two agree to the last bits.  Prices outside the no-arbitrage bounds give NaN.
#pragma once
#include <cmath>

#include "../../include/b200sv.h"

namespace b200sv {

__device__ __forceinline__ double black_call(double F, double K, double sdev) {
  const double d1 = log(F / K) / sdev + 0.5 * sdev;
  return F * normcdf(d1) - K * normcdf(d1 - sdev);
}

__device__ __forceinline__ double black_implied_vol(double forward, double strike, double ttm, double discfactor, double price, int type) {
  const double p = price / discfactor;
  const bool is_call = (type == B200SV_CALL || type == B200SV_INV_CALL);
  const double c = is_call ? p : p + (forward - strike);            // put-call parity: work on the call
  const double intrinsic = fmax(forward - strike, 0.0);
  const bool ok = (c > intrinsic) && (c < forward) && isfinite(c);
  const double srt = sqrt(ttm);
  double a = 1e-8, b = 10.0;
  for (int it = 0; it < 80; ++it) {
    const double mid = 0.5 * (a + b);
    const bool up = black_call(forward, strike, mid * srt) < c;
    a = up ? mid : a;
    b = up ? b : mid;
  }
  return ok ? 0.5 * (a + b) : NAN;
}

}  // namespace b200sv
