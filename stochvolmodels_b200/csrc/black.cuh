// black.cuh -- Black-76 implied volatility of one quote (shared by ivol_kernels.cu and the batched chain pricers of mgf_kernels.cu /
// the fused finalize kernel of mc_kernels.cu).
// Safeguarded Newton on the call-equivalent undiscounted price in total-volatility space s = vol * sqrt(ttm): start at the inflection
// point s = sqrt(2 |ln(F/K)|) (Manaster-Koehler: from there Newton is monotone), keep the bracket [1e-8, 10] * sqrt(ttm) updated by the
// sign of every residual and fall back to its midpoint whenever a Newton step leaves it or the vega underflows.  5-9 iterations of two
// normcdf + one exp instead of the 80 halvings of round 1 (the fused inversion was 50 of the 350 us of a batched chain call).  The
// checker oracle/bsm.py stays the plain 80-step bisection (an independent method): the two agree to the conditioning of the quote,
// 1e-16 * F / vega.  Prices outside the no-arbitrage bounds give NaN; roots outside the bracket return its end (as the bisection did).
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
  if (!ok) return NAN;
  const double srt = sqrt(ttm);
  const double x = log(forward / strike);
  double a = 1e-8 * srt, b = 10.0 * srt;
  double s = sqrt(2.0 * fabs(x));
  if (!(s > a && s < b)) s = fmin(fmax(2.5066282746310002 * c / forward, 2.0 * a), 0.5 * b);   // at the money: Brenner-Subrahmanyam
  for (int it = 0; it < 64; ++it) {
    const double d1 = x / s + 0.5 * s;
    const double f = forward * normcdf(d1) - strike * normcdf(d1 - s) - c;
    const double vega = forward * 0.3989422804014327 * exp(-0.5 * d1 * d1);
    if (f < 0.0) a = s; else b = s;
    double sn = s - f / vega;
    if (!(sn > a && sn < b)) sn = 0.5 * (a + b);                     // also catches NaN / inf from a vanishing vega
    const double step = fabs(sn - s);
    s = sn;
    if (step <= 4.5e-16 * s || b - a <= 4.5e-16 * b) break;     // within 2 ulp
  }
  return s / srt;
}

}  // namespace b200sv
