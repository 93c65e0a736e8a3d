/* oracle_mc.c -- plain-C (OpenMP) port of the reference Monte Carlo chain pricer.  TEST INFRASTRUCTURE ONLY:
 * used (a) as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) and (b) as a large-N checker
 * of the fp64 GPU path in tests/.  Never linked into or called by the product (stochvolmodels_b200).
 *
 * Follows, expression by expression (compiled with -ffp-contract=off):
 *   simulate_logsv_x_vol_terminal   /root/reference/src/stochvolmodels/pricers/logsv_pricer.py:1021-1047
 *   logsv_mc_chain_pricer           pricers/logsv_pricer.py:806-867
 *   simulate_heston_x_vol_terminal  pricers/heston_pricer.py:366-381
 *   heston_mc_chain_pricer          pricers/heston_pricer.py:285-331
 *   compute_mc_vars_payoff          utils/mc_payoffs.py:61-88
 *   set_time_grid                   utils/funcs.py:44-47
 * The only departure from the reference is the source of the Gaussians: the reference draws them from Numba's global
 * MT19937 and materialises W0/W1[steps, paths]; this port draws the SAME Philox4x32-10 + Box-Muller stream as the device
 * (stochvolmodels_b200/csrc/philox.cuh, restated in oracle/mc.py) per path, so that GPU and CPU results can be compared
 * path-for-path, and keeps per-path state in registers (which only makes the CPU baseline faster than the reference).
 * Pinned through tests/test_oracle_golden.py::test_c_port_matches_numpy_oracle (numpy oracle <- golden <- reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

static inline double u52(uint32_t hi, uint32_t lo) {
  const uint64_t bits = ((uint64_t)0x3FF << 52) | ((uint64_t)(hi >> 12) << 32) | (uint64_t)lo;
  double d; memcpy(&d, &bits, 8); return d;
}

/* normals of (path, slice, step): gauss64 != 0 -> one call per step; else the float Box-Muller, one call per two steps */
static inline void step_normals(uint64_t seed, uint64_t path, uint32_t slice, uint32_t step, int gauss64, double* z0, double* z1) {
  uint32_t c[4] = {(uint32_t)path, (uint32_t)(path >> 32), gauss64 ? step : step >> 1, slice};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  if (gauss64) {
    const double u1 = 2.0 - u52(c[0], c[1] | 1u), u2 = u52(c[2], c[3]) - 1.0;   /* lowest bit forced: u1 < 1 (gauss64.cuh) */
    const double rad = sqrt(-2.0 * log(u1));
    *z0 = rad * cos(2.0 * M_PI * u2);
    *z1 = rad * sin(2.0 * M_PI * u2);
  } else {
    const uint32_t ra = (step & 1) ? c[2] : c[0], rb = (step & 1) ? c[3] : c[1];
    const float u1 = fmaf((float)ra, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float rad = sqrtf(-2.0f * logf(u1));
    const float ang = (float)(int32_t)rb * 1.4629180792671596e-09f;
    *z0 = (double)(rad * cosf(ang));
    *z1 = (double)(rad * sinf(ang));
  }
}

static void time_grid(double ttm, int n, int* S, double* dt) {
  *S = (int)(ttm * (double)n) + 1;
  *dt = ttm / (double)(*S);
}

/* payoffs of one slice: utils/mc_payoffs.py:61-88 (nanmean / population nanstd, SE / sqrt(N)) */
static void payoffs(const double* x, const double* q, long long N, double ttm, double F, const double* strikes, const int8_t* types,
                    int J, double df, int variable_type, double* prices, double* stderrs) {
  double sum = 0.0, cnt = 0.0;
#pragma omp parallel for reduction(+ : sum, cnt) schedule(static)
  for (long long i = 0; i < N; ++i) {
    const double s = F * exp(x[i]);
    if (s == s) { sum += s; cnt += 1.0; }
  }
  const double corr = sum / cnt - F;
  for (int j = 0; j < J; ++j) {
    const double K = strikes[j];
    const int ty = types[j];
    double s1 = 0.0, n1 = 0.0;
#pragma omp parallel for reduction(+ : s1, n1) schedule(static)
    for (long long i = 0; i < N; ++i) {
      const double spot = F * exp(x[i]) - corr;
      const double u = variable_type == 2 ? q[i] / ttm : spot;
      double pay = (ty & 1) ? (u < K ? K - u : 0.0) : (u > K ? u - K : 0.0);
      if (ty >= 2) pay = pay / spot;
      if (pay == pay) { s1 += pay; n1 += 1.0; }
    }
    const double mean = s1 / n1;
    double s2 = 0.0;
#pragma omp parallel for reduction(+ : s2) schedule(static)
    for (long long i = 0; i < N; ++i) {
      const double spot = F * exp(x[i]) - corr;
      const double u = variable_type == 2 ? q[i] / ttm : spot;
      double pay = (ty & 1) ? (u < K ? K - u : 0.0) : (u > K ? u - K : 0.0);
      if (ty >= 2) pay = pay / spot;
      if (pay == pay) s2 += (pay - mean) * (pay - mean);
    }
    prices[j] = df * mean;
    stderrs[j] = df * sqrt(s2 / n1) / sqrt((double)N);
  }
}

/* model: 0 = LogSV (params = sigma0, theta, kappa1, kappa2, beta, volvol), 1 = Heston (v0, theta, kappa, rho, volvol).
 * states_out (optional): 3*N doubles (x, sigma|var, qvar) after the LAST slice.  returns 0, or -1 on allocation failure. */
int oracle_mc_chain(int model, const double* params, int M, const double* ttms, const double* forwards, const double* discfactors,
                    const double* etas, const int* offsets, const double* strikes, const int8_t* types, long long nb_path,
                    long long path_offset, int nb_steps_per_year, int is_spot, int variable_type, uint64_t seed, int gauss64,
                    int nthreads, double* prices, double* stderrs, double* states_out) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  double* x = (double*)malloc(sizeof(double) * 3 * (size_t)nb_path);
  if (!x) return -1;
  double *v = x + nb_path, *q = x + 2 * nb_path;
  const double v_init = params[0];
  double t0 = 0.0;
  for (int m = 0; m < M; ++m) {
    int S; double dt;
    time_grid(ttms[m] - t0, nb_steps_per_year, &S, &dt);
    t0 = ttms[m];
    const double sdt = sqrt(dt);
    if (model == 0) {
      const double theta = params[1], kappa1 = params[2], kappa2 = params[3], beta = params[4], volvol = params[5];
      const double eta = etas ? etas[m] : 1.0;
      const double alpha = is_spot ? -1.0 : 1.0, adj = is_spot ? 0.0 : beta * eta;
      const double vartheta2 = beta * beta + volvol * volvol, eta2 = eta * eta;
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < nb_path; ++i) {
        double xi = m ? x[i] : 0.0, si = m ? v[i] : v_init, qi = m ? q[i] : 0.0;
        double L = log(si);
        for (int s = 0; s < S; ++s) {
          double z0, z1;
          step_normals(seed, (uint64_t)(path_offset + i), (uint32_t)m, (uint32_t)s, gauss64, &z0, &z1);
          const double w0 = sdt * z0, w1 = sdt * z1;
          const double s2dt = eta2 * si * si * dt;
          xi = xi + alpha * 0.5 * s2dt + eta * si * w0;
          L = L + ((kappa1 * theta / si - kappa1) + kappa2 * (theta - si) + adj * si - 0.5 * vartheta2) * dt + beta * w0 + volvol * w1;
          si = exp(L);
          qi = qi + 0.5 * (s2dt + eta2 * si * si * dt);
        }
        x[i] = xi; v[i] = si; q[i] = qi;
      }
    } else {
      const double theta = params[1], kappa = params[2], rho = params[3], volvol = params[4];
      const double rho_1 = sqrt(1.0 - rho * rho);
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < nb_path; ++i) {
        double xi = m ? x[i] : 0.0, vi = m ? v[i] : v_init, qi = m ? q[i] : 0.0;
        for (int s = 0; s < S; ++s) {
          double z0, z1;
          step_normals(seed, (uint64_t)(path_offset + i), (uint32_t)m, (uint32_t)s, gauss64, &z0, &z1);
          const double w0 = sdt * z0, w1 = sdt * z1;
          const double sig = sqrt(vi), vdt = vi * dt;
          xi = xi - 0.5 * vdt + sig * w0;
          qi = qi + vdt;
          vi = vi + kappa * (theta - vi) * dt + sig * volvol * (rho * w0 + rho_1 * w1);
          vi = vi < 1e-4 ? 1e-4 : vi;
        }
        x[i] = xi; v[i] = vi; q[i] = qi;
      }
    }
    const int J = offsets[m + 1] - offsets[m];
    if (J > 0 && prices)
      payoffs(x, q, nb_path, ttms[m], forwards[m], strikes + offsets[m], types + offsets[m], J, discfactors[m], variable_type,
              prices + (offsets[m] - offsets[0]), stderrs + (offsets[m] - offsets[0]));
  }
  if (states_out) memcpy(states_out, x, sizeof(double) * 3 * (size_t)nb_path);
  free(x);
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
