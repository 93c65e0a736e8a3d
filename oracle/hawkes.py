"""CPU restatement (numpy) of the reference's Hawkes jump-diffusion Monte Carlo.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).
Pinned by tests/golden/hawkes_mc.npz (tests/golden/make_golden.py --only-hawkes ran the unmodified reference).

Follows /root/reference/src/stochvolmodels/pricers/hawkes_jd_pricer.py:
  simulate_hawkesjd_terminal   :718-779   1800 steps per year; per step one scaled normal, two exponential clocks -ln(U)/dt compared with
                                          the two intensities, two shifted-exponential jump sizes; Euler update of (x, lambda_p, lambda_m)
  hawkesjd_mc_chain_pricer     :644-715   slices chained on the terminal state; payoffs = utils/mc_payoffs.py with x as every argument
``draw_inputs`` re-draws what the reference draws from numpy's GLOBAL legacy generator after ``np.random.seed(seed)``: one block of shape
(S, N) each, in the order W0 (normal), U_P, U_M (uniform(1e-16, 1)), J_P, J_M (exponential).
"""
from __future__ import annotations

import numpy as np

from . import mc as _mc

KEYS = ("mu", "sigma", "shift_p", "mean_p", "shift_m", "mean_m", "lambda_p", "theta_p", "kappa_p", "beta1_p", "beta2_p", "lambda_m", "theta_m",
        "kappa_m", "beta1_m", "beta2_m")
STEPS_PER_YEAR = 5 * 360      # :752


def draw_inputs(rng, ttm, nb_path, shift_p, mean_p, shift_m, mean_m):
    """(W0, U_P, U_M, J_P, J_M, dt) exactly as :752-757 forms them from the generator ``rng`` (a RandomState)"""
    S, dt = _mc.set_time_grid(ttm, STEPS_PER_YEAR)
    W0 = np.sqrt(dt) * rng.normal(0, 1, size=(S, nb_path))
    U_P = -np.log(rng.uniform(low=1e-16, high=1.0, size=(S, nb_path))) / dt
    U_M = -np.log(rng.uniform(low=1e-16, high=1.0, size=(S, nb_path))) / dt
    J_P = shift_p + rng.exponential(scale=mean_p, size=(S, nb_path))
    J_M = shift_m - rng.exponential(scale=-mean_m, size=(S, nb_path))
    return W0, U_P, U_M, J_P, J_M, dt


def step_fixed(x, lp, lm, W0, U_P, U_M, J_P, J_M, dt, mu, sigma, shift_p, mean_p, shift_m, mean_m, theta_p, kappa_p, beta1_p, beta2_p, theta_m,
               kappa_m, beta1_m, beta2_m, **_):
    """the step loop :764-777 on caller-supplied inputs"""
    comp_p = dt * (np.exp(shift_p) / (1.0 - mean_p) - 1.0)
    comp_m = dt * (np.exp(shift_m) / (1.0 - mean_m) - 1.0)
    drift_dt = (mu - 0.5 * sigma * sigma) * dt
    x, lp, lm = x.copy(), lp.copy(), lm.copy()
    for w0, u_p, u_m, j_p, j_m in zip(W0, U_P, U_M, J_P, J_M):
        diffusion = drift_dt - comp_p * lp - comp_m * lm + sigma * w0
        jump_p = np.where(lp > u_p, j_p, 0.0)
        jump_m = np.where(lm > u_m, j_m, 0.0)
        x = x + diffusion + jump_p + jump_m
        load_p = beta1_p * jump_p + beta2_p * jump_m
        load_m = beta1_m * jump_p + beta2_m * jump_m
        lp = lp + kappa_p * (theta_p - lp) * dt + load_p
        lm = lm + kappa_m * (theta_m - lm) * dt + load_m
    return x, lp, lm


def chain_prices(params: dict, ttms, forwards, discfactors, strikes_ttms, types_ttms, nb_path, rng=None, inputs=None):
    """hawkesjd_mc_chain_pricer (:644-715); ``inputs`` = list of per-slice (W0, U_P, U_M, J_P, J_M, dt) or drawn from ``rng``"""
    x, lp, lm = np.zeros(nb_path), params["lambda_p"] * np.ones(nb_path), params["lambda_m"] * np.ones(nb_path)
    t0, prices, stds = 0.0, [], []
    for m, (ttm, fwd, df, K, T) in enumerate(zip(ttms, forwards, discfactors, strikes_ttms, types_ttms)):
        blk = inputs[m] if inputs is not None else draw_inputs(rng, ttm - t0, nb_path, params["shift_p"], params["mean_p"], params["shift_m"], params["mean_m"])
        x, lp, lm = step_fixed(x, lp, lm, *blk, **{k: v for k, v in params.items() if k not in ("lambda_p", "lambda_m")})
        t0 = ttm
        p, e = _mc.mc_payoffs(x, x, ttm, fwd, K, T, df, 1)
        prices.append(p)
        stds.append(e)
    return prices, stds
