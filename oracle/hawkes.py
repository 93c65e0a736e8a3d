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


# ---------------------------------------------------------------------------------------------------------------------------------
# Fourier route (hawkes_jd_pricer.py:365-641), pinned by tests/golden/hawkes_fourier.npz (make_golden.py --only-hawkes-fourier)
#   solve_ode_for_a            :582-641   Riccati system for (a0, a_p, a_m) per transform point, SciPy RK45 with default tolerances
#   compute_hawkes_a_mgf_grid  :518-547   log-MGF = a0 + a_p lambda_p + a_m lambda_m
#   hawkesjd_chain_pricer      :365-417   500-point grid, vol_scaler = clip(sigma, 0.2, 0.5) sqrt(min(ttm_min, 1/12)), A carried over maturities
#   ..._with_risk_premia       :420-515   grid on Re = -1/2 - gamma, normalisers / forwards under the risk kernel from two single-point solves
# ---------------------------------------------------------------------------------------------------------------------------------
MAX_PHI = 500                 # :37


def fourier_vol_scaler(sigma: float, ttm_min: float) -> float:
    return float(np.clip(sigma, 0.2, 0.5) * np.sqrt(min(ttm_min, 1.0 / 12.0)))          # :360-362


class HawkesRhs:
    """right-hand side of :607-626 for rows of A (one row per transform point)"""

    def __init__(self, p: dict, phi: np.ndarray, psi: np.ndarray):
        self.p, self.phi, self.psi = p, np.asarray(phi, dtype=np.complex128), np.asarray(psi, dtype=np.complex128)
        self.comp_p = np.exp(p["shift_p"]) / (1.0 - p["mean_p"]) - 1.0              # HawkesJDParams.__post_init__ :67-70
        self.comp_m = np.exp(p["shift_m"]) / (1.0 - p["mean_m"]) - 1.0

    def __call__(self, A: np.ndarray, rows=None) -> np.ndarray:
        """``rows``: the grid points the rows of A belong to (mgf.rk45_grid passes the active subset)"""
        p = self.p
        phi, psi = (self.phi, self.psi) if rows is None else (self.phi[rows], self.psi[rows])
        zp = phi - p["beta1_p"] * A[:, 1] - p["beta1_m"] * A[:, 2]
        zm = phi - p["beta2_p"] * A[:, 1] - p["beta2_m"] * A[:, 2]
        j_p = np.exp(-p["shift_p"] * zp) / (1.0 + p["mean_p"] * zp) - 1.0
        j_m = np.exp(-p["shift_m"] * zm) / (1.0 + p["mean_m"] * zm) - 1.0
        out = np.empty_like(A)
        out[:, 0] = p["kappa_p"] * p["theta_p"] * A[:, 1] + p["kappa_m"] * p["theta_m"] * A[:, 2] + np.square(p["sigma"]) * (0.5 * (phi + 1.0) * phi - psi)
        out[:, 1] = j_p - p["kappa_p"] * A[:, 1] + self.comp_p * phi
        out[:, 2] = j_m - p["kappa_m"] * A[:, 2] + self.comp_m * phi
        return out


def a_mgf_grid(dtau: float, phi: np.ndarray, p: dict, a_t0: np.ndarray = None, psi: np.ndarray = None):
    """(a_t1 [P, 3], log_mgf [P]) of compute_hawkes_a_mgf_grid"""
    from . import mgf as _mgf
    phi = np.asarray(phi, dtype=np.complex128)
    psi = np.zeros_like(phi) if psi is None else psi
    a_t0 = np.zeros((phi.shape[0], 3), dtype=np.complex128) if a_t0 is None else a_t0
    a_t1 = _mgf.rk45_grid(HawkesRhs(p, phi, psi), a_t0, float(dtau))
    return a_t1, a_t1[:, 0] + a_t1[:, 1] * p["lambda_p"] + a_t1[:, 2] * p["lambda_m"]


def forwards_under_risk_kernel(p: dict, gamma: float, ttms, forwards):
    """(normalizers, gamma_forwards) of :487-515: two single-point solves per maturity, each restarted from A = 0"""
    norm, fwd = np.ones(len(ttms)), np.ones(len(ttms))
    for m, (ttm, f) in enumerate(zip(ttms, forwards)):
        lm0 = a_mgf_grid(ttm, np.array([-gamma + 0j]), p)[1]
        lm1 = a_mgf_grid(ttm, np.array([-gamma - 1.0 + 0j]), p)[1]
        norm[m] = 1.0 / np.exp(lm0[0].real)
        fwd[m] = f * np.exp(lm1[0].real) * norm[m]
    return norm, fwd


def gamma_slice_prices(log_mgf, phi, gamma, forward, normalizer, gamma_forward, strikes, types):
    """slice_pricer_with_mgf_grid_with_gamma (utils/mgf_pricer.py:273-320).  Its fast branch tests Re(phi) == +(1/2 + gamma) while the grid
    sits on -(1/2 + gamma), so the general weights are what runs; the discount factor is not applied (both reproduced)."""
    from . import mgf as _mgf
    dp = _mgf.legacy_simpson_weights(phi)
    if np.all(np.abs(np.real(phi) - (0.5 + gamma)) < 1e-10):
        w = (dp / np.pi) / (np.imag(phi) ** 2 + 0.25) + 0j
    else:
        w = -(dp / np.pi) / ((phi + gamma + 1.0) * (phi + gamma))
    out = np.zeros(len(strikes))
    for j, (k, ty) in enumerate(zip(strikes, types)):
        capped = np.nansum(np.real(w * np.exp(-np.log(forward / k) * phi + log_mgf)))
        gk = np.power(k, 1.0 + gamma)
        if str(ty) == "C":
            out[j] = gamma_forward - normalizer * gk * capped
        elif str(ty) == "P":
            out[j] = k - normalizer * gk * capped
        else:
            raise ValueError("not implemented")
    return out


def fourier_chain_prices(p: dict, ttms, forwards, discfactors, strikes_ttms, types_ttms, is_spot_measure=True, vol_scaler=None, gamma=None,
                         return_grids=False):
    """hawkesjd_chain_pricer (gamma None) / hawkesjd_chain_pricer_with_risk_premia"""
    from . import mgf as _mgf
    if vol_scaler is None:
        vol_scaler = fourier_vol_scaler(p["sigma"], float(np.min(ttms)))
    phi = _mgf.phi_grid(vol_scaler, True, MAX_PHI)
    if gamma is not None:
        phi = (-0.5 - gamma) + 1j * np.imag(phi)
        norm, gfw = forwards_under_risk_kernel(p, gamma, ttms, forwards)
    a, t0, prices, grids = np.zeros((phi.shape[0], 3), dtype=np.complex128), 0.0, [], []
    for m, ttm in enumerate(ttms):
        a, lm = a_mgf_grid(ttm - t0, phi, p, a)
        t0 = ttm
        if gamma is None:
            prices.append(_mgf.vanilla_slice_prices(lm, phi, forwards[m], strikes_ttms[m], types_ttms[m], discfactors[m], is_spot_measure))
        else:
            prices.append(gamma_slice_prices(lm, phi, gamma, forwards[m], norm[m], gfw[m], strikes_ttms[m], types_ttms[m]))
        grids.append((a.copy(), lm.copy()))
    return (prices, grids, phi) if return_grids else prices
