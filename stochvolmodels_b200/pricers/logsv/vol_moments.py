"""Moments of the mean-adjusted volatility Y = sigma - theta and the model variance-swap rate -- host-side 4x4 linear algebra used by the
``PARAMS_WITH_VARSWAP_FIT`` calibration mode (reference pricers/logsv/vol_moments_ode.py:29-217, LogSvParams.get_vol_moments_lambda
logsv_params.py:269-323).  Truncated moment system  dM/dtau = Lambda M + C  (Eq. (3.48) of Sepp & Rakhmonov), closed by freezing
the (k*+1)-th moment at its initial value (Eq. (3.51)):

    M(tau)            = E(tau) M0 + R(tau) C,                     E = expm(Lambda tau),  R = Lambda^-1 (E - I)
    int_0^tau M dt    = R(tau) M0 + Lambda^-1 (R(tau) - tau I) C
    expected qvar/tau = (m2_int + 2 theta m1_int) / tau + theta^2                                   (Eq. (3.53))

Nothing here touches the GPU: it runs once per optimizer point on k* = 4 unknowns.
"""
from __future__ import annotations

import numpy as np


def vol_moments_generator(params, n_terms: int = 4) -> np.ndarray:
    """Lambda^(1,k*): row n (1-based) holds c(n) theta^2, 2 c(n) theta, c(n) - n kappa, -n kappa2 on the columns n-2 .. n+1, with
    c(n) = vartheta^2 n (n-1) / 2 and kappa = kappa1 + kappa2 theta."""
    kappa, kappa2, theta, v2 = params.kappa1 + params.kappa2 * params.theta, params.kappa2, params.theta, params.vartheta2
    lam = np.zeros((n_terms, n_terms))
    for row in range(n_terms):
        n = row + 1
        c_n = 0.5 * v2 * n * (n - 1.0)
        if row >= 2:
            lam[row, row - 2] = c_n * theta * theta
        if row >= 1:
            lam[row, row - 1] = 2.0 * c_n * theta
        lam[row, row] = c_n - n * kappa
        if row + 1 < n_terms:
            lam[row, row + 1] = -n * kappa2
    return lam


def vol_moments(params, t: float = 1.0, n_terms: int = 4, integrated: bool = False) -> np.ndarray:
    """(E[Y^1], ..., E[Y^k*]) at ``t``, or their time integrals over [0, t] (``integrated``); at t = 0 the initial powers of Y0."""
    from scipy.linalg import expm
    y = params.sigma0 - params.theta
    m0 = y ** np.arange(1, n_terms + 1)
    if np.isclose(abs(t), 0.0):
        return m0
    free = np.zeros(n_terms)
    free[1] = params.vartheta2 * params.theta * params.theta
    free[-1] = -n_terms * params.kappa2 * y ** (n_terms + 1)
    lam = vol_moments_generator(params, n_terms)
    lam_inv = np.linalg.inv(lam)
    e_t = expm(lam * t)
    r_t = lam_inv @ (e_t - np.eye(n_terms))
    if integrated:
        return r_t @ m0 + (lam_inv @ (r_t - t * np.eye(n_terms))) @ free
    return e_t @ m0 + r_t @ free


def expected_qvar(params, ttm: float = 1.0, n_terms: int = 4) -> float:
    """annualised expected quadratic variance = fair variance of a continuously monitored variance swap."""
    if np.isclose(ttm, 0.0):
        return float(np.square(params.sigma0))
    m = vol_moments(params, ttm, n_terms, integrated=True)
    return float((m[1] + 2.0 * params.theta * m[0]) / ttm + params.theta * params.theta)


def fit_vol_backbone_to_varswaps(params, varswap_strikes, n_terms: int = 4):
    """eta per quoted maturity such that the model's forward variance matches the market's between consecutive maturities
    (``varswap_strikes``: pandas Series of var-swap VOLS indexed by ttm).  Non-positive ratios fall back to 1; maturities under 0.06y
    get the square root of the ratio (the reference's ad-hoc short-end damping, vol_moments_ode.py:207-210).  Returns a Series."""
    import pandas as pd
    ttms = np.asarray(varswap_strikes.index, dtype=float)
    market = ttms * np.square(np.asarray(varswap_strikes, dtype=float))
    model = np.array([expected_qvar(params, t, n_terms) for t in ttms]) * ttms
    d_market, d_model = np.diff(market, prepend=0.0), np.diff(model, prepend=0.0)
    eta = d_market / d_model
    eta = np.where(eta > 0.0, eta, 1.0)
    eta = np.where(ttms < 0.06, np.sqrt(eta), eta)
    return pd.Series(eta, index=ttms)


# ---- the reference's entry-point names and signatures (pricers/logsv/vol_moments_ode.py:29-217) over the functions above -------------
def compute_analytic_vol_moments(params, t: float = 1.0, n_terms: int = 4, is_qvar: bool = False) -> np.ndarray:
    """moments of Y at ``t`` (``is_qvar``: their integrals over [0, t]) -- vol_moments_ode.py:29"""
    return vol_moments(params, t=t, n_terms=n_terms, integrated=is_qvar)


def compute_analytic_qvar(params, ttm: float = 1.0, n_terms: int = 4) -> float:
    """annualised expected quadratic variance, Eq. (3.53) -- vol_moments_ode.py:110"""
    return expected_qvar(params, ttm=ttm, n_terms=n_terms)


def compute_vol_moments_t(params, ttm: np.ndarray, n_terms: int = 4, is_print: bool = False) -> np.ndarray:
    """[len(ttm), n_terms] table of the moments (vol_moments_ode.py:149)"""
    table = np.array([vol_moments(params, t=float(t_), n_terms=n_terms) for t_ in ttm]).reshape(len(ttm), n_terms)
    if is_print:
        for t_, row in zip(ttm, table):
            print(f"t={t_}: {row}")
    return table


def compute_expected_vol_t(params, t: np.ndarray, n_terms: int = 4) -> np.ndarray:
    """E[sigma_t] = E[Y_t] + theta over an array of times (vol_moments_ode.py:164)"""
    return np.array([vol_moments(params, t=float(t_), n_terms=n_terms)[0] + params.theta for t_ in t])


def compute_sqrt_qvar_t(params, t: np.ndarray, n_terms: int = 4) -> np.ndarray:
    """model variance-swap rate in vol terms over an array of maturities (vol_moments_ode.py:178)"""
    return np.sqrt(np.array([expected_qvar(params, ttm=float(t_), n_terms=n_terms) for t_ in t]))


def fit_model_vol_backbone_to_varswaps(log_sv_params, varswap_strikes, n_terms: int = 4, verbose: bool = False):
    """vol_moments_ode.py:186; ``verbose`` prints the market / model forward variances next to the fitted eta"""
    eta = fit_vol_backbone_to_varswaps(log_sv_params, varswap_strikes, n_terms=n_terms)
    if verbose:
        ttms = np.asarray(varswap_strikes.index, dtype=float)
        print("vars_swaps")
        for t_, k, e in zip(ttms, np.asarray(varswap_strikes, dtype=float), np.asarray(eta)):
            print(f"  ttm={t_:.4f} strike={k:.4f} market_qvar_dt={t_ * k * k:.6f} model_qvar_dt={expected_qvar(log_sv_params, t_, n_terms) * t_:.6f} "
                  f"model_eta={e:.4f}")
    return eta
