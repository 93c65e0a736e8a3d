"""CPU restatement (numpy) of the reference's rough-LogSV multi-factor Monte Carlo with caller-supplied normals.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of the CUDA path, never imported by the product.
Pinned by tests/golden/rough_mc_*.npz (tests/golden/make_golden.py --only-rough ran the unmodified reference).

Follows, function by function (paths under /root/reference/src/stochvolmodels/pricers):
  drift_ode_solve2                 rough_logsv/split_simulation.py:76-124    classical RK4 on  z' = -x (z - v0) + (k1 + k2 <w,z>)(theta - <w,z>)
  diffus_sde_solve_f64             :231-250                                  exact log-normal step of the weighted sum, shared shift Q
  drift_diffus_strand_f64          :253-284                                  Strang splitting  D(h/2) S(h) D(h/2)
  log_spot_full_solve2_f64         :287-337                                  log-spot increment from the vol increment (trapezoid c1 = c2 = 1/2)
  log_spot_full_combined_f64       :340-361                                  loop over the time grid, h = grid[1] - grid[0]
  rough_logsv_mc_chain_pricer_fixed_randoms   logsv_pricer.py:1164-1232      per maturity: RESTART from t = 0 with the first S_m rows of Z0 / Z1
  get_randoms_for_rough_vol_chain_valuation   logsv_pricer.py:1076-1097
The payoff step is utils/mc_payoffs.py:10-88 applied to arrays of shape (1, nb_path): ``x0.shape[0]`` is then 1, so the route's
"standard errors" are discfactor * nanstd(payoff) WITHOUT the 1/sqrt(nb_path) (reference quirk, reproduced and documented).
"""
from __future__ import annotations

import numpy as np

from . import mc as _mc


def drift_rk4(nodes, v0, theta, kappa1, kappa2, z0, weight, h):
    """split_simulation.py:76-124; all arrays (n, nb_path)"""
    def slope(z):
        zw = np.sum(weight * z, axis=0)
        return -nodes * (z - v0) + (kappa1 + kappa2 * zw) * (theta - zw)
    s1 = slope(z0)
    s2 = slope(z0 + 0.5 * h * s1)
    s3 = slope(z0 + 0.5 * h * s2)
    s4 = slope(z0 + h * s3)
    return z0 + (h / 6.0) * (s1 + 2.0 * s2 + 2.0 * s3 + s4)


def diffusion_exact(y0, weight, volvol, h, z_rand):
    """split_simulation.py:231-250"""
    weight_sum = np.sum(weight, axis=0)
    vv = volvol * weight_sum
    yw = np.sum(weight * y0, axis=0)
    Yh = yw * np.exp(-0.5 * vv ** 2 * h + vv * (z_rand * np.sqrt(h)))
    return y0 + (1.0 / weight_sum * (Yh - yw))[None, :]


def step(nodes, weight, v0, theta, kappa1, kappa2, log_s, v, y, rho, volvol, h, z0, z1):
    """log_spot_full_solve2_f64 (split_simulation.py:287-337): returns (vol_h, y_h, log_spot_h)"""
    D = drift_rk4(nodes, v0, theta, kappa1, kappa2, v, weight, 0.5 * h)
    S = diffusion_exact(D, weight, volvol, h, z0)
    vol_h = drift_rk4(nodes, v0, theta, kappa1, kappa2, S, weight, 0.5 * h)
    w_vol_h = np.sum(weight * vol_h, axis=0)
    bad = np.isnan(w_vol_h) | (w_vol_h <= 0.0)
    vol_h[:, bad] = 1e-6
    wlam = weight * nodes
    vw = np.sum(weight * v, axis=0)
    volw_h = np.sum(weight * vol_h, axis=0)
    w_inv = 1.0 / np.sum(weight, axis=0)
    rho_comp = np.sqrt(1.0 - rho * rho)
    sq_vw, sq_vhw = np.square(vw), np.square(volw_h)
    w_lam_vol, w_lam_vol_h, w_lam_v0 = np.sum(wlam * v, axis=0), np.sum(wlam * vol_h, axis=0), np.sum(wlam * v0, axis=0)
    term1 = 1.0 / volvol * (((volw_h - vw) / h + 0.5 * w_lam_vol + 0.5 * w_lam_vol_h - w_lam_v0) * w_inv
                            - kappa1 * theta + (kappa1 - kappa2 * theta) * (0.5 * vw + 0.5 * volw_h)
                            + kappa2 * (0.5 * sq_vw + 0.5 * sq_vhw)) * h
    term2 = 0.5 * h * sq_vw + 0.5 * h * sq_vhw
    log_spot_h = log_s - 0.5 * term2 + rho * term1 + rho_comp * np.sqrt(term2) * z1
    y_h = y + 0.5 * h * (vw * vw + volw_h * volw_h)
    return vol_h, y_h, log_spot_h


def log_spot_full_combined(nodes, weight, v0, theta, kappa1, kappa2, log_s0, v_init, rho, volvol, timegrid, Z0, Z1):
    """(log_spot (1, P), vol (n, P), qv (1, P)) after the whole grid (split_simulation.py:340-361)"""
    h = timegrid[1] - timegrid[0]
    P = Z0.shape[1]
    vol, y, ls = v_init.copy(), np.zeros((1, P)), np.ones((1, P)) * log_s0
    for idx in range(timegrid.size - 1):
        vol, y, ls = step(nodes, weight, v0, theta, kappa1, kappa2, ls, vol, y, rho, volvol, h, Z0[idx], Z1[idx])
    return ls, vol, y


def rough_randoms(ttms, nb_path, nb_steps_per_year, seed):
    """logsv_pricer.py:1076-1097: Z0 then Z1 of shape (S_last, nb_path) from a LOCAL RandomState + the per-maturity grids"""
    rng = np.random.RandomState(seed)
    grids, S = [], 0
    for ttm in ttms:
        S, _ = _mc.set_time_grid(ttm, nb_steps_per_year)
        grids.append(np.linspace(0.0, ttm, S + 1))          # utils/funcs.py:45-46
    Z0 = rng.normal(0, 1, size=(S, nb_path))
    Z1 = rng.normal(0, 1, size=(S, nb_path))
    return Z0, Z1, grids


def rough_chain_fixed(ttms, forwards, discfactors, strikes_ttms, types_ttms, Z0, Z1, sigma0, theta, kappa1, kappa2, beta, orthog_vol, weights, nodes,
                      timegrids, variable_type: int = 1, return_states: bool = False):
    """rough_logsv_mc_chain_pricer_fixed_randoms (logsv_pricer.py:1164-1232)"""
    weights, nodes = np.asarray(weights, dtype=float), np.asarray(nodes, dtype=float)
    n, P = nodes.size, Z0.shape[1]
    volvol = np.sqrt(beta ** 2 + orthog_vol ** 2)
    rho = beta / volvol
    v0_vec = np.repeat(np.full((n,), sigma0 / np.sum(weights))[:, None], P, axis=1)
    w_vec, n_vec = np.repeat(weights[:, None], P, axis=1), np.repeat(nodes[:, None], P, axis=1)
    prices, stds, states = [], [], []
    for ttm, fwd, df, K, T, grid in zip(ttms, forwards, discfactors, strikes_ttms, types_ttms, timegrids):
        S = grid.size - 1
        ls, vol, qv = log_spot_full_combined(n_vec, w_vec, v0_vec, theta, kappa1, kappa2, 0.0, v0_vec.copy(), rho, volvol, grid, Z0[:S], Z1[:S])
        p, e = _mc.mc_payoffs(ls[0], qv[0], ttm, fwd, K, T, df, variable_type)
        prices.append(p)
        stds.append(e * np.sqrt(P))          # x0.shape[0] == 1 in the reference's 2-d call: no 1/sqrt(nb_path)
        states.append((ls, vol, qv))
    return (prices, stds, states) if return_states else (prices, stds)
