"""Hawkes jump-diffusion Monte Carlo on the GPU against the reference goldens and the numpy oracle.

tests/golden/hawkes_mc.npz = outputs of the UNMODIFIED reference (pricers/hawkes_jd_pricer.py:644-779) after np.random.seed(seed); the
oracle re-draws the same legacy-generator blocks, the strict GPU kernel consumes them and must reproduce the reference operation by
operation; the fused kernel (in-kernel Philox draws) is checked against the oracle on the draws it exports."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import hawkes

pytestmark = pytest.mark.gpu


def test_strict_kernel_vs_reference_golden_terminal_states(cuda_lib):
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import simulate_hawkesjd_terminal
    g = load_golden("hawkes_mc.npz")
    for name in ("dflt", "drift"):
        params = dict(zip(hawkes.KEYS, g[f"{name}_params"]))
        N, ttm = int(g[f"{name}_N"]), float(g[f"{name}_ttm"])
        blk = hawkes.draw_inputs(np.random.RandomState(int(g[f"{name}_seed"])), ttm, N, params["shift_p"], params["mean_p"], params["shift_m"], params["mean_m"])
        kw = {k: v for k, v in params.items() if k not in ("lambda_p", "lambda_m")}
        x, lp, lm = simulate_hawkesjd_terminal(ttm=ttm, x0=g[f"{name}_x0"], lambda_p0=g[f"{name}_lp0"], lambda_m0=g[f"{name}_lm0"], nb_path=N,
                                               inputs=blk[:5], **kw)
        np.testing.assert_allclose(x, g[f"{name}_x"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(lp, g[f"{name}_lp"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lm, g[f"{name}_lm"], rtol=1e-12, atol=1e-12)
        xo, lpo, lmo = hawkes.step_fixed(g[f"{name}_x0"], g[f"{name}_lp0"], g[f"{name}_lm0"], *blk, **params)
        np.testing.assert_array_equal(x, xo)              # IEEE operation by operation: bit-identical to the numpy restatement
        np.testing.assert_array_equal(lp, lpo)
        np.testing.assert_array_equal(lm, lmo)
        assert np.any(lp != g[f"{name}_lp0"]) and np.max(np.abs(x)) > 0.05           # jumps did fire


@pytest.mark.parametrize("gauss", ["fp32", "fp64"])
def test_fused_kernel_vs_oracle_on_its_own_draws(cuda_lib, gauss):
    """chain prices of the in-kernel-draw kernel == oracle chain on the exported draws (two slices: the second continues from the first's
    state on Philox sub-stream 1); terminal states path by path"""
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, OptionChain
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import hawkesjd_device_draws, simulate_hawkesjd_terminal
    p = HawkesJDParams(mu=0.02, sigma=0.4, lambda_p=9.0, lambda_m=11.0)
    d = {k: v for k, v in p.to_dict().items() if k in hawkes.KEYS}
    N, seed = 6000, 123
    ttms, fw, df = np.array([0.04, 0.1]), np.array([1.0, 1.01]), np.array([0.999, 0.99])
    K = np.array([0.85, 0.95, 1.0, 1.05, 1.15]); T = np.array(["P", "P", "C", "C", "IC"])
    chain = OptionChain(ttms=ttms, forwards=fw, discfactors=df, strikes_ttms=[K, K * 1.01], optiontypes_ttms=[T, T])
    pg, eg = HawkesJDPricer().model_mc_price_chain(chain, p, nb_path=N, seed=seed, gauss=gauss)
    inputs, t0 = [], 0.0
    for m, ttm in enumerate(ttms):
        inputs.append(hawkesjd_device_draws(seed, 0, N, m, ttm - t0, gauss=gauss, **d))
        t0 = ttm
    po, eo = hawkes.chain_prices(d, ttms, fw, df, chain.strikes_ttms, chain.optiontypes_ttms, N, inputs=inputs)
    for m in range(2):
        np.testing.assert_allclose(pg[m], po[m], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(eg[m], eo[m], rtol=1e-10, atol=1e-15)
    # terminal values API: constant start, then per-path continuation
    x, lp, lm = HawkesJDPricer().simulate_terminal_values(p, ttm=0.04, nb_path=N, seed=seed, gauss=gauss)
    xo, lpo, lmo = hawkes.step_fixed(np.zeros(N), p.lambda_p * np.ones(N), p.lambda_m * np.ones(N), *inputs[0], **d)
    tol = dict(rtol=1e-12, atol=1e-14)          # throughput stepper: contracted arithmetic, same draws and jump decisions as the strict one
    np.testing.assert_allclose(x, xo, **tol)
    np.testing.assert_allclose(lp, lpo, **tol)
    kw = {k: v for k, v in d.items() if k not in ("lambda_p", "lambda_m")}
    x2, lp2, lm2 = simulate_hawkesjd_terminal(ttm=ttms[1] - ttms[0], x0=x, lambda_p0=lp, lambda_m0=lm, nb_path=N, seed=seed, gauss=gauss, slice_index=1, **kw)
    xo2, lpo2, lmo2 = hawkes.step_fixed(xo, lpo, lmo, *inputs[1], **d)
    np.testing.assert_allclose(x2, xo2, **tol)
    np.testing.assert_allclose(lm2, lmo2, **tol)
    # the draws themselves: uniform clocks, exponential sizes, unit-variance increments
    W0, U_P, U_M, J_P, J_M, dt = inputs[1]
    n = W0.size
    assert abs(W0.mean()) < 4 * np.sqrt(dt / n) and abs(W0.var() / dt - 1) < 4 * np.sqrt(2 / n)
    assert abs((U_P * dt).mean() - 1) < 4 / np.sqrt(n) and abs((U_M * dt).mean() - 1) < 4 / np.sqrt(n)
    assert abs((J_P - p.shift_p).mean() / p.mean_p - 1) < 4 / np.sqrt(n) and np.all(J_P >= p.shift_p) and np.all(J_M <= p.shift_m)


def test_chain_prices_within_mc_error_of_reference_mc_golden(cuda_lib):
    """statistical agreement with the reference's own MC sample (4000 paths, golden) using 2e6 GPU paths: |diff| < 4 combined SE"""
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, OptionChain
    g = load_golden("hawkes_mc.npz")
    params = HawkesJDParams(**dict(zip(hawkes.KEYS, g["chain_params"])))
    M = g["chain_ttms"].shape[0]
    chain = OptionChain(ttms=g["chain_ttms"], forwards=g["chain_forwards"], discfactors=g["chain_discfactors"],
                        strikes_ttms=[g["chain_strikes"] * f for f in g["chain_forwards"]], optiontypes_ttms=[g["chain_types"]] * M)
    p, e = HawkesJDPricer().model_mc_price_chain(chain, params, nb_path=2_000_000, seed=5)
    for m in range(M):
        z = (p[m] - g["chain_prices"][m]) / np.sqrt(e[m] ** 2 + g["chain_stds"][m] ** 2)
        assert np.all(np.abs(z) < 4.0), z
    # martingale check: E[exp(x_T)] = exp(mu T) under the compensated dynamics
    x, _, _ = HawkesJDPricer().simulate_terminal_values(params, ttm=0.1, nb_path=2_000_000, seed=6)
    m1 = np.exp(x)
    assert abs(m1.mean() - np.exp(params.mu * 0.1)) < 4 * m1.std() / np.sqrt(x.size)
    # and the Fourier route prices the same chain within the MC error
    four = HawkesJDPricer().price_chain(chain, params)
    for m in range(M):
        assert np.all(np.abs(p[m] - four[m]) < 4.5 * e[m] + 2e-4), (p[m], four[m], e[m])


def test_sharded_driver_equals_the_host_level_chain(cuda_lib):
    """multi_gpu.mc_chain_distributed('hawkes', ...) (device-level b200sv_dev_hawkesjd_slice + the LogSV chain's payoff / finalize calls) on
    one rank == b200sv_hawkesjd_mc_chain; two half-shards run back to back on this GPU hold the same terminal states as the unsharded run
    (global path ids), which is what makes the prices independent of the number of GPUs."""
    import torch
    from ctypes import byref, c_void_p
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, engine, get_btc_test_chain_data
    from stochvolmodels_b200 import _capi as C
    from stochvolmodels_b200.multi_gpu import CudaMcEngine, mc_chain_distributed
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import STEPS_PER_YEAR, _params_c
    chain = get_btc_test_chain_data()
    params = HawkesJDParams()
    d = params.to_dict()
    d.pop("risk_premia_gamma", None)
    pc = _params_c(**d)
    N, seed = 200_003, 17
    flags = engine.mc_flags("fp64", "fp32")
    p_h, e_h = HawkesJDPricer().model_mc_price_chain(chain, params, nb_path=N, seed=seed)
    p_d, e_d = mc_chain_distributed("hawkes", pc, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms, chain.optiontypes_ttms,
                                    N, STEPS_PER_YEAR, True, C.LOG_RETURN, seed, flags)
    for m in range(len(p_h)):
        np.testing.assert_allclose(p_d[m], p_h[m], rtol=1e-12)
        np.testing.assert_allclose(e_d[m], e_h[m], rtol=1e-10)
    # two shards vs one engine over all paths: identical terminal states
    from stochvolmodels_b200.utils.funcs import set_time_grid
    S, dt, _ = set_time_grid(float(chain.ttms[0]), STEPS_PER_YEAR)
    whole = CudaMcEngine("hawkes", pc, N, 0, flags, 4)
    whole.simulate_slice(0, True, S, dt, 1.0, True, float(chain.forwards[0]), seed)
    n0 = N // 2
    parts = [CudaMcEngine("hawkes", pc, n0, 0, flags, 4), CudaMcEngine("hawkes", pc, N - n0, n0, flags, 4)]
    moments = 0.0
    for e in parts:
        moments = moments + e.simulate_slice(0, True, S, dt, 1.0, True, float(chain.forwards[0]), seed).cpu().numpy()
    torch.cuda.synchronize()
    joined = torch.cat([e.state for e in parts], dim=1)
    assert torch.equal(joined, whole.state)
    np.testing.assert_allclose(moments, whole.moments.cpu().numpy(), rtol=1e-12)


def test_fourier_route_vs_reference_golden_and_oracle(cuda_lib):
    """hawkesjd_chain_pricer / hawkesjd_chain_pricer_with_risk_premia (reference :365-515): per-maturity ODE grids (a0, a_p, a_m), log-MGF and chain
    prices of the CUDA path against the reference's own outputs (hawkes_fourier.npz, two parameter sets, three carried maturities), the
    risk-kernel normalisers / forwards / prices, the pricer API, and the error conventions"""
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, OptionChain
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import (hawkesjd_chain_pricer, hawkesjd_chain_pricer_with_risk_premia,
                                                              hawkesjd_forwards_under_risk_kernel, set_vol_scaler)
    g = load_golden("hawkes_fourier.npz")
    K, T, ttms, fw, df = g["strikes"], g["types"], g["ttms"], g["forwards"], g["discfactors"]
    Ks, Ts = [K * f for f in fw], [T] * 3
    for name in ("dflt", "alt"):
        params = HawkesJDParams(**dict(zip(hawkes.KEYS, g[f"{name}_params"])))
        prices, a, lm, _, _ = hawkesjd_chain_pricer(params, ttms, fw, df, Ks, Ts, return_grids=True)
        for m in range(3):
            np.testing.assert_allclose(a[m], g[f"{name}_a_{m}"], rtol=1e-10, atol=1e-11)
            np.testing.assert_allclose(lm[m], g[f"{name}_lm_{m}"], rtol=1e-10, atol=1e-11)
            np.testing.assert_allclose(prices[m], g[f"{name}_prices"][m], rtol=1e-10, atol=1e-13)
    params = HawkesJDParams(**dict(zip(hawkes.KEYS, g["dflt_params"])), risk_premia_gamma=float(g["gamma"]))
    norm, gfw = hawkesjd_forwards_under_risk_kernel(params, float(g["gamma"]), ttms, fw)
    np.testing.assert_allclose(norm, g["gamma_normalizers"], rtol=1e-11)
    np.testing.assert_allclose(gfw, g["gamma_forwards"], rtol=1e-11)
    gp = hawkesjd_chain_pricer_with_risk_premia(params, ttms, fw, df, Ks, Ts)
    chain = OptionChain(ttms=ttms, forwards=fw, discfactors=df, strikes_ttms=Ks, optiontypes_ttms=Ts)
    api = HawkesJDPricer().price_chain(chain, params)
    for m in range(3):
        np.testing.assert_allclose(gp[m], g["gamma_prices"][m], rtol=1e-10, atol=1e-13)
        np.testing.assert_array_equal(api[m], gp[m])
    assert set_vol_scaler(0.45, 0.05) == np.clip(0.45, 0.2, 0.5) * np.sqrt(0.05)
    # grid entry points: the second maturity from the first one's A on a sub-grid; a psi grid against the oracle
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import compute_hawkes_a_mgf_grid, solve_a_ode_grid
    pa = HawkesJDParams(**dict(zip(hawkes.KEYS, g["alt_params"])))
    sub = slice(0, 500, 7)
    a1, lm1 = compute_hawkes_a_mgf_grid(ttms[1] - ttms[0], g["alt_phi"][sub], pa, a_t0=g["alt_a_0"][sub])
    np.testing.assert_allclose(a1, g["alt_a_1"][sub], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(lm1, g["alt_lm_1"][sub], rtol=1e-10, atol=1e-11)
    psi = -0.3 + 1j * np.linspace(0.0, 40.0, 33)
    phi0 = np.zeros_like(psi)
    ao, _ = hawkes.a_mgf_grid(0.2, phi0, dict(zip(hawkes.KEYS, g["alt_params"])), psi=psi)
    np.testing.assert_allclose(solve_a_ode_grid(phi0, 0.2, pa, psi_grid=psi), ao, rtol=1e-10, atol=1e-11)
    # a grid the goldens do not cover: the oracle (SciPy-RK45 clone on the same Riccati system), inverse-measure payoffs
    po = hawkes.fourier_chain_prices(dict(zip(hawkes.KEYS, g["alt_params"])), ttms[:2], fw[:2], df[:2], Ks[:2], [np.array(["IP", "P", "C", "IC", "IC"])] * 2,
                                     is_spot_measure=False, vol_scaler=0.2)
    pg = hawkesjd_chain_pricer(HawkesJDParams(**dict(zip(hawkes.KEYS, g["alt_params"]))), ttms[:2], fw[:2], df[:2], Ks[:2],
                               [np.array(["IP", "P", "C", "IC", "IC"])] * 2, is_spot_measure=False, vol_scaler=0.2)
    for m in range(2):
        np.testing.assert_allclose(pg[m], po[m], rtol=1e-10, atol=1e-13)
    with pytest.raises(ValueError, match="not implemented"):
        hawkesjd_chain_pricer_with_risk_premia(params, ttms, fw, df, Ks, [np.array(["IP", "P", "C", "C", "C"])] * 3)
    with pytest.raises(NotImplementedError):
        hawkesjd_chain_pricer(params, ttms, fw, df, Ks, Ts, is_stiff_solver=True)


def test_hawkes_calibration_drivers_recover_a_synthetic_market(cuda_lib):
    """calibrate_model_params_to_chain / calibrate_risk_premia_gamma_to_chain (reference :230-357) around the GPU Fourier pricer: a market
    generated by the model itself is re-fitted from a perturbed start (objective -> ~0, implied vols reproduced)"""
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, OptionChain
    pricer = HawkesJDPricer()
    K = np.linspace(0.75, 1.3, 12)
    ttms, fw = np.array([0.08, 0.25, 0.5]), np.array([1.0, 1.0, 1.0])
    mk = lambda vols: OptionChain(ttms=ttms, forwards=fw, strikes_ttms=[K] * 3, optiontypes_ttms=[np.where(K >= 1.0, "C", "P")] * 3,
                                  ids=np.array(["1m", "3m", "6m"]), bid_ivs=vols, ask_ivs=[v.copy() for v in vols])
    flat = mk([0.5 * np.ones_like(K)] * 3)
    # the reference's parametrisation: shared kappa, +-beta pairs
    true = HawkesJDParams(sigma=0.4, mean_p=0.04, mean_m=-0.05, theta_p=6.0, theta_m=8.0, kappa_p=25.0, kappa_m=25.0, beta1_p=30.0, beta2_p=-30.0,
                          beta1_m=40.0, beta2_m=-40.0)
    market = pricer.compute_model_ivols_for_chain(flat, true)
    assert all(np.all(np.isfinite(v)) for v in market)
    chain = mk(market)
    start = HawkesJDParams(sigma=0.45, mean_p=0.03, mean_m=-0.04, theta_p=7.0, theta_m=7.0, kappa_p=22.0, kappa_m=28.0, beta1_p=33.0, beta2_p=-33.0,
                           beta1_m=36.0, beta2_m=-36.0)
    fit = pricer.calibrate_model_params_to_chain(chain, start, maxiter=60)
    refit = pricer.compute_model_ivols_for_chain(chain, fit)
    before = pricer.compute_model_ivols_for_chain(chain, start)
    err = lambda v: max(np.max(np.abs(a - b)) for a, b in zip(v, market))
    assert err(refit) < 5e-3 and err(refit) < 0.2 * err(before), (err(refit), err(before))     # 8 weakly identified parameters, ftol 1e-8
    assert fit.kappa_p == fit.kappa_m and fit.beta2_p == -fit.beta1_p and fit.beta2_m == -fit.beta1_m and fit.shift_p == start.shift_p
    assert fit.jump1_cond + fit.jump2_cond >= -1e-8
    # risk-premium kernel: market generated with gamma = 0.5 and sigma = 0.42, start from (0.45, 0.2)
    true_g = HawkesJDParams(sigma=0.42, risk_premia_gamma=0.5)
    chain_g = mk(pricer.compute_model_ivols_for_chain(flat, true_g))
    fit_g = pricer.calibrate_risk_premia_gamma_to_chain(chain_g, HawkesJDParams(sigma=0.45, risk_premia_gamma=0.2), print_iter=False)
    vols_g = lambda p_: pricer.compute_model_ivols_for_chain(chain_g, p_)
    err_g = lambda v: max(np.max(np.abs(a - b)) for a, b in zip(v, chain_g.get_mid_vols()))
    e_fit, e_start = err_g(vols_g(fit_g)), err_g(vols_g(HawkesJDParams(sigma=0.45, risk_premia_gamma=0.2)))
    # sigma and gamma trade off along a shallow valley and the reference's finite-difference step is 0.025: the fit is judged on the vols
    assert e_fit < 0.5 * e_start and 0.3 < fit_g.risk_premia_gamma < 0.7 and abs(fit_g.sigma - 0.42) < 0.01, (e_fit, e_start, fit_g)
