"""Rough-LogSV multi-factor Monte Carlo on the GPU against the reference goldens and the numpy oracle.

goldens: tests/golden/rough_mc_*.npz = outputs of the UNMODIFIED reference ``rough_logsv_mc_chain_pricer_fixed_randoms``
(pricers/logsv_pricer.py:1164-1232) and ``log_spot_full_combined`` (rough_logsv/split_simulation.py:466) on RandomState normals,
n = 3 / 2 / 1 factors (H = 0.30 / 0.45 / 0.50)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import mc, rough

pytestmark = pytest.mark.gpu


def _case(name):
    g = load_golden(f"rough_mc_{name}.npz")
    M = int(g["nslices"])
    return g, M, [g[f"strikes_{m}"] for m in range(M)], [g[f"types_{m}"] for m in range(M)]


@pytest.mark.parametrize("name", ["h030_n3", "h045_n2", "h050_n1"])
def test_rough_chain_fixed_randoms_vs_reference_golden(cuda_lib, name):
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
    g, M, strikes, types = _case(name)
    sigma0, theta, kappa1, kappa2, beta, volvol, H = g["params"]
    Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(g["ttms"], nb_path=int(g["nb_path"]), nb_steps_per_year=int(g["npy"]), seed=int(g["seed"]))
    np.testing.assert_array_equal(Z0[:3, :5], g["Z0_head"])            # same draw order as the reference (Z0 block, then Z1 block)
    for m in range(M):
        np.testing.assert_array_equal(grids[m], g[f"grid_{m}"])
    prices, stds, states = rough_logsv_mc_chain_pricer_fixed_randoms(ttms=g["ttms"], forwards=g["forwards"], discfactors=g["discfactors"],
                                                                     strikes_ttms=strikes, optiontypes_ttms=types, Z0=Z0, Z1=Z1, sigma0=sigma0,
                                                                     theta=theta, kappa1=kappa1, kappa2=kappa2, beta=beta, orthog_vol=volvol,
                                                                     weights=g["weights"], nodes=g["nodes"], timegrids=grids, return_states=True)
    n = g["nodes"].size
    for m in range(M):
        np.testing.assert_allclose(states[m, 0], g[f"log_spot_{m}"][0], rtol=0, atol=5e-12)      # reference kernels are fastmath=True
        np.testing.assert_allclose(states[m, 1:1 + n], g[f"vol_{m}"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(states[m, 1 + n], g[f"qv_{m}"][0], rtol=1e-11, atol=0)
        np.testing.assert_allclose(prices[m], g[f"prices_{m}"], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(stds[m], g[f"stds_{m}"], rtol=1e-9, atol=1e-14)              # NOT divided by sqrt(nb_path): reference quirk


def test_rough_chain_vs_the_references_own_regression_fixture(cuda_lib):
    """the known-answer vector the reference's own test suite holds for this path (tests/test_rough_logsv_pricer_regression.py:
    BTC chain, H = 0.1 -> 3 factors, 10000 paths, seed 10), at the reference's own tolerance rtol 1e-7"""
    from stochvolmodels_b200 import get_btc_test_chain_data
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
    g = load_golden("rough_mc_reference_regression.npz")
    chain = get_btc_test_chain_data()
    sigma0, theta, kappa1, kappa2, beta, volvol, H = g["params"]
    Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(ttms=chain.ttms, nb_path=int(g["nb_path"]), nb_steps_per_year=int(g["npy"]), seed=int(g["seed"]))
    prices, _ = rough_logsv_mc_chain_pricer_fixed_randoms(ttms=chain.ttms, forwards=chain.forwards, discfactors=chain.discfactors,
                                                          strikes_ttms=chain.strikes_ttms, optiontypes_ttms=chain.optiontypes_ttms, Z0=Z0, Z1=Z1,
                                                          sigma0=sigma0, theta=theta, kappa1=kappa1, kappa2=kappa2, beta=beta, orthog_vol=volvol,
                                                          weights=g["weights"], nodes=g["nodes"], timegrids=grids)
    for m in range(int(g["nslices"])):
        np.testing.assert_allclose(prices[m], g[f"expected_prices_{m}"], rtol=1e-7, atol=0)


def test_rough_fixed_random_pricer_is_finite_and_deterministic(cuda_lib):
    """port of the reference's tests/test_rough_logsv_characterization.py::test_rough_fixed_random_pricer_is_finite_and_deterministic
    (H = 0.1 -> three factors; the nodes / weights are the reference optimiser's output for this fixture, stored in the regression golden)"""
    from stochvolmodels_b200 import OptionChain
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
    g = load_golden("rough_mc_reference_regression.npz")
    chain = OptionChain.slice_to_chain(ttm=0.05, forward=1.0, strikes=np.array([0.95, 1.0, 1.05]), optiontypes=np.array(["P", "C", "C"]), id="rough")
    z0, z1, grids = get_randoms_for_rough_vol_chain_valuation(chain.ttms, nb_path=128, nb_steps_per_year=100, seed=123)

    def price():
        return rough_logsv_mc_chain_pricer_fixed_randoms(ttms=chain.ttms, forwards=chain.forwards, discfactors=chain.discfactors,
                                                         strikes_ttms=chain.strikes_ttms, optiontypes_ttms=chain.optiontypes_ttms, Z0=z0, Z1=z1,
                                                         sigma0=0.2, theta=0.2, kappa1=2.0, kappa2=8.0, beta=-0.2, orthog_vol=0.3,
                                                         weights=g["weights"], nodes=g["nodes"], timegrids=grids)
    p1, e1 = price()
    p2, e2 = price()
    assert np.asarray(p1[0]).shape == (3,)
    assert np.all(np.isfinite(p1[0])) and np.all(np.isfinite(e1[0])) and np.all(np.asarray(e1[0]) >= 0.0)
    np.testing.assert_array_equal(p1[0], p2[0])
    np.testing.assert_array_equal(e1[0], e2[0])


def test_rough_qvar_payoffs_and_bad_vol_reset_vs_oracle(cuda_lib):
    """Q_VAR payoffs, and a configuration that drives the weighted vol through zero so that the reference's reset-to-1e-6 branch
    (split_simulation.py:310-312) is exercised: GPU == numpy oracle on the same normals."""
    from stochvolmodels_b200 import VariableType
    from stochvolmodels_b200.pricers.logsv_pricer import rough_logsv_mc_chain_pricer_fixed_randoms
    P, ttms = 4000, np.array([0.25, 0.5])
    Z0, Z1, grids = rough.rough_randoms(ttms, P, 12, 77)      # coarse grid (h = 1/16) + strong quadratic mean reversion + doubled shocks:
    Z0 = Z0 * 2.0                                             # the explicit RK4 drift overshoots through zero on ~2 % of the paths
    w, x = np.array([0.8, 0.35]), np.array([0.0015625, 1.316])
    K = [np.array([0.02, 0.05, 0.1]), np.array([0.05, 0.1, 0.2])]
    T = [np.array(["C", "P", "C"])] * 2
    args = dict(sigma0=0.5, theta=0.6, kappa1=1.0, kappa2=10.0, beta=-0.3, orthog_vol=2.0, weights=w, nodes=x, timegrids=grids)
    po, eo, so = rough.rough_chain_fixed(ttms, np.ones(2), np.ones(2), K, T, Z0, Z1, variable_type=2, return_states=True, **args)
    pg, eg, sg = rough_logsv_mc_chain_pricer_fixed_randoms(ttms=ttms, forwards=np.ones(2), discfactors=np.ones(2), strikes_ttms=K, optiontypes_ttms=T,
                                                           Z0=Z0, Z1=Z1, variable_type=VariableType.Q_VAR, return_states=True, **args)
    nbad = int(np.sum(np.all(so[1][1] == 1e-6, axis=0)))
    assert nbad > 20, "the case no longer reaches the reset branch"
    for m in range(2):
        assert np.all(np.isfinite(so[m][0]))
        np.testing.assert_allclose(sg[m, 0], so[m][0][0], rtol=0, atol=1e-10)
        np.testing.assert_allclose(sg[m, 1:3], so[m][1], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(sg[m, 3], so[m][2][0], rtol=1e-10, atol=0)
        np.testing.assert_allclose(pg[m], po[m], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(eg[m], eo[m], rtol=1e-8, atol=1e-13)


def test_rough_philox_route_and_api(cuda_lib):
    """model_mc_price_chain(use_rough_mc=True): (a) default = the reference's RandomState normals -> equals the fixed-random pricer;
    (b) gauss='fp64' = in-kernel Philox draws -> equals the oracle fed with the exported device normals; (c) H = 1/2 (one node at 1e-3)
    prices agree with the plain LogSV Fourier route within MC error (the rough scheme degenerates to the article's dynamics)."""
    from stochvolmodels_b200 import LogSvParams, LogSVPricer, OptionChain, engine
    from stochvolmodels_b200 import _capi as C
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
    K = np.array([0.8, 0.9, 1.0, 1.1, 1.2]); T = np.array(["P", "P", "C", "C", "C"])
    chain = OptionChain(ttms=np.array([0.1, 0.25]), forwards=np.ones(2), strikes_ttms=[K, K], optiontypes_ttms=[T, T])
    p = LogSvParams(sigma0=0.8, theta=1.0, kappa1=3.0, kappa2=3.0, beta=0.15, volvol=1.8, H=0.3,
                    weights=np.array([0.97702551, 0.75117174, 1.66711921]), nodes=np.array([3.33333333e-02, 6.31209359e+00, 1.01078897e+02]))
    pricer = LogSVPricer()
    N, npy, seed = 20000, 360, 5
    pa, ea = pricer.model_mc_price_chain(chain, p, nb_path=N, nb_steps=npy, use_rough_mc=True, seed=seed)
    Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(chain.ttms, N, npy, seed)
    pb, eb = rough_logsv_mc_chain_pricer_fixed_randoms(ttms=chain.ttms, forwards=chain.forwards, discfactors=chain.discfactors, strikes_ttms=[K, K],
                                                       optiontypes_ttms=[T, T], Z0=Z0, Z1=Z1, sigma0=p.sigma0, theta=p.theta, kappa1=p.kappa1,
                                                       kappa2=p.kappa2, beta=p.beta, orthog_vol=p.volvol, weights=p.weights, nodes=p.nodes, timegrids=grids)
    for m in range(2):
        np.testing.assert_array_equal(pa[m], pb[m])
        np.testing.assert_array_equal(ea[m], eb[m])
    # (b) Philox draws: slice 0 for every maturity, maturity m consumes the first S_m steps
    pc, ec = pricer.model_mc_price_chain(chain, p, nb_path=N, nb_steps=npy, use_rough_mc=True, seed=seed, gauss="fp64")
    S = grids[-1].size - 1
    z0, z1 = engine.device_normals(seed, 0, N, 0, S, C.GAUSS_F64)
    po, eo = rough.rough_chain_fixed(chain.ttms, chain.forwards, chain.discfactors, [K, K], [T, T], z0, z1, p.sigma0, p.theta, p.kappa1, p.kappa2,
                                     p.beta, p.volvol, p.weights, p.nodes, grids)
    for m in range(2):
        np.testing.assert_allclose(pc[m], po[m], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(ec[m], eo[m], rtol=1e-8, atol=1e-13)
    # (c) H = 1/2
    q = LogSvParams(sigma0=1.0, theta=1.0, kappa1=5.0, kappa2=5.0, beta=0.2, volvol=2.0)
    q.approximate_kernel(T=0.25)
    np.testing.assert_array_equal(q.nodes, [1e-3])
    Nh = 400_000
    ph, eh = pricer.model_mc_price_chain(chain, q, nb_path=Nh, nb_steps=npy, use_rough_mc=True, seed=11, gauss="fp32")
    four = pricer.price_chain(chain, q)
    for m in range(2):
        se = eh[m] / np.sqrt(Nh)                      # this route returns the plain std
        assert np.all(np.abs(ph[m] - four[m]) < 4.5 * se + 2e-4), (ph[m], four[m], se)
    with pytest.raises(NotImplementedError):
        LogSvParams(H=0.3).approximate_kernel(T=1.0)


def test_rough_mc_calibration_engine(cuda_lib):
    """CalibrationEngine.ROUGH_MC (reference logsv_pricer.py:266-289, 528-533): market = the model's own rough-MC vols on the fixed
    normals => the objective at the truth is 0 and SLSQP returns to it from a nearby start; the batched objective equals per-set calls."""
    from stochvolmodels_b200 import CalibrationEngine, LogSvParams, LogSVPricer, LogsvModelCalibrationType, OptionChain
    K = np.array([0.85, 0.95, 1.0, 1.05, 1.15]); T = np.array(["P", "P", "C", "C", "C"])
    ttms = np.array([0.1, 0.25])
    w, x = np.array([0.80082171, 0.34941995]), np.array([0.0015625, 1.31611315])
    truth = LogSvParams(sigma0=0.9, theta=1.0, kappa1=4.0, kappa2=4.0, beta=0.3, volvol=1.5, H=0.45, weights=w, nodes=x)
    pricer = LogSVPricer()
    N, npy, seed = 50_000, 120, 10
    flat = OptionChain(ttms=ttms, forwards=np.ones(2), strikes_ttms=[K, K], optiontypes_ttms=[T, T], ids=np.array(["a", "b"]))
    prices, _ = pricer.model_mc_price_chain(flat, truth, nb_path=N, nb_steps=npy, use_rough_mc=True, seed=seed)
    vols = flat.compute_model_ivols_from_chain_data(model_prices=prices)
    chain = OptionChain(ttms=ttms, forwards=np.ones(2), strikes_ttms=[K, K], optiontypes_ttms=[T, T], ids=np.array(["a", "b"]),
                        bid_ivs=[v.copy() for v in vols], ask_ivs=[v.copy() for v in vols])
    start = LogSvParams(sigma0=0.8, theta=1.1, kappa1=4.0, kappa2=4.0, beta=0.2, volvol=1.3, H=0.45, weights=w, nodes=x)
    fit, info = pricer.calibrate_model_params_to_chain(chain, start, model_calibration_type=LogsvModelCalibrationType.PARAMS4,
                                                       calibration_engine=CalibrationEngine.ROUGH_MC, nb_path=N, nb_steps=npy, seed=seed, return_info=True)
    assert info["fun"] < 2e-6 and info["nit"] >= 3, info        # start objective ~1e-3; vol-of-vol is weakly identified on two maturities
    fit_vols = flat.compute_model_ivols_from_chain_data(model_prices=pricer.model_mc_price_chain(flat, fit, nb_path=N, nb_steps=npy, use_rough_mc=True, seed=seed)[0])
    assert max(np.max(np.abs(a - b)) for a, b in zip(fit_vols, vols)) < 3e-3
    assert fit.H == 0.45 and fit.weights is w


def test_rough_and_hawkes_edge_cases(cuda_lib):
    """ragged chain with an EMPTY slice, a single path, a single-step grid, 8 factors (the kernel's maximum) and a refused 9th; Hawkes with an
    empty slice and one path -- no crashes, finite outputs of the right shapes, errors where the boundary promises them"""
    from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, OptionChain
    from stochvolmodels_b200._capi import B200svError
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
    ttms = np.array([0.002, 0.05, 0.1])                       # first grid: int(0.002 * 100) + 1 = 1 step
    K = [np.array([0.9, 1.0, 1.1]), np.zeros(0), np.array([1.0])]
    T = [np.array(["P", "C", "IC"]), np.zeros(0, dtype="U2"), np.array(["IP"])]
    for P in (1, 33):
        Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(ttms, nb_path=P, nb_steps_per_year=100, seed=2)
        assert grids[0].size == 2
        for n in (1, 8):
            w, x = np.full(n, 1.0 / n), np.geomspace(1e-3, 50.0, n)
            kw = dict(ttms=ttms, forwards=np.ones(3), discfactors=np.ones(3), strikes_ttms=K, optiontypes_ttms=T, Z0=Z0, Z1=Z1, sigma0=0.4, theta=0.5,
                      kappa1=2.0, kappa2=1.0, beta=-0.2, orthog_vol=0.6, weights=w, nodes=x, timegrids=grids)
            p, e = rough_logsv_mc_chain_pricer_fixed_randoms(**kw)
            po, eo = rough.rough_chain_fixed(ttms, np.ones(3), np.ones(3), K, T, Z0, Z1, 0.4, 0.5, 2.0, 1.0, -0.2, 0.6, w, x, grids)
            assert [a.shape for a in p] == [(3,), (0,), (1,)] and [a.shape for a in e] == [(3,), (0,), (1,)]
            for m in (0, 2):
                np.testing.assert_allclose(p[m], po[m], rtol=1e-10, atol=1e-14)
                np.testing.assert_allclose(e[m], eo[m], rtol=1e-8, atol=1e-13)
    with pytest.raises((B200svError, ValueError)):
        rough_logsv_mc_chain_pricer_fixed_randoms(**{**kw, "weights": np.full(9, 1.0 / 9), "nodes": np.geomspace(1e-3, 50.0, 9)})
    with pytest.raises(ValueError):
        rough_logsv_mc_chain_pricer_fixed_randoms(**{**kw, "Z1": None})
    with pytest.raises(ValueError, match="unknown option payoff code"):
        rough_logsv_mc_chain_pricer_fixed_randoms(**{**kw, "optiontypes_ttms": [np.array(["P", "C", "X"]), T[1], T[2]]})
    # Hawkes: an empty slice through the function-level entry (OptionChain itself rejects empty slices, like the reference's), 1 / 257 paths
    from stochvolmodels_b200.pricers.hawkes_jd_pricer import hawkesjd_mc_chain_pricer
    d = {k: v for k, v in HawkesJDParams().to_dict().items() if k != "risk_premia_gamma"}
    for P in (1, 257):
        p, e = hawkesjd_mc_chain_pricer(ttms=np.array([0.01, 0.03]), forwards=np.ones(2), discfactors=np.ones(2),
                                        strikes_ttms=[np.zeros(0), np.array([0.95, 1.05])],
                                        optiontypes_ttms=[np.zeros(0, dtype="U2"), np.array(["P", "C"])], nb_path=P, seed=4, **d)
        assert p[0].shape == (0,) and p[1].shape == (2,) and np.all(np.isfinite(p[1])) and np.all(e[1] >= 0.0)


def test_rough_sharded_driver_equals_the_host_level_chain(cuda_lib):
    """multi_gpu.mc_chain_distributed('rough', ...) (device-level b200sv_dev_rough_logsv_slice, every maturity from t = 0 on its own grid,
    se_paths = 1) on one rank == b200sv_rough_logsv_mc_chain with in-kernel draws; two half-shards hold the unsharded terminal states."""
    import torch
    from stochvolmodels_b200 import LogSvParams, LogSVPricer, engine, get_btc_test_chain_data
    from stochvolmodels_b200 import _capi as C
    from stochvolmodels_b200.multi_gpu import CudaMcEngine, mc_chain_distributed
    from stochvolmodels_b200.utils.funcs import set_time_grid
    chain = get_btc_test_chain_data()
    p = LogSvParams(sigma0=0.8, theta=1.0, kappa1=2.2, kappa2=2.2, beta=0.2, volvol=1.6, H=0.3, weights=np.array([0.7, 0.5, 0.3]),
                    nodes=np.array([0.05, 1.5, 20.0]))
    N, npy, seed = 100_003, 360, 9
    for gauss in ("fp32", "fp64"):
        flags = engine.mc_flags("fp64", gauss)
        p_h, e_h = LogSVPricer().model_mc_price_chain(chain, p, nb_path=N, nb_steps=npy, use_rough_mc=True, seed=seed, gauss=gauss)
        grids = [set_time_grid(float(t), npy)[2] for t in chain.ttms]
        grid = [(g.size - 1, float(g[1] - g[0])) for g in grids]        # split_simulation.py:346
        pc = engine.logsv_params_c(p.sigma0, p.theta, p.kappa1, p.kappa2, p.beta, p.volvol)
        p_d, e_d = mc_chain_distributed("rough", pc, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms, chain.optiontypes_ttms,
                                        N, 0, True, C.LOG_RETURN, seed, flags, grid=grid, factors=(p.weights, p.nodes), se_paths=1)
        for m in range(len(p_h)):
            np.testing.assert_allclose(p_d[m], p_h[m], rtol=1e-12)
            np.testing.assert_allclose(e_d[m], e_h[m], rtol=1e-10)
    whole = CudaMcEngine("rough", pc, N, 0, flags, 4, factors=(p.weights, p.nodes))
    whole.simulate_slice(0, True, grid[1][0], grid[1][1], 1.0, True, float(chain.forwards[1]), seed)
    n0 = N // 3
    parts = [CudaMcEngine("rough", pc, n0, 0, flags, 4, factors=(p.weights, p.nodes)),
             CudaMcEngine("rough", pc, N - n0, n0, flags, 4, factors=(p.weights, p.nodes))]
    for e in parts:
        e.simulate_slice(0, True, grid[1][0], grid[1][1], 1.0, True, float(chain.forwards[1]), seed)
    torch.cuda.synchronize()
    assert whole.state.shape[0] == 5 and torch.equal(torch.cat([e.state for e in parts], dim=1), whole.state)
    with pytest.raises(ValueError):
        CudaMcEngine("rough", pc, 10, 0, flags, 4)                       # factors are required
