"""CPU: host-side logic of the boundary (no GPU): time grid rule, chain flattening, validation and error conventions,
default-step quirks, Black implied vols, path sharding."""
import inspect

import numpy as np
import pytest

from conftest import load_golden


def test_set_time_grid_matches_reference_rule():
    from stochvolmodels_b200.utils.funcs import set_time_grid
    for ttm, n, S, dt in load_golden("time_grid.npz")["cases"]:
        s2, dt2, grid = set_time_grid(ttm, int(n))
        assert s2 == int(S) and dt2 == dt and grid.shape == (s2 + 1,)


def test_flatten_chain_and_type_codes():
    from stochvolmodels_b200 import _capi as C
    off, k, t = C.flatten_chain([np.array([1.0, 2.0]), np.array([3.0])], [np.array(["C", "IP"]), np.array(["P"])])
    assert off.tolist() == [0, 2, 3] and k.tolist() == [1.0, 2.0, 3.0] and t.tolist() == [C.CALL, C.INV_PUT, C.PUT]
    assert t.dtype == np.int8 and off.dtype == np.int32
    with pytest.raises(ValueError, match="payoff"):
        C.encode_types(np.array(["X"]))
    assert [a.tolist() for a in C.split_chain(np.arange(3.0), off)] == [[0.0, 1.0], [2.0]]


def test_option_chain_validation_rules():
    from stochvolmodels_b200 import OptionChain
    K, T = np.array([0.9, 1.1]), np.array(["P", "C"])
    ok = OptionChain(ttms=np.array([0.1, 0.2]), forwards=np.ones(2), strikes_ttms=[K, K], optiontypes_ttms=[T, T])
    assert np.all(ok.discfactors == 1.0) and np.all(ok.discount_rates == 0.0)
    with pytest.raises(ValueError, match="strictly increasing"):
        OptionChain(ttms=np.array([0.2, 0.1]), forwards=np.ones(2), strikes_ttms=[K, K], optiontypes_ttms=[T, T])
    with pytest.raises(ValueError, match="positive"):
        OptionChain(ttms=np.array([0.1]), forwards=np.ones(1), strikes_ttms=[np.array([-1.0, 1.0])], optiontypes_ttms=[T])
    with pytest.raises(ValueError, match="optiontypes"):
        OptionChain(ttms=np.array([0.1]), forwards=np.ones(1), strikes_ttms=[K], optiontypes_ttms=[np.array(["P", "Z"])])
    with pytest.raises(ValueError, match="same length"):
        OptionChain(ttms=np.array([0.1, 0.2]), forwards=np.ones(2), strikes_ttms=[K], optiontypes_ttms=[T])
    u = OptionChain.get_uniform_chain(ttms=np.array([0.25, 0.5]), ids=np.array(["3m", "6m"]), forwards=np.array([1.0, 1.0]),
                                      strikes=np.array([0.8, 0.9, 1.0, 1.1, 1.2]))
    assert u.optiontypes_ttms[0].tolist() == ["P", "P", "C", "C", "C"]          # np.where(K >= F, 'C', 'P'), option_chain.py:492
    s = OptionChain.slice_to_chain(0.25, 1.0, K, T, discfactor=0.99)
    assert s.ttms.tolist() == [0.25] and s.discfactors.tolist() == [0.99]


def test_btc_chain_shape():
    from stochvolmodels_b200 import get_btc_test_chain_data
    c = get_btc_test_chain_data()
    assert [len(k) for k in c.strikes_ttms] == [12, 13, 15, 9] and c.ttms.shape == (4,)
    g = load_golden("logsv_fourier_btc.npz")
    for m in range(4):
        np.testing.assert_array_equal(c.strikes_ttms[m], g[f"strikes_{m}"])
        np.testing.assert_array_equal(c.optiontypes_ttms[m], g[f"types_{m}"])
    np.testing.assert_array_equal(c.ttms, g["ttms"])
    np.testing.assert_array_equal(c.forwards, g["forwards"])


def test_pricer_api_surface_matches_reference_signatures():
    """names / defaults of the reference methods (pricers/logsv_pricer.py:345-377, 590-596; heston_pricer.py:52-96)."""
    from stochvolmodels_b200 import HestonParams, HestonPricer, LogSvParams, LogSVPricer, ModelPricer, VariableType
    assert issubclass(LogSVPricer, ModelPricer) and issubclass(HestonPricer, ModelPricer)
    sig = inspect.signature(LogSVPricer.model_mc_price_chain)
    assert sig.parameters["nb_path"].default == 100000 and sig.parameters["nb_steps"].default is None
    assert sig.parameters["is_spot_measure"].default is True and sig.parameters["variable_type"].default == VariableType.LOG_RETURN
    sig = inspect.signature(LogSVPricer.simulate_terminal_values)
    assert sig.parameters["ttm"].default == 1.0 and sig.parameters["nb_path"].default == 100000
    sig = inspect.signature(HestonPricer.model_mc_price_chain)
    assert sig.parameters["nb_path"].default == 100000
    for name in ("price_chain", "compute_chain_prices_with_vols", "compute_model_ivols_for_chain", "price_slice", "price_vanilla",
                 "model_mc_price_chain", "simulate_terminal_values", "compute_mc_chain_implied_vols"):
        assert callable(getattr(LogSVPricer, name)) and callable(getattr(HestonPricer, name))
    p = LogSvParams(sigma0=0.2, theta=0.2, kappa1=1.0, kappa2=None)
    assert p.kappa2 == 5.0                                              # kappa2=None -> kappa1/theta (logsv_params.py:92-93)
    assert LogSvParams().get_vol_backbone_eta(0.3) == 1.0 and np.all(LogSvParams().get_vol_backbone_etas(np.array([0.1, 0.2])) == 1.0)
    assert HestonParams().v0 == 0.04 and HestonParams().rho == -0.5
    with pytest.raises(NotImplementedError):
        ModelPricer.model_mc_price_chain(LogSVPricer(), None, None)


def test_vol_backbone_lookup_is_equal_or_largest():
    import pandas as pd
    from stochvolmodels_b200 import LogSvParams
    p = LogSvParams()
    p.set_vol_backbone(pd.Series([0.9, 1.1], index=[0.1, 0.3]))
    assert p.get_vol_backbone_eta(0.1) == 0.9 and p.get_vol_backbone_eta(0.2) == 1.1 and p.get_vol_backbone_eta(0.3) == 1.1


def test_vol_scaler_and_phi_grid_host_side():
    from stochvolmodels_b200.pricers.logsv_pricer import set_vol_scaler
    from stochvolmodels_b200.utils.mgf_pricer import get_phi_grid
    g = load_golden("grids.npz")
    assert set_vol_scaler(1.0, np.array([0.25])) == float(g["vol_scaler_q_025"])
    np.testing.assert_array_equal(get_phi_grid(True, 1000, 0.2041241452319315), g["phi_mma"])
    np.testing.assert_array_equal(get_phi_grid(False, 1000, 0.2041241452319315), g["phi_inv"])


def test_error_conventions_raised_before_any_gpu_work():
    from stochvolmodels_b200 import engine, VariableType
    with pytest.raises(NotImplementedError):
        engine.variable_code(VariableType.SIGMA)
    with pytest.raises(ValueError, match="not implemented"):
        engine._check_fourier_types([np.array(["IC"])], True)
    engine._check_fourier_types([np.array(["IC", "C", "P", "IP"])], False)      # inverse measure accepts C == IC, P == IP
    with pytest.raises(ValueError):
        engine.mc_flags("fp16", "fp32")
    assert engine.mc_flags("fp64", "fp32") == 0 and engine.mc_flags("fp32", "fp64") == 3
    assert engine.fresh_seed() != engine.fresh_seed()


def test_fixed_randoms_helper_replays_and_leaves_global_state_untouched():
    """reference tests/test_logsv_characterization.py:583-602."""
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_chain_valuation
    np.random.seed(1)
    before = np.random.get_state()[1].copy()
    a = get_randoms_for_chain_valuation(np.array([0.1, 0.25]), nb_path=16, nb_steps_per_year=360, seed=10)
    b = get_randoms_for_chain_valuation(np.array([0.1, 0.25]), nb_path=16, nb_steps_per_year=360, seed=10)
    np.testing.assert_array_equal(np.random.get_state()[1], before)
    for u, v in zip(a[0] + a[1], b[0] + b[1]):
        np.testing.assert_array_equal(u, v)
    assert a[0][0].shape == (37, 16) and a[0][1].shape == (55, 16) and a[2][0] == 0.1 / 37
    g = load_golden("logsv_mc_fixed_g5_c1.npz")
    W0s, _, dts = get_randoms_for_chain_valuation(g["ttms"], 10000, 252, 10)
    np.testing.assert_array_equal(W0s[0][0, :3], g["W0_head"])
    assert dts[0] == g["dts"][0]


def test_oracle_black_implied_vols_round_trip_and_quickstart_values():
    from oracle import bsm
    K = np.array([0.6, 0.9, 1.0, 1.1, 1.6])
    types = np.array(["P", "P", "C", "C", "C"])
    for vol in (0.05, 0.2, 1.0, 2.5):
        p = bsm.compute_bsm_vanilla_price(1.0, K, 0.5, vol, types, 0.97)
        iv = bsm.infer_bsm_implied_vol(1.0, 0.5, K, p, types, 0.97)
        good = p > 1e-8
        np.testing.assert_allclose(iv[good], vol, rtol=1e-8)
    # examples/getting_started/quickstart.py:43-46
    np.testing.assert_allclose(bsm.infer_bsm_implied_vol(1.0, 0.25, [1.0], [0.197330882838064], ["C"]), 0.999577, rtol=5e-6)
    np.testing.assert_allclose(bsm.infer_bsm_implied_vol(1.0, 0.5, [1.0], [0.275201697631672], ["C"]), 0.995757, rtol=5e-6)
    assert np.isnan(bsm.infer_bsm_implied_vol(1.0, 0.5, [1.0], [1.5], ["C"]))[0]          # above the no-arbitrage bound


def test_shard_paths_partitions_exactly():
    from stochvolmodels_b200.multi_gpu import shard_paths
    for n, w in ((100, 8), (7, 8), (10**8, 8), (12345, 3), (5, 1)):
        parts = [shard_paths(n, w, r) for r in range(w)]
        assert sum(p[0] for p in parts) == n
        assert parts[0][1] == 0 and all(parts[r][1] == parts[r - 1][1] + parts[r - 1][0] for r in range(1, w))
        assert max(p[0] for p in parts) - min(p[0] for p in parts) <= 1


def test_model_pricer_mc_pdf_and_default_interfaces():
    """ModelPricer contract of the reference's tests/test_model_calibration_contracts.py:121-157 on a host-only toy pricer."""
    import pytest
    from stochvolmodels_b200 import ModelParams, ModelPricer

    class Flat(ModelPricer):
        def price_chain(self, option_chain, params, **kwargs):
            return [np.zeros(1)]

        def simulate_terminal_values(self, ttm, params, nb_path):
            return np.array([np.nan, np.inf, -np.inf, -0.2, -0.1, 0.0, 0.1, 0.2])

    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        density = Flat().get_log_return_mc_pdf(ttm=0.25, params=ModelParams(), x_grid=np.linspace(-0.5, 0.5, 51), nb_path=8)
    assert np.all(np.isfinite(density)) and np.all(density >= 0.0) and abs(np.sum(density) - 1.0) < 1e-14
    out = buf.getvalue()
    assert "num -inf = 1" in out and "num +inf = 1" in out and "num nans = 1" in out

    class PriceOnly(ModelPricer):
        def price_chain(self, option_chain, params, **kwargs):
            return [np.zeros(1)]
    for call in (lambda p: p.model_mc_price_chain(None, None), lambda p: p.calibrate_model_params_to_chain(None),
                 lambda p: p.simulate_vol_paths(None), lambda p: p.simulate_terminal_values(None), lambda p: p.compute_logreturn_pdf(None)):
        with pytest.raises(NotImplementedError):
            call(PriceOnly())


def test_rough_random_grid_replays_seed_and_has_aligned_shapes():
    """port of the reference's tests/test_rough_logsv_characterization.py::test_rough_random_grid_replays_seed_and_has_aligned_shapes:
    local RandomState (global numpy state untouched), Z0 / Z1 of the LAST maturity's length, grids end at the maturities"""
    from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation
    kwargs = dict(ttms=np.array([0.05, 0.1]), nb_path=16, nb_steps_per_year=100, seed=123)
    np.random.seed(91)
    expected_global_draw = np.random.random()
    np.random.seed(91)
    z0a, z1a, ga = get_randoms_for_rough_vol_chain_valuation(**kwargs)
    actual_global_draw = np.random.random()
    z0b, z1b, gb = get_randoms_for_rough_vol_chain_valuation(**kwargs)
    assert z0a.shape == z1a.shape == (11, 16)
    np.testing.assert_array_equal(z0a, z0b)
    np.testing.assert_array_equal(z1a, z1b)
    assert len(ga) == len(gb) == 2
    for a, b in zip(ga, gb):
        np.testing.assert_array_equal(a, b)
        assert a[0] == 0.0
    np.testing.assert_allclose([g[-1] for g in ga], kwargs["ttms"], rtol=0.0, atol=0.0)
    np.testing.assert_allclose(actual_global_draw, expected_global_draw, rtol=0.0, atol=0.0)
    # H = 1/2 kernel: the single node of the reference (logsv_params.py:110-113), deterministic
    from stochvolmodels_b200 import LogSvParams
    p, q = LogSvParams(H=0.5), LogSvParams(H=0.5)
    p.approximate_kernel(T=0.05); q.approximate_kernel(T=0.05)
    assert p.nodes.shape == p.weights.shape == (1,) and np.all(p.weights > 0) and np.all(p.nodes >= 0)
    np.testing.assert_array_equal(p.nodes, q.nodes)
    with pytest.raises(AssertionError):
        LogSvParams(H=0.7)


def test_logsv_params_helpers_and_vol_moment_entry_points_vs_reference_golden(capsys):
    """LogSvParams.gamma / eta / spatial grids / get_vol_moments_lambda and the reference-named functions of pricers/logsv/vol_moments
    (compute_analytic_vol_moments, compute_vol_moments_t, compute_expected_vol_t, compute_sqrt_qvar_t, fit_model_vol_backbone_to_varswaps)
    against the reference's own outputs (logsv_params_helpers.npz, three parameter sets)"""
    import pandas as pd
    from conftest import load_golden
    from stochvolmodels_b200 import LogSvParams, VariableType
    from stochvolmodels_b200.pricers.logsv import vol_moments as vm
    g = load_golden("logsv_params_helpers.npz")
    ts = g["ts"]
    for i, row in enumerate(g["sets"]):
        p = LogSvParams(*row)
        np.testing.assert_allclose([p.gamma, p.eta, p.kappa, p.theta2, p.vartheta2], g[f"scalars_{i}"], rtol=1e-15)
        for name, vt in (("x", VariableType.LOG_RETURN), ("sigma", VariableType.SIGMA), ("qvar", VariableType.Q_VAR)):
            np.testing.assert_allclose(p.get_variable_space_grid(variable_type=vt, ttm=0.7, n_stdevs=2.5, n=37), g[f"grid_{name}_{i}"], rtol=1e-14,
                                       atol=1e-15)
        np.testing.assert_allclose(p.get_x_grid(), g[f"grid_x_default_{i}"], rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(p.get_vol_moments_lambda(4), g[f"lambda4_{i}"], rtol=1e-15)
        np.testing.assert_allclose(p.get_vol_moments_lambda(n_terms=8), g[f"lambda8_{i}"], rtol=1e-15)
        np.testing.assert_allclose(vm.compute_vol_moments_t(params=p, ttm=ts, n_terms=4), g[f"moments_{i}"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(vm.compute_vol_moments_t(params=p, ttm=ts[1:], n_terms=8), g[f"moments8_{i}"], rtol=1e-7, atol=1e-10)
        ints = np.array([vm.compute_analytic_vol_moments(params=p, t=t, n_terms=4, is_qvar=True) for t in ts[1:]])
        np.testing.assert_allclose(ints, g[f"int_moments_{i}"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(vm.compute_expected_vol_t(params=p, t=ts), g[f"expected_vol_{i}"], rtol=1e-10)
        np.testing.assert_allclose(vm.compute_sqrt_qvar_t(params=p, t=ts), g[f"sqrt_qvar_{i}"], rtol=1e-10)
        idx, k = g[f"varswap_strikes_{i}"]
        eta = vm.fit_model_vol_backbone_to_varswaps(log_sv_params=p, varswap_strikes=pd.Series(k, index=idx), verbose=(i == 0))
        np.testing.assert_allclose(eta.to_numpy(), g[f"backbone_{i}"], rtol=1e-9)
        np.testing.assert_array_equal(eta.index.to_numpy(), idx)
    with pytest.raises(NotImplementedError):
        p.get_variable_space_grid(variable_type=7)
    p.assert_vol_moments_stability()
    p.print_vol_moments_stability()
    out = capsys.readouterr().out
    assert "vol moments stable = True" in out and "lambda_m" in out and "vars_swaps" in out


def test_compute_integration_weights_properties_and_errors():
    """the reference's own assertions about its public quadrature helper (its tests/test_numerical_utilities.py:9-60): Simpson integrates
    cubics exactly, trapezoid integrates linears, even-sized / short / non-uniform / decreasing / non-finite grids raise ValueError"""
    from stochvolmodels_b200 import compute_integration_weights
    x = np.linspace(-2.0, 2.0, 9)
    w = compute_integration_weights(0.5 + 1j * x, is_simpson=True)
    np.testing.assert_allclose([w @ np.ones(9), w @ x, w @ x ** 2, w @ x ** 3], [4.0, 0.0, 16.0 / 3.0, 0.0], atol=1e-14)
    y = np.linspace(0.0, 1.0, 5)
    w = compute_integration_weights(-0.5 + 1j * y, is_simpson=False)
    np.testing.assert_allclose([w @ np.ones(5), w @ y], [1.0, 0.5], atol=1e-14)
    with pytest.raises(ValueError, match="odd"):
        compute_integration_weights(1j * np.linspace(0.0, 1.0, 4), is_simpson=True)
    for grid, simpson, message in ((np.array([0.0]), False, "too short"), (np.array([0.0, 0.5, 1.1]), True, "uniform"),
                                   (np.array([0.0, 0.5, 0.4]), True, "increasing"), (np.array([0.0, np.nan, 1.0]), True, "finite")):
        with pytest.raises(ValueError, match=message):
            compute_integration_weights(1j * grid, is_simpson=simpson)


def test_set_seed_makes_seedless_calls_reproducible():
    from stochvolmodels_b200 import engine, set_seed
    set_seed(8)
    a = [engine.fresh_seed() for _ in range(3)]
    set_seed(8)
    b = [engine.fresh_seed() for _ in range(3)]
    set_seed(9)
    c = engine.fresh_seed()
    set_seed(None)
    d, e = engine.fresh_seed(), engine.fresh_seed()
    assert a == b and len(set(a)) == 3 and c != a[0] and d != e and all(0 <= s < 1 << 64 for s in a + [c, d, e])


def test_v0_implied_short_maturity_helpers():
    """the reference's assertions about v0_implied (its tests/test_logsv_characterization.py:680-700) + the Heston placeholder"""
    from stochvolmodels_b200.pricers.heston_pricer import v0_implied as heston_v0
    from stochvolmodels_b200.pricers.logsv_pricer import LOGSV_BTC_PARAMS as p, v0_implied
    np.testing.assert_allclose(v0_implied(p.sigma0, 1.1, p.volvol, p.theta, p.kappa1, 0.02), p.sigma0 - (1.1 ** 2 + p.volvol ** 2) * 0.02 / 4.0)
    np.testing.assert_allclose(v0_implied(p.sigma0, 0.0, p.volvol, p.theta, p.kappa1, 0.02), p.sigma0 - p.volvol ** 2 * 0.02 / 4.0)
    v = v0_implied(p.sigma0, p.beta, p.volvol, p.theta, p.kappa1, 0.02)
    assert np.isfinite(v) and abs(v - p.sigma0) < 0.05
    # root of the reference's quadratic: 6 beta t v^2 + (24 + beta^2 t + 2 vartheta^2 t - 12 kappa1 t) v + 12 (theta kappa1 t - 2 atm) = 0
    t, b = 0.02, p.beta
    res = 6 * b * t * v * v + (24 + b * b * t + 2 * p.vartheta2 * t - 12 * p.kappa1 * t) * v + 12 * (p.theta * p.kappa1 * t - 2 * p.sigma0)
    assert abs(res) < 1e-10
    assert heston_v0(0.04, 0.5, 0.1) == 0.04 - 0.25 * 0.1 / 8.0


def test_option_chain_slice_views_and_strike_transforms_vs_reference_golden(capsys):
    """OptionChain.get_slice / get_slices_as_chain / to_forward_normalised_strikes / to_uniform_strikes and the BTC sample chain's market
    quotes against the reference's own outputs (option_chain_transforms.npz); OptionSlice validation"""
    from conftest import load_golden
    from stochvolmodels_b200 import OptionChain, get_btc_test_chain_data
    from stochvolmodels_b200.data.option_chain import OptionSlice
    g = load_golden("option_chain_transforms.npz")
    c = get_btc_test_chain_data()
    np.testing.assert_array_equal(np.asarray(c.ids), g["ids"])
    np.testing.assert_array_equal(c.ttms, g["ttms"])
    np.testing.assert_array_equal(c.forwards, g["forwards"])
    n, u = OptionChain.to_forward_normalised_strikes(c), OptionChain.to_uniform_strikes(c, num_strikes=7)
    for m in range(4):
        np.testing.assert_array_equal(c.strikes_ttms[m], g[f"strikes_{m}"])
        np.testing.assert_array_equal(c.bid_ivs[m], g[f"bid_ivs_{m}"])
        np.testing.assert_array_equal(c.ask_ivs[m], g[f"ask_ivs_{m}"])
        np.testing.assert_array_equal(n.strikes_ttms[m], g[f"norm_strikes_{m}"])
        np.testing.assert_array_equal(u.strikes_ttms[m], g[f"uni_strikes_{m}"])
        np.testing.assert_array_equal(np.asarray(u.optiontypes_ttms[m]), g[f"uni_types_{m}"])
    np.testing.assert_array_equal(n.forwards, g["norm_forwards"])
    np.testing.assert_array_equal(n.forwards0, g["norm_forwards0"])
    assert u.bid_ivs is None and n.bid_ivs is c.bid_ivs
    sub = OptionChain.get_slices_as_chain(c, ids=["3m", "1m"])
    np.testing.assert_array_equal(sub.ttms, g["sub_ttms"])
    np.testing.assert_array_equal(sub.forwards, g["sub_forwards"])
    np.testing.assert_array_equal(np.asarray(sub.ids), g["sub_ids"])
    np.testing.assert_array_equal(sub.strikes_ttms[0], g["sub_strikes_0"])
    np.testing.assert_array_equal(sub.bid_ivs[1], g["sub_bid_1"])
    one = OptionChain.get_slices_as_chain(c, ids=["2m"])
    np.testing.assert_array_equal(one.ttms, g["one_ttms"])
    np.testing.assert_array_equal(one.strikes_ttms[0], g["one_strikes"])
    sl = c.get_slice("2w")
    np.testing.assert_allclose([sl.ttm, sl.forward, sl.discfactor, sl.discount_rate], g["slice_scalars"], rtol=0, atol=0)
    np.testing.assert_array_equal(sl.ask_ivs, g["slice_ask"])
    np.testing.assert_array_equal(c.get_mid_vols()[0], g["mid_vols_0"])
    with pytest.raises(ValueError):
        c.get_slice("9y")
    k, t = np.array([0.9, 1.0]), np.array(["P", "C"])
    assert OptionSlice(ttm=0.5, forward=1.0, strikes=k, optiontypes=t, id="x", discount_rate=0.02).discfactor == np.exp(-0.01)
    for bad in (dict(ttm=-1.0), dict(bid_ivs=np.array([0.3, 0.2]), ask_ivs=np.array([0.2, 0.3])), dict(bid_prices=np.array([-1.0, 0.1])),
                dict(ask_ivs=np.array([0.2])), dict(discfactor=0.0)):
        with pytest.raises(ValueError):
            OptionSlice(**{**dict(ttm=0.5, forward=1.0, strikes=k, optiontypes=t, id="x"), **bad})
    c.print()
    assert "strikes_ttms" in capsys.readouterr().out


def test_chain_deltas_and_skews():
    """Black-76 deltas at the mid vols (put-call parity in delta, monotone in strike) and the delta-interpolated skew: on a flat-vol chain the
    skew is zero, on a downward-sloping smile it is positive; a one-sided slice raises like the reference (option_chain.py:301-302)"""
    from stochvolmodels_b200 import OptionChain, get_btc_test_chain_data
    c = get_btc_test_chain_data()
    d = c.get_chain_deltas()
    for m in range(4):
        puts = np.asarray(c.optiontypes_ttms[m]) == "P"
        assert np.all(d[m][puts] < 0) and np.all(d[m][~puts] > 0) and np.all(np.abs(d[m]) < 1)
        n_d1 = np.where(puts, d[m] + 1.0, d[m])                   # N(d1) falls with the strike
        assert np.all(np.diff(n_d1) < 0)
    assert c.get_chain_skews(0.25).shape == (4,)
    K = np.linspace(0.7, 1.4, 15)
    types = np.where(K >= 1.0, "C", "P")
    flat = OptionChain(ttms=np.array([0.25]), forwards=np.ones(1), strikes_ttms=[K], optiontypes_ttms=[types], bid_ivs=[0.5 * np.ones(15)],
                       ask_ivs=[0.5 * np.ones(15)])
    np.testing.assert_allclose(flat.get_chain_skews(), 0.0, atol=1e-15)
    smile = 0.5 - 0.3 * np.log(K)
    down = OptionChain(ttms=np.array([0.25]), forwards=np.ones(1), strikes_ttms=[K], optiontypes_ttms=[types], bid_ivs=[smile], ask_ivs=[smile])
    assert down.get_chain_skews()[0] > 0.05
    calls_only = OptionChain(ttms=np.array([0.25]), forwards=np.ones(1), strikes_ttms=[K], optiontypes_ttms=[np.array(["C"] * 15)], bid_ivs=[smile],
                             ask_ivs=[smile])
    with pytest.raises(ValueError, match="both put and call"):
        calls_only.get_chain_skews()


def test_small_host_utilities_of_utils_funcs():
    from stochvolmodels_b200.utils.funcs import compute_histogram_data, find_nearest, ncdf, npdf, update_kwargs, erfcc
    a = np.array([0.1, 0.25, 0.5, 1.0])
    assert find_nearest(a, 0.3) == 0.25 and find_nearest(a, 0.3, is_equal_or_largest=True) == 0.5 and find_nearest(a, 0.25, is_equal_or_largest=True) == 0.25
    assert find_nearest(a, 5.0) == 1.0 and find_nearest(a[::-1], 0.45, is_sorted=False) == 0.5 and find_nearest(a, 0.375) == 0.5      # tie: upper
    assert update_kwargs({"a": 1}, None) == {"a": 1} and update_kwargs({"a": 1}, {"a": 2, "b": 3}) == {"a": 2, "b": 3}
    np.testing.assert_allclose([ncdf(0.0), ncdf(1.959963984540054), npdf(0.0), erfcc(0.0)], [0.5, 0.975, 1 / np.sqrt(2 * np.pi), 1.0], rtol=1e-12)
    h = compute_histogram_data(np.array([0.1, 0.2, 0.2, 0.9]), np.linspace(0.0, 1.0, 5))
    assert h.shape == (5,) and h.iloc[1] == 0.75 and h.iloc[4] == 0.25 and h.index[-1] == 1.0
