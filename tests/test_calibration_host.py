"""CPU tests of the calibration host logic (pricers/calibration.py): codec, constraints, weights, the batched objective's
forward-difference gradient and the SLSQP wiring -- with an analytic stand-in for the GPU pricer (no compute calls)."""
import numpy as np
import pytest

from stochvolmodels_b200 import LogSvParams, OptionChain
from stochvolmodels_b200.pricers import calibration as cal


def test_codec_params4_and_params5():
    p0 = LogSvParams(sigma0=0.8, theta=0.9, kappa1=4.0, kappa2=3.0, beta=0.1, volvol=1.2)
    lo = LogSvParams(sigma0=0.1, theta=0.1, kappa1=0.25, kappa2=0.25, beta=-3.0, volvol=0.2)
    hi = LogSvParams(sigma0=1.5, theta=1.5, kappa1=10.0, kappa2=10.0, beta=3.0, volvol=3.0)
    c4 = cal.LogSvParameterCodec(p0, lo, hi, cal.LogsvModelCalibrationType.PARAMS4)
    x0, b = c4.initial_and_bounds()
    np.testing.assert_array_equal(x0, [0.8, 0.9, 0.1, 1.2])
    assert b == ((0.1, 1.5), (0.1, 1.5), (-3.0, 3.0), (0.2, 3.0))
    q = c4.parse(np.array([0.5, 0.6, 0.7, 0.8]))
    assert (q.sigma0, q.theta, q.kappa1, q.kappa2, q.beta, q.volvol) == (0.5, 0.6, 4.0, 3.0, 0.7, 0.8)
    c5 = cal.LogSvParameterCodec(p0, lo, hi, cal.LogsvModelCalibrationType.PARAMS5)
    x0, b = c5.initial_and_bounds()
    np.testing.assert_array_equal(x0, [0.8, 0.9, 4.0, 0.1, 1.2])
    q = c5.parse(np.array([0.5, 0.8, 2.0, 0.7, 0.9]))
    assert q.kappa2 == 2.0 / 0.8                      # kappa2=None -> kappa1/theta (logsv_params.py:92-93)
    cv = cal.LogSvParameterCodec(p0, lo, hi, cal.LogsvModelCalibrationType.PARAMS_WITH_VARSWAP_FIT)
    x0, b = cv.initial_and_bounds()
    np.testing.assert_array_equal(x0, [0.1, 1.2])
    assert b == ((-3.0, 3.0), (0.2, 3.0))
    with pytest.raises(NotImplementedError):
        cal.LogSvParameterCodec(p0, lo, hi, cal.LogsvModelCalibrationType.PARAMS6).initial_and_bounds()


def test_constraints_match_theorem_3_7():
    p0 = LogSvParams(sigma0=0.8, theta=0.9, kappa1=4.0, kappa2=3.0, beta=0.1, volvol=1.2)
    c4 = cal.LogSvParameterCodec(p0, p0, p0, cal.LogsvModelCalibrationType.PARAMS4)
    x = np.array([0.5, 0.6, 0.7, 0.8])
    assert cal.build_logsv_constraints(c4, cal.ConstraintsType.UNCONSTRAINT) is None
    assert cal.build_logsv_constraints(c4, cal.ConstraintsType.MMA_MARTINGALE)["fun"](x) == pytest.approx(3.0 - 0.7)
    assert cal.build_logsv_constraints(c4, cal.ConstraintsType.INVERSE_MARTINGALE)["fun"](x) == pytest.approx(3.0 - 1.4)
    m4 = cal.build_logsv_constraints(c4, cal.ConstraintsType.INVERSE_MARTINGALE_MOMENT4)
    assert len(m4) == 2 and m4[1]["fun"](x) == pytest.approx((4.0 + 3.0 * 0.6) - 1.5 * (0.7 ** 2 + 0.8 ** 2))


def test_weights_and_chain_helpers():
    chain = OptionChain.get_uniform_chain(ttms=np.array([0.1, 0.5]), ids=np.array(["a", "b"]), forwards=np.array([1.0, 1.1]),
                                          strikes=np.array([0.9, 1.0, 1.1]), flat_vol=0.3)
    x, y = chain.get_chain_data_as_xy()
    assert x[0] is chain.ttms and all(np.all(v == 0.3) for v in y)
    np.testing.assert_allclose(chain.get_chain_atm_vols(), 0.3)
    vegas = chain.get_chain_vegas()
    d1 = np.log(1.1 / 1.0) / (0.3 * np.sqrt(0.5)) + 0.5 * 0.3 * np.sqrt(0.5)
    assert vegas[1][1] == pytest.approx(1.1 * np.exp(-0.5 * d1 * d1) / np.sqrt(2 * np.pi) * np.sqrt(0.5), rel=1e-14)
    unit = chain.get_chain_vegas(is_unit_ttm_vega=True)
    assert unit[0][1] != pytest.approx(vegas[0][1])
    mv = np.full(6, 0.3)
    w = cal.calibration_weights(chain, mv, True, False)
    assert w.shape == (6,) and w[:3].sum() == pytest.approx(1.0) and w[3:].sum() == pytest.approx(1.0)
    np.testing.assert_array_equal(cal.calibration_weights(chain, mv, False, False), np.ones(6))


def _quadratic_objective(bounds, target):
    A = np.array([[1.0, 0.2, 0.0], [0.0, 1.0, 0.3], [0.1, 0.0, 1.0], [0.5, 0.5, 0.5]])
    calls = []

    def batch_vols(points):
        calls.append(points.copy())
        return points @ A.T
    obj = cal.BatchedObjective(batch_vols=batch_vols, market_vols=A @ target, weights=np.array([1.0, 2.0, 0.5, 1.0]), bounds=bounds)
    return obj, A, calls


def test_batched_objective_value_gradient_and_memo():
    bounds = ((0.0, 2.0), (0.0, 2.0), (0.0, 1.0))
    target = np.array([0.7, 1.1, 0.4])
    obj, A, calls = _quadratic_objective(bounds, target)
    x = np.array([0.5, 1.0, 1.0])                         # third coordinate AT its upper bound -> backward step
    f, g = obj.fun(x), obj.jac(x)
    assert len(calls) == 1 and calls[0].shape == (4, 3)   # one batch of n+1 points served both calls
    np.testing.assert_array_equal(obj.steps(x), [cal.SLSQP_EPS, cal.SLSQP_EPS, -cal.SLSQP_EPS])
    assert np.all(calls[0][1:] <= np.array([b[1] for b in bounds]) + 0.0)          # never evaluated outside the box
    W = np.diag(obj.weights)
    r = A @ (x - target)
    assert f == pytest.approx(r @ W @ r, rel=1e-14)
    np.testing.assert_allclose(g, 2.0 * A.T @ W @ r, rtol=1e-6)
    obj.fun(x + 1e-3)
    assert len(calls) == 2
    nan_vols = lambda pts: np.where(np.arange(4)[None, :] == 0, np.nan, pts @ A.T)   # nansum semantics (logsv_pricer.py:292-294)
    o2 = cal.BatchedObjective(batch_vols=nan_vols, market_vols=A @ target, weights=obj.weights, bounds=bounds)
    assert np.isfinite(o2.fun(x))


def test_slsqp_recovers_target_with_batched_gradient():
    bounds = ((0.0, 2.0), (0.0, 2.0), (0.0, 1.0))
    target = np.array([0.7, 1.1, 0.4])
    obj, _, calls = _quadratic_objective(bounds, target)
    x, res = cal.run_slsqp(obj, np.array([1.5, 0.2, 0.9]), bounds)
    np.testing.assert_allclose(x, target, atol=2e-5)
    assert obj.nb_batches == len(calls) and res.success
    con = {"type": "ineq", "fun": lambda p: 0.6 - p[0]}   # active constraint: x0 <= 0.6
    obj2, _, _ = _quadratic_objective(bounds, target)
    x2, _ = cal.run_slsqp(obj2, np.array([0.3, 0.2, 0.9]), bounds, con)
    assert x2[0] == pytest.approx(0.6, abs=1e-6)


def test_validate_optimization_result():
    class R:
        success, message, x = True, "ok", np.array([0.5, 0.5])
    b = ((0.0, 1.0), (0.0, 1.0))
    np.testing.assert_array_equal(cal.validate_optimization_result(R, b), [0.5, 0.5])
    R.success = False
    with pytest.raises(cal.CalibrationError, match="Calibration failed"):
        cal.validate_optimization_result(R, b)
    R.success, R.x = True, np.array([0.5, 1.5])
    with pytest.raises(cal.CalibrationError, match="above bounds"):
        cal.validate_optimization_result(R, b)
    R.x = np.array([0.5, np.nan])
    with pytest.raises(cal.CalibrationError, match="non-finite"):
        cal.validate_optimization_result(R, b)
    R.x = np.array([0.5])
    with pytest.raises(cal.CalibrationError, match="wrong shape"):
        cal.validate_optimization_result(R, b)


def _varswap_chain(g):
    vols = [np.asarray(v) for v in g["market_vols"]]
    M = len(g["ttms"])
    return OptionChain(ttms=g["ttms"], ids=np.array([f"{t:0.2f}" for t in g["ttms"]]), forwards=g["forwards"], strikes_ttms=[g["strikes"]] * M,
                       optiontypes_ttms=[g["types"]] * M, bid_ivs=vols, ask_ivs=[v.copy() for v in vols])


def test_vol_moments_varswap_strikes_and_backbone_vs_reference_golden():
    """host pieces of PARAMS_WITH_VARSWAP_FIT against the reference's own outputs (tests/golden/make_golden.py --only-calib):
    Lambda matrix, moments, integrated moments, expected quadratic variance, replicated var-swap strikes, fitted eta backbone."""
    from conftest import load_golden
    from stochvolmodels_b200.pricers.logsv import vol_moments as vm
    g = load_golden("calib_logsv_varswap.npz")
    p = LogSvParams(*g["start"])
    np.testing.assert_allclose(vm.vol_moments_generator(p, 4), g["lambda4"], rtol=1e-15, atol=0)
    np.testing.assert_allclose([vm.vol_moments(p, t, 4) for t in (0.0, 0.04, 0.5, 2.0)], g["moments"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose([vm.vol_moments(p, t, 4, integrated=True) for t in (0.04, 0.5, 2.0)], g["int_moments"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose([vm.expected_qvar(p, t) for t in (0.0, 0.04, 0.25, 0.5, 2.0)], g["qvars"], rtol=1e-11)
    chain = _varswap_chain(g)
    np.testing.assert_allclose(chain.get_slice_varswap_strikes(False).to_numpy(), g["varswap_strikes_raw"], rtol=1e-12)
    vs = chain.get_slice_varswap_strikes(True)
    np.testing.assert_allclose(vs.to_numpy(), g["varswap_strikes"], rtol=1e-12)
    eta = vm.fit_vol_backbone_to_varswaps(p, vs)
    np.testing.assert_allclose(eta.to_numpy(), g["eta_start"], rtol=1e-10)
    np.testing.assert_array_equal(eta.index.to_numpy(), g["ttms"])
    # the codec attaches the refitted backbone to every parsed point
    codec = cal.LogSvParameterCodec(p, p, p, cal.LogsvModelCalibrationType.PARAMS_WITH_VARSWAP_FIT, vs)
    q = codec.parse(np.array([0.1, 1.2]))
    np.testing.assert_allclose(q.get_vol_backbone_etas(g["ttms"]), g["eta_start"], rtol=1e-10)
    assert (q.sigma0, q.theta, q.kappa1, q.kappa2) == (p.sigma0, p.theta, p.kappa1, p.kappa2)
