"""Calibration drivers on top of the batched GPU chain pricers (SURVEY.md §8f "next" #2).

Mirrors ``LogSVPricer.calibrate_model_params_to_chain`` (reference pricers/logsv_pricer.py:441-558 with the codec :103-193, the
weights :196-207, the objective :222-294 and the constraints :297-330) and ``HestonPricer.calibrate_model_params_to_chain``
(pricers/heston_pricer.py:111-180): same arguments, bounds, objective (vega-weighted squared implied-vol errors, ``nansum``), SLSQP
with ``ftol=1e-8``, result validation and return type.

What is different is how the optimizer gets its numbers.  The reference lets scipy difference the objective: n+1 sequential chain
pricings (~2 s each on the CPU) per SLSQP iteration.  Here every evaluation point x is priced TOGETHER with its n forward-difference
neighbours x + h e_i as ONE batch of n+1 parameter sets (``b200sv_logsv_price_chain_batch``: (n+1) x 1000 ODE threads, Fourier sums
and the Black inversion fused, one launch pair, one D2H copy), so the objective and its gradient cost one ~1 ms GPU call.  The step is
SLSQP's own default (``eps = 1.4901161193847656e-08``, absolute, flipped at an upper bound), so the gradient is the number scipy would
have computed from the same objective.

The MC engine (:251-266) has two sources of fixed normals: ``mc_randoms="numpy"`` (default) draws them exactly as the reference does
(``np.random.RandomState(seed)``, :1051-1074), keeps them resident in HBM (``DeviceRandoms``) and prices each of the n+1 sets with the
fixed-random stepper -- the same objective values as the reference for the same seed; ``mc_randoms="philox"`` prices the n+1 sets in one
``b200sv_logsv_mc_chain_batch`` call in which the counter-based generator re-draws identical normals for every set (nothing stored or
streamed; statistically equivalent objective, different sample).

Quotes whose model price falls outside the no-arbitrage bounds have no implied vol: the GPU inversion returns NaN for them and the
objective's ``nansum`` (the reference's own reduction, :292-294) leaves them out of that evaluation, exactly as the reference does with the NaNs
of its third-party inverter.  Which quotes that inverter would flag is not pinned (the package is absent); ``BatchedObjective``
counts the skipped quotes per evaluation in ``nan_quotes`` so that a caller can see when a fit was driven by a shrinking set of quotes.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .. import engine
from ..utils.funcs import to_flat_np_array

SLSQP_EPS = 1.4901161193847656e-08      # scipy.optimize._slsqp_py default `eps` (sqrt of the double epsilon)


class LogsvModelCalibrationType(Enum):
    """which parameters the calibration solves for (reference logsv_pricer.py:56-68)."""
    PARAMS4 = 1                  # sigma0, theta, beta, volvol; kappa1, kappa2 fixed
    PARAMS5 = 2                  # sigma0, theta, kappa1, beta, volvol; kappa2 = kappa1 / theta
    PARAMS6 = 3
    PARAMS_WITH_VARSWAP_FIT = 4  # beta, volvol; the eta backbone is re-fitted to the chain's variance-swap strikes at every point


class ConstraintsType(Enum):
    """parameter constraints (reference :71-90)."""
    UNCONSTRAINT = 1
    MMA_MARTINGALE = 2               # kappa2 >= beta
    INVERSE_MARTINGALE = 3           # kappa2 >= 2 beta
    MMA_MARTINGALE_MOMENT4 = 4       # + kappa >= 1.5 vartheta^2
    INVERSE_MARTINGALE_MOMENT4 = 5


class CalibrationEngine(Enum):
    """how model vols are produced inside the objective (reference :93-102)."""
    ANALYTIC = 1
    MC = 2
    ROUGH_MC = 3                     # rough-LogSV multi-factor MC with fixed normals (reference logsv_pricer.py:266-289)


class CalibrationError(RuntimeError):
    """raised when the optimizer fails or returns an unusable vector (reference pricers/model_pricer.py:44-80)."""


def validate_optimization_result(result, bounds) -> np.ndarray:
    """the optimizer's vector as float64[len(bounds)], or ``CalibrationError`` naming what is wrong with it.

    Same acceptance rule and error wording as the reference's guard (pricers/model_pricer.py:48-80: success flag, numeric 1-d vector
    of the right length, finite, inside the bounds up to 1e-10) -- callers match on the message text."""
    note = str(getattr(result, "message", "no optimizer message"))

    def reject(what: str, cause=None):
        raise CalibrationError(f"Calibration {what}: {note}") from cause

    if not getattr(result, "success", False):
        reject("failed")
    x = getattr(result, "x", None)
    if x is None:
        reject("returned no parameter vector")
    try:
        x = np.array(x, dtype=np.float64, copy=True)
    except (TypeError, ValueError) as exc:
        reject("returned a non-numeric parameter vector", exc)
    if x.shape != (len(bounds),):
        reject("returned a parameter vector with the wrong shape")
    if not np.isfinite(x).all():
        reject("returned non-finite parameters")
    slack = 1.0e-10
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
    if (x < lo - slack).any():
        reject("returned parameters below bounds")
    if (x > hi + slack).any():
        reject("returned parameters above bounds")
    return x


def calibration_weights(option_chain, market_vols: np.ndarray, is_vega_weighted: bool, is_unit_ttm_vega: bool) -> np.ndarray:
    """flattened objective weights: per-slice normalised Black vegas or ones (reference logsv_pricer.py:196-207)."""
    if not is_vega_weighted:
        return np.ones_like(market_vols)
    vegas_ttms = option_chain.get_chain_vegas(is_unit_ttm_vega=is_unit_ttm_vega)
    return to_flat_np_array([v / sum(v) for v in vegas_ttms])


@dataclass
class BatchedObjective:
    """objective + forward-difference gradient from ONE batched evaluation, memoised on the last point so that SLSQP's separate
    ``fun(x)`` / ``jac(x)`` calls share it.  ``batch_vols(points [B, n]) -> model vols [B, J]``."""
    batch_vols: Callable[[np.ndarray], np.ndarray]
    market_vols: np.ndarray
    weights: np.ndarray
    bounds: Sequence[Tuple[float, float]]
    eps: float = SLSQP_EPS
    nb_batches: int = 0
    nan_quotes: int = 0          # quotes without an implied vol at the LAST evaluation point (dropped there by nansum, as in the reference)
    _x: Optional[np.ndarray] = None
    _f: float = np.nan
    _g: Optional[np.ndarray] = None

    def steps(self, x: np.ndarray) -> np.ndarray:
        """+eps, or -eps where x + eps would leave the box (scipy approx_derivative's one-sided bound handling)."""
        ub = np.array([np.inf if b[1] is None else b[1] for b in self.bounds], dtype=float)
        return np.where(x + self.eps > ub, -self.eps, self.eps)

    def values(self, vols: np.ndarray) -> np.ndarray:
        return np.nansum(self.weights[None, :] * np.square(vols - self.market_vols[None, :]), axis=1)

    def _evaluate(self, x: np.ndarray) -> None:
        x = np.asarray(x, dtype=float)
        if self._x is not None and np.array_equal(x, self._x):
            return
        h = self.steps(x)
        pts = np.vstack([x[None, :], x[None, :] + np.diag(h)])
        vols = self.batch_vols(pts)
        f = self.values(vols)
        self.nb_batches += 1
        self.nan_quotes = int(np.count_nonzero(np.isnan(vols[0])))
        self._x, self._f, self._g = x.copy(), float(f[0]), (f[1:] - f[0]) / h

    def fun(self, x: np.ndarray, *args) -> float:
        self._evaluate(x)
        return self._f

    def jac(self, x: np.ndarray, *args) -> np.ndarray:
        self._evaluate(x)
        return self._g.copy()


def run_slsqp(objective: BatchedObjective, p0: np.ndarray, bounds, constraints=None, disp: bool = False):
    """``scipy.optimize.minimize(method='SLSQP', options={'ftol': 1e-8})`` as in the reference (:543-555), with the batched gradient."""
    from scipy.optimize import minimize      # host-side optimizer; scipy is already a dependency of the reference
    kwargs = dict(method="SLSQP", jac=objective.jac, bounds=bounds, options={"disp": disp, "ftol": 1e-8})
    if constraints is not None:
        kwargs["constraints"] = constraints
    result = minimize(objective.fun, np.asarray(p0, dtype=float), **kwargs)
    return validate_optimization_result(result, bounds), result


# ---- LogSV ------------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class LogSvParameterCodec:
    """optimizer vector <-> LogSvParams (reference _LogSvParameterCodec, :103-193)."""
    params0: "LogSvParams"
    params_min: "LogSvParams"
    params_max: "LogSvParams"
    calibration_type: LogsvModelCalibrationType
    varswap_strikes: object = None          # pandas Series (PARAMS_WITH_VARSWAP_FIT only)

    def parse(self, pars: np.ndarray):
        from .logsv.vol_moments import fit_vol_backbone_to_varswaps
        from .logsv_pricer import LogSvParams
        p0 = self.params0
        rough = dict(H=getattr(p0, "H", 0.5), nodes=getattr(p0, "nodes", None), weights=getattr(p0, "weights", None))   # reference :125-127
        if self.calibration_type == LogsvModelCalibrationType.PARAMS_WITH_VARSWAP_FIT:
            out = LogSvParams(sigma0=p0.sigma0, theta=p0.theta, kappa1=p0.kappa1, kappa2=p0.kappa2, beta=pars[0], volvol=pars[1], **rough)
            out.set_vol_backbone(fit_vol_backbone_to_varswaps(out, self.varswap_strikes))     # reference :145-160
            return out
        if self.calibration_type == LogsvModelCalibrationType.PARAMS4:
            out = LogSvParams(sigma0=pars[0], theta=pars[1], kappa1=p0.kappa1, kappa2=p0.kappa2, beta=pars[2], volvol=pars[3], **rough)
        elif self.calibration_type == LogsvModelCalibrationType.PARAMS5:
            out = LogSvParams(sigma0=pars[0], theta=pars[1], kappa1=pars[2], kappa2=None, beta=pars[3], volvol=pars[4], **rough)
        else:
            raise NotImplementedError(f"{self.calibration_type}")
        # no vol_backbone on the trial point: the reference codec (logsv_pricer.py:112-137) builds PARAMS4 / PARAMS5 candidates without
        # one, so eta = 1 during the fit and the fitted object carries none, whatever params0 holds
        return out

    def initial_and_bounds(self) -> Tuple[np.ndarray, Tuple[Tuple[float, float], ...]]:
        p0, lo, hi = self.params0, self.params_min, self.params_max
        if self.calibration_type == LogsvModelCalibrationType.PARAMS4:
            names = ("sigma0", "theta", "beta", "volvol")
        elif self.calibration_type == LogsvModelCalibrationType.PARAMS5:
            names = ("sigma0", "theta", "kappa1", "beta", "volvol")
        elif self.calibration_type == LogsvModelCalibrationType.PARAMS_WITH_VARSWAP_FIT:
            names = ("beta", "volvol")
        else:
            raise NotImplementedError(f"{self.calibration_type}")
        return (np.array([getattr(p0, n) for n in names], dtype=float),
                tuple((getattr(lo, n), getattr(hi, n)) for n in names))


def build_logsv_constraints(codec: LogSvParameterCodec, constraints_type: ConstraintsType):
    """SLSQP inequality constraints of Theorem 3.7 (reference :297-330)."""
    def martingale_measure(pars):
        p = codec.parse(pars)
        return p.kappa2 - p.beta

    def inverse_measure(pars):
        p = codec.parse(pars)
        return p.kappa2 - 2.0 * p.beta

    def vol_4thmoment_finite(pars):
        p = codec.parse(pars)
        return (p.kappa1 + p.kappa2 * p.theta) - 1.5 * p.vartheta2

    if constraints_type == ConstraintsType.UNCONSTRAINT:
        return None
    if constraints_type == ConstraintsType.MMA_MARTINGALE:
        return {"type": "ineq", "fun": martingale_measure}
    if constraints_type == ConstraintsType.INVERSE_MARTINGALE:
        return {"type": "ineq", "fun": inverse_measure}
    if constraints_type == ConstraintsType.MMA_MARTINGALE_MOMENT4:
        return ({"type": "ineq", "fun": martingale_measure}, {"type": "ineq", "fun": vol_4thmoment_finite})
    if constraints_type == ConstraintsType.INVERSE_MARTINGALE_MOMENT4:
        return ({"type": "ineq", "fun": inverse_measure}, {"type": "ineq", "fun": vol_4thmoment_finite})
    raise NotImplementedError(f"{constraints_type}")


def calibrate_logsv(pricer, option_chain, params0, params_min, params_max, is_vega_weighted: bool, is_unit_ttm_vega: bool,
                    model_calibration_type: LogsvModelCalibrationType, constraints_type: ConstraintsType,
                    calibration_engine: CalibrationEngine, nb_path: int, nb_steps: int, seed: int, is_spot_measure: bool = True,
                    disp: bool = False, return_info: bool = False, fd_step: Optional[float] = None, mc_randoms: str = "numpy"):
    from .logsv_pricer import (DeviceRandoms, _fixed_randoms_chain_device, _params_c, get_randoms_for_chain_valuation)
    vol_scaler = pricer.set_vol_scaler(option_chain=option_chain)
    _, market_vols_ttms = option_chain.get_chain_data_as_xy()
    market_vols = to_flat_np_array(market_vols_ttms)
    weights = calibration_weights(option_chain, market_vols, is_vega_weighted, is_unit_ttm_vega)
    varswap_strikes = (option_chain.get_slice_varswap_strikes(floor_with_atm_vols=True)
                       if model_calibration_type == LogsvModelCalibrationType.PARAMS_WITH_VARSWAP_FIT else None)
    codec = LogSvParameterCodec(params0, params_min, params_max, model_calibration_type, varswap_strikes)
    p0, bounds = codec.initial_and_bounds()
    ttms = option_chain.ttms

    if calibration_engine == CalibrationEngine.ANALYTIC:
        def batch_vols(points: np.ndarray) -> np.ndarray:
            sets = [codec.parse(p) for p in points]
            etas = np.array([s.get_vol_backbone_etas(ttms=ttms) for s in sets], dtype=float)
            _, ivols = engine.logsv_price_chain_batch([_params_c(s) for s in sets], ttms, option_chain.forwards, option_chain.discfactors,
                                                      etas, option_chain.strikes_ttms, option_chain.optiontypes_ttms,
                                                      is_spot_measure=is_spot_measure, vol_scaler=vol_scaler)
            return ivols
    elif calibration_engine == CalibrationEngine.MC and mc_randoms == "philox":
        flags = engine.mc_flags("fp64", "fp32")

        def batch_vols(points: np.ndarray) -> np.ndarray:
            sets = [codec.parse(p) for p in points]
            etas = np.array([s.get_vol_backbone_etas(ttms=ttms) for s in sets], dtype=float)
            _, _, ivols = engine.logsv_mc_chain_batch([_params_c(s) for s in sets], ttms, option_chain.forwards, option_chain.discfactors, etas,
                                                      option_chain.strikes_ttms, option_chain.optiontypes_ttms, nb_path, nb_steps,
                                                      is_spot_measure, seed, flags)
            return ivols
    elif calibration_engine == CalibrationEngine.MC:
        if mc_randoms != "numpy":
            raise ValueError("mc_randoms must be 'numpy' or 'philox'")
        rnd = get_randoms_for_chain_valuation(ttms=ttms, nb_path=nb_path, nb_steps_per_year=nb_steps, seed=seed, device=True)
        assert isinstance(rnd, DeviceRandoms)

        def batch_vols(points: np.ndarray) -> np.ndarray:
            rows = []
            for p in points:
                s = codec.parse(p)
                prices, _ = _fixed_randoms_chain_device(rnd, ttms, option_chain.forwards, option_chain.discfactors, option_chain.strikes_ttms,
                                                        option_chain.optiontypes_ttms, _params_c(s), s.get_vol_backbone_etas(ttms=ttms),
                                                        is_spot_measure, 1, True)
                rows.append(to_flat_np_array(option_chain.compute_model_ivols_from_chain_data(model_prices=prices)))
            return np.vstack(rows)
    elif calibration_engine == CalibrationEngine.ROUGH_MC:
        # reference :528-533 + :266-289: fixed normals of get_randoms_for_rough_vol_chain_valuation; here the n + 1 finite-difference
        # parameter sets of an optimizer step share ONE upload of them (or none: mc_randoms="philox" draws in-kernel on a fixed seed)
        from .logsv_pricer import get_randoms_for_rough_vol_chain_valuation
        if mc_randoms not in ("numpy", "philox"):
            raise ValueError("mc_randoms must be 'numpy' or 'philox'")
        if params0.weights is None or params0.nodes is None:
            raise ValueError("ROUGH_MC needs params0.weights / params0.nodes (LogSvParams.approximate_kernel)")
        Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(ttms=ttms, nb_path=nb_path, nb_steps_per_year=nb_steps, seed=seed)
        if mc_randoms == "philox":
            Z0 = Z1 = None
        nsteps = [int(g.size) - 1 for g in grids]
        hs = [float(g[1] - g[0]) for g in grids]
        flags = engine.mc_flags("fp64", "fp32")

        def batch_vols(points: np.ndarray) -> np.ndarray:
            sets = [codec.parse(p) for p in points]
            _, _, ivols, _, _ = engine.rough_logsv_mc_chain([_params_c(s) for s in sets], params0.weights, params0.nodes, ttms, option_chain.forwards,
                                                            option_chain.discfactors, option_chain.strikes_ttms, option_chain.optiontypes_ttms,
                                                            nb_path, nsteps, hs, Z0, Z1, 1, seed, flags, with_ivols=True)
            return ivols
    else:
        raise NotImplementedError(f"{calibration_engine}")

    objective = BatchedObjective(batch_vols=batch_vols, market_vols=market_vols, weights=weights, bounds=bounds, eps=fd_step or SLSQP_EPS)
    x, result = run_slsqp(objective, p0, bounds, build_logsv_constraints(codec, constraints_type), disp=disp)
    fit = codec.parse(x)
    if return_info:
        return fit, {"fun": float(result.fun), "nit": int(result.nit), "nb_batches": objective.nb_batches, "x": x,
                     "nan_quotes": objective.nan_quotes}
    return fit


# ---- Heston -----------------------------------------------------------------------------------------------------------------
HESTON_BOUNDS = ((0.01, 2.0), (0.01, 2.0), (0.1, 30.0), (-0.99, 0.99), (0.1, 5.0))      # heston_pricer.py:127


def calibrate_heston(pricer, option_chain, params0, is_vega_weighted: bool, is_unit_ttm_vega: bool, disp: bool = False,
                     return_info: bool = False, fd_step: Optional[float] = None):
    from .heston_pricer import HestonParams, _params_c
    p0 = (np.array([params0.v0, params0.theta, params0.kappa, params0.rho, params0.volvol], dtype=float) if params0 is not None
          else np.array([0.1, 0.1, 2.0, -0.2, 1.0]))
    _, market_vols_ttms = option_chain.get_chain_data_as_xy()
    market_vols = to_flat_np_array(market_vols_ttms)
    weights = calibration_weights(option_chain, market_vols, is_vega_weighted, is_unit_ttm_vega)

    def batch_vols(points: np.ndarray) -> np.ndarray:
        sets = [_params_c(*p) for p in points]
        _, ivols = engine.heston_price_chain_batch(sets, option_chain.ttms, option_chain.forwards, option_chain.discfactors,
                                                   option_chain.strikes_ttms, option_chain.optiontypes_ttms)
        return ivols

    def feller(pars):                     # 2 kappa theta - volvol^2 >= 0 (heston_pricer.py:152-160)
        return 2.0 * pars[2] * pars[1] - pars[4] * pars[4]

    objective = BatchedObjective(batch_vols=batch_vols, market_vols=market_vols, weights=weights, bounds=HESTON_BOUNDS, eps=fd_step or SLSQP_EPS)
    x, result = run_slsqp(objective, p0, HESTON_BOUNDS, {"type": "ineq", "fun": feller}, disp=disp)
    fit = HestonParams(v0=x[0], theta=x[1], kappa=x[2], rho=x[3], volvol=x[4])
    if return_info:
        return fit, {"fun": float(result.fun), "nit": int(result.nit), "nb_batches": objective.nb_batches, "x": x}
    return fit
