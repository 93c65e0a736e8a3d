"""HawkesJDPricer on B200: the reference's ``pricers/hawkes_jd_pricer.py``, Monte Carlo and Fourier routes.

* ``HawkesJDParams``                          (hawkes_jd_pricer.py:41-119)
* ``HawkesJDPricer.model_mc_price_chain``     (:156-171) -> ``hawkesjd_mc_chain_pricer`` (:644-715)
* ``HawkesJDPricer.simulate_terminal_values`` (:193-224) -> ``simulate_hawkesjd_terminal`` (:718-779)
* ``HawkesJDPricer.price_chain``              (:125-155) -> ``hawkesjd_chain_pricer`` (:365-417) / ``hawkesjd_chain_pricer_with_risk_premia`` (:420-484),
  ``hawkesjd_forwards_under_risk_kernel`` (:487-515), ``compute_hawkes_a_mgf_grid`` / ``solve_a_ode_grid`` (:518-579)
* ``HawkesJDPricer.calibrate_model_params_to_chain`` (:230-300), ``calibrate_risk_premia_gamma_to_chain`` (:303-357): host SLSQP loops around the
  GPU chain pricer

Monte Carlo: the reference draws every random input from numpy's process-global generator; here they are drawn in-kernel from the Philox
stream keyed by an explicit ``seed`` (default: the ``set_seed`` stream or OS entropy).  ``simulate_hawkesjd_terminal`` additionally accepts the five
input blocks (``W0, U_P, U_M, J_P, J_M`` in the reference's own form) and then reproduces the reference arithmetic operation by operation.
Fourier: the 3-equation Riccati system per transform point goes through the same SciPy-RK45 clone as the LogSV coefficient ODEs
(csrc/mgf_kernels.cu, ``hawkes_mgf_kernel``); its stiff (BDF) variant is not built and raises.
"""
from __future__ import annotations

from ctypes import byref
from dataclasses import asdict, dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .. import _capi as C
from .. import engine
from ..data.option_chain import OptionChain
from ..utils.config import VariableType
from ..utils.funcs import set_time_grid, timer
from .model_pricer import ModelParams, ModelPricer

_KEYS = tuple(k for k, _ in C.HawkesParamsC._fields_)
STEPS_PER_YEAR = 5 * 360          # hawkes_jd_pricer.py:752 "need small dt step for large intensities"


@dataclass
class HawkesJDParams(ModelParams):
    """parameters of the 2-factor Hawkes jump-diffusion, annualised (reference :41-65, same defaults)"""
    mu: float = 0.0
    sigma: float = 0.45
    shift_p: float = 0.06
    mean_p: float = 0.03
    shift_m: float = -0.06
    mean_m: float = -0.03
    lambda_p: float = 6.55
    theta_p: float = 6.55
    kappa_p: float = 22.29
    beta1_p: float = 76.0
    beta2_p: float = -67.58
    lambda_m: float = 8.50
    theta_m: float = 8.50
    kappa_m: float = 29.0
    beta1_m: float = 104.55
    beta2_m: float = -109.6
    risk_premia_gamma: Optional[float] = None

    def __post_init__(self):
        self.compensator_p = np.exp(self.shift_p) / (1.0 - self.mean_p) - 1.0
        self.compensator_m = np.exp(self.shift_m) / (1.0 - self.mean_m) - 1.0

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)

    @property
    def exp_jump_p(self) -> float:
        return self.shift_p + self.mean_p

    @property
    def exp_jump_m(self) -> float:
        return self.shift_m + self.mean_m

    @property
    def jump1_cond(self) -> float:
        """stationarity margin of the positive-jump intensity (:87-96)"""
        return self.kappa_p - self.beta1_p * self.exp_jump_p - self.beta2_p * self.exp_jump_m

    @property
    def jump2_cond(self) -> float:
        return self.kappa_m - self.beta2_m * self.exp_jump_m - self.beta1_m * self.exp_jump_p


def _params_c(**kw) -> C.HawkesParamsC:
    return C.HawkesParamsC(*[float(kw[k]) for k in _KEYS])


class HawkesJDPricer(ModelPricer):
    """ModelPricer for the Hawkes jump-diffusion model: Fourier and Monte Carlo routes on the GPU."""

    def price_chain(self, option_chain: OptionChain, params: HawkesJDParams, is_spot_measure: bool = True, **kwargs) -> List[np.ndarray]:
        """Fourier prices of the chain (reference :125-155): the risk-kernel pricer when ``params.risk_premia_gamma`` is set, else the plain one"""
        pricer = hawkesjd_chain_pricer if params.risk_premia_gamma is None else hawkesjd_chain_pricer_with_risk_premia
        return pricer(model_params=params, ttms=option_chain.ttms, forwards=option_chain.forwards, discfactors=option_chain.discfactors,
                      strikes_ttms=option_chain.strikes_ttms, optiontypes_ttms=option_chain.optiontypes_ttms, is_spot_measure=is_spot_measure, **kwargs)

    @timer
    def model_mc_price_chain(self, option_chain: OptionChain, params: HawkesJDParams, nb_path: int = 100000, **kwargs
                             ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        d = params.to_dict()
        d.pop("risk_premia_gamma", None)
        return hawkesjd_mc_chain_pricer(ttms=option_chain.ttms, forwards=option_chain.forwards, discfactors=option_chain.discfactors,
                                        strikes_ttms=option_chain.strikes_ttms, optiontypes_ttms=option_chain.optiontypes_ttms, nb_path=nb_path,
                                        seed=kwargs.get("seed"), gauss=kwargs.get("gauss", "fp32"),
                                        variable_type=kwargs.get("variable_type", VariableType.LOG_RETURN),
                                        distributed=kwargs.get("distributed", True), exchange=kwargs.get("exchange"), **d)

    # ---- calibration drivers (reference :230-357): SLSQP on the host around the GPU Fourier chain pricer -----------------------------
    def _vol_objective(self, option_chain: OptionChain, unpack, is_vega_weighted: bool, is_unit_ttm_vega: bool, scale: float = 1.0):
        """weighted squared difference of model and mid implied vols (nansum: quotes whose model price leaves the no-arbitrage bounds drop out)"""
        market = np.concatenate([np.asarray(v, dtype=float) for v in option_chain.get_mid_vols()])
        if is_vega_weighted:
            weights = np.concatenate([v / np.sum(v) for v in option_chain.get_chain_vegas(is_unit_ttm_vega=is_unit_ttm_vega)])
        else:
            weights = np.ones_like(market)

        def objective(pars: np.ndarray, *_) -> float:
            model = np.concatenate(self.compute_model_ivols_for_chain(option_chain=option_chain, params=unpack(pars)))
            return float(np.nansum(scale * weights * np.square(model - market)))
        return objective

    @timer
    def calibrate_model_params_to_chain(self, option_chain: OptionChain, params0: HawkesJDParams, is_vega_weighted: bool = True,
                                        is_unit_ttm_vega: bool = False, **kwargs) -> HawkesJDParams:
        """fit (sigma, mean_p, mean_m, theta_p, theta_m, kappa, beta_p, beta_m) to the chain's mid vols with shifts and initial intensities
        fixed at ``params0``'s; kappa is shared by the two intensities and beta_p / beta_m enter as +beta / -beta pairs (reference :230-300, same
        start vector -- including its beta_m start 0.5 (beta2_p - beta2_m) --, bounds, stationarity constraint and SLSQP options).
        ``disp`` / ``maxiter`` (extras) are handed to SLSQP."""
        from scipy.optimize import minimize
        p0 = np.array([params0.sigma, params0.mean_p, params0.mean_m, params0.theta_p, params0.theta_m, 0.5 * (params0.kappa_p + params0.kappa_m),
                       0.5 * (params0.beta1_p - params0.beta2_p), 0.5 * (params0.beta2_p - params0.beta2_m)])
        bounds = ((0.10, 2.0), (0.01, 0.99), (-0.99, -0.01), (0.01, 100.0), (0.01, 100.0), (1.0, 100.0), (1.0, 100.0), (1.0, 100.0))

        def unpack(x: np.ndarray) -> HawkesJDParams:
            return HawkesJDParams(mu=0.0, sigma=x[0], shift_p=params0.shift_p, mean_p=x[1], shift_m=params0.shift_m, mean_m=x[2],
                                  lambda_p=params0.lambda_p, theta_p=x[3], kappa_p=x[5], beta1_p=x[6], beta2_p=-x[6], lambda_m=params0.lambda_m,
                                  theta_m=x[4], kappa_m=x[5], beta1_m=x[7], beta2_m=-x[7])

        stationary = {"type": "ineq", "fun": lambda x: unpack(x).jump1_cond + unpack(x).jump2_cond}
        options = {"disp": bool(kwargs.get("disp", False)), "ftol": 1e-8}
        if "maxiter" in kwargs:
            options["maxiter"] = int(kwargs["maxiter"])
        res = minimize(self._vol_objective(option_chain, unpack, is_vega_weighted, is_unit_ttm_vega), p0, args=None, method="SLSQP",
                       constraints=stationary, bounds=bounds, options=options)
        return unpack(res.x)

    @timer
    def calibrate_risk_premia_gamma_to_chain(self, option_chain: OptionChain, params0: HawkesJDParams, is_vega_weighted: bool = True,
                                             is_unit_ttm_vega: bool = False, maxiter: int = 100, print_iter: bool = True, **kwargs) -> HawkesJDParams:
        """fit (sigma, risk_premia_gamma) with every other parameter of ``params0`` fixed (reference :303-357: gamma scaled by 8 into (-1, 1),
        weights x 1e4, SLSQP with finite-difference step 0.025, ftol = tol = 1e-16).  Returns a new object; the reference mutates ``params0``."""
        from scipy.optimize import minimize
        gamma_scaler = 8.0
        fixed = {k: v for k, v in params0.to_dict().items() if k in _KEYS}

        def unpack(x: np.ndarray) -> HawkesJDParams:
            if print_iter:
                print(f"unpack_pars: sigma={x[0]}, gamma={gamma_scaler * x[1]}")
            return HawkesJDParams(**{**fixed, "sigma": float(x[0])}, risk_premia_gamma=float(gamma_scaler * x[1]))

        p0 = np.array([params0.sigma, (params0.risk_premia_gamma or 0.0) / gamma_scaler])
        res = minimize(self._vol_objective(option_chain, unpack, is_vega_weighted, is_unit_ttm_vega, scale=10000.0), p0, args=None, method="SLSQP",
                       bounds=((0.01, 1.5), (-1.0, 1.0)), options={"disp": bool(kwargs.get("disp", False)), "ftol": 1e-16, "maxiter": maxiter, "eps": 0.025},
                       tol=1e-16)
        return unpack(res.x)

    @timer
    def simulate_terminal_values(self, params: HawkesJDParams, ttm: float = 1.0, nb_path: int = 100000, is_spot_measure: bool = True, **kwargs
                                 ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        d = {k: v for k, v in params.to_dict().items() if k in _KEYS and k not in ("lambda_p", "lambda_m")}
        return simulate_hawkesjd_terminal(ttm=ttm, x0=np.zeros(1), lambda_p0=params.lambda_p * np.ones(1), lambda_m0=params.lambda_m * np.ones(1),
                                          nb_path=nb_path, seed=kwargs.get("seed"), gauss=kwargs.get("gauss", "fp32"), **d)


def hawkesjd_mc_chain_pricer(ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms, lambda_p: float, lambda_m: float, mu: float, sigma: float,
                             shift_p: float, mean_p: float, shift_m: float, mean_m: float, theta_p: float, kappa_p: float, beta1_p: float,
                             beta2_p: float, theta_m: float, kappa_m: float, beta1_m: float, beta2_m: float, risk_premia_gamma: float = 0.0,
                             nb_path: int = 100000, variable_type: VariableType = VariableType.LOG_RETURN, seed: Optional[int] = None,
                             gauss: str = "fp32", distributed: bool = True, exchange: Optional[str] = None
                             ) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """chain prices and standard errors by simulating the jump-diffusion (reference :644-715): the terminal (x, lambda_p, lambda_m) of slice m
    seeds slice m+1; payoffs are forward-recentred on x (utils/mc_payoffs.py).  Under an initialised torch.distributed world ``nb_path`` is
    the TOTAL path count, sharded over the ranks by global path id (the LogSV chain's two exchanges per maturity, multi_gpu.py)."""
    vt = engine.variable_code(variable_type)
    pc = _params_c(**locals())
    from .logsv_pricer import _shared_seed, _use_distributed
    seed = _shared_seed(seed)
    from .logsv_pricer import _use_distributed
    if _use_distributed({"distributed": distributed, "nb_path": nb_path, "exchange": exchange}):
        from ..multi_gpu import mc_chain_distributed
        C.encode_types(np.concatenate([np.asarray(t) for t in optiontypes_ttms]))
        return mc_chain_distributed("hawkes", pc, ttms, forwards, discfactors, None, strikes_ttms, optiontypes_ttms, nb_path, STEPS_PER_YEAR,
                                    True, vt, seed, engine.mc_flags("fp64", gauss), exchange=exchange)
    M, ttms, forwards, discfactors, offsets, strikes, types = engine._chain_arrays(ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms)
    prices, stds = np.empty(strikes.shape[0]), np.empty(strikes.shape[0])
    C.call("b200sv_hawkesjd_mc_chain", byref(pc), M, C.dptr(ttms), C.dptr(forwards), C.dptr(discfactors), C.iptr(offsets), C.dptr(strikes),
           C.i8ptr(types), int(nb_path), vt, seed & 0xFFFFFFFFFFFFFFFF, engine.mc_flags("fp64", gauss), C.dptr(prices), C.dptr(stds))
    return C.split_chain(prices, offsets), C.split_chain(stds, offsets)


def simulate_hawkesjd_terminal(ttm: float, x0: np.ndarray, lambda_p0: np.ndarray, lambda_m0: np.ndarray, mu: float, sigma: float, shift_p: float,
                               mean_p: float, shift_m: float, mean_m: float, theta_p: float, kappa_p: float, beta1_p: float, beta2_p: float,
                               theta_m: float, kappa_m: float, beta1_m: float, beta2_m: float, nb_path: int = 100000, seed: Optional[int] = None,
                               gauss: str = "fp32", slice_index: int = 0, inputs=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """terminal (log-return, lambda_p, lambda_m) with both Hawkes intensities evolving jointly (reference :718-779).  Length-1 initial
    arrays broadcast as in the reference (:742-750: x0 -> zeros, intensities -> constants).  ``inputs = (W0, U_P, U_M, J_P, J_M)`` (extra;
    [nb_steps, nb_path] each, in the form :753-757 builds them) runs the strict kernel on caller-supplied draws."""
    x0, lambda_p0, lambda_m0 = np.atleast_1d(x0), np.atleast_1d(lambda_p0), np.atleast_1d(lambda_m0)
    for a in (x0, lambda_p0, lambda_m0):
        assert a.shape[0] in (1, nb_path)
    x = np.zeros(nb_path) if x0.shape[0] == 1 else np.array(x0, dtype=np.float64)
    lp = lambda_p0 * np.ones(nb_path) if lambda_p0.shape[0] == 1 else np.array(lambda_p0, dtype=np.float64)
    lm = lambda_m0 * np.ones(nb_path) if lambda_m0.shape[0] == 1 else np.array(lambda_m0, dtype=np.float64)
    x, lp, lm = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, lp, lm))
    pc = _params_c(lambda_p=float(lp[0]), lambda_m=float(lm[0]), **{k: v for k, v in locals().items() if k in _KEYS})
    nb_steps, dt, _ = set_time_grid(ttm=ttm, nb_steps_per_year=STEPS_PER_YEAR)
    if inputs is not None:
        blocks = [C.f64(b) for b in inputs]
        if len(blocks) != 5 or any(b.shape != (nb_steps, nb_path) for b in blocks):
            raise ValueError(f"inputs must be five arrays of shape ({nb_steps}, {nb_path})")
        C.call("b200sv_hawkesjd_step_fixed", C.dptr(x), C.dptr(lp), C.dptr(lm), *[C.dptr(b) for b in blocks], nb_steps, nb_path, dt, byref(pc))
        return x, lp, lm
    const_start = x0.shape[0] == 1 and lambda_p0.shape[0] == 1 and lambda_m0.shape[0] == 1
    C.call("b200sv_hawkesjd_terminal", byref(pc), float(ttm), int(nb_path), (engine.fresh_seed() if seed is None else int(seed)) & 0xFFFFFFFFFFFFFFFF,
           engine.mc_flags("fp64", gauss), int(slice_index), 0 if const_start else 1, C.dptr(x), C.dptr(lp), C.dptr(lm))
    return x, lp, lm


def hawkesjd_device_draws(seed: int, path0: int, n: int, slice_index: int, ttm: float, gauss: str = "fp32", **params):
    """(W0, U_P, U_M, J_P, J_M, dt): what the in-kernel generator draws for paths [path0, path0 + n) of a slice of length ``ttm``, in the
    reference's form (test / diagnostics hook)."""
    pc = _params_c(**{**{k: 0.0 for k in _KEYS}, **params})
    S, dt, _ = set_time_grid(ttm=ttm, nb_steps_per_year=STEPS_PER_YEAR)
    out = [np.empty((S, n)) for _ in range(5)]
    C.call("b200sv_hawkesjd_device_draws", int(seed) & 0xFFFFFFFFFFFFFFFF, int(path0), int(n), int(slice_index), S, dt, byref(pc),
           engine.mc_flags("fp64", gauss), *[C.dptr(o) for o in out])
    return (*out, dt)


MAX_PHI = 500           # transform grid size of the Fourier route (reference :37)


def set_vol_scaler(sigma0: float, ttm: float) -> float:
    """grid scaler of the Fourier route (reference :360-362)"""
    return float(np.clip(sigma0, 0.2, 0.5) * np.sqrt(np.minimum(ttm, 1.0 / 12.0)))


def _fourier_chain(model_params: HawkesJDParams, ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms, is_stiff_solver, is_spot_measure,
                   variable_type, vol_scaler, gamma, return_grids=False):
    if getattr(variable_type, "value", variable_type) != VariableType.LOG_RETURN.value:
        raise NotImplementedError           # reference :413-414
    if is_stiff_solver:
        raise NotImplementedError("is_stiff_solver: the Hawkes Fourier route runs SciPy's RK45 control law on the GPU; its BDF variant is not built")
    if gamma is None:
        engine._check_fourier_types(optiontypes_ttms, bool(is_spot_measure))
    elif not is_spot_measure or any(str(t) not in ("C", "P") for types in optiontypes_ttms for t in types):
        raise ValueError("not implemented")           # utils/mgf_pricer.py:310-318
    M, ttms, forwards, discfactors, offsets, strikes, types = engine._chain_arrays(ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms)
    d = model_params.to_dict()
    d.pop("risk_premia_gamma", None)
    pc = _params_c(**d)
    prices = np.empty(strikes.shape[0])
    a = np.empty((M, MAX_PHI, 3), dtype=np.complex128) if return_grids else None
    lm = np.empty((M, MAX_PHI), dtype=np.complex128) if return_grids else None
    norm, gfw = np.empty(M), np.empty(M)
    cptr = lambda arr: None if arr is None else arr.ctypes.data_as(C._dp)
    C.call("b200sv_hawkesjd_price_chain", byref(pc), M, C.dptr(ttms), C.dptr(forwards), C.dptr(discfactors), C.iptr(offsets), C.dptr(strikes),
           C.i8ptr(types), int(bool(is_spot_measure)), float(vol_scaler) if vol_scaler is not None else 0.0, MAX_PHI,
           float("nan") if gamma is None else float(gamma), C.dptr(prices), cptr(a), cptr(lm), C.dptr(norm), C.dptr(gfw))
    out = C.split_chain(prices, offsets)
    return (out, a, lm, norm, gfw) if return_grids else out


def hawkesjd_chain_pricer(model_params: HawkesJDParams, ttms: np.ndarray, forwards: np.ndarray, discfactors: np.ndarray, strikes_ttms, optiontypes_ttms,
                          is_stiff_solver: bool = False, is_spot_measure: bool = True, variable_type: VariableType = VariableType.LOG_RETURN,
                          vol_scaler: float = None, return_grids: bool = False) -> List[np.ndarray]:
    """Fourier chain prices of the Hawkes jump-diffusion (reference :365-417): Riccati ODEs for (a0, a_p, a_m) on a 500-point transform grid
    with SciPy's RK45 control law per point, carried over the maturities, then the vanilla slice pricer -- one GPU call"""
    return _fourier_chain(model_params, ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms, is_stiff_solver, is_spot_measure, variable_type,
                          vol_scaler, None, return_grids)


def hawkesjd_chain_pricer_with_risk_premia(model_params: HawkesJDParams, ttms: np.ndarray, forwards: np.ndarray, discfactors: np.ndarray, strikes_ttms,
                                           optiontypes_ttms, is_stiff_solver: bool = False, is_spot_measure: bool = True,
                                           variable_type: VariableType = VariableType.LOG_RETURN, vol_scaler: float = None,
                                           return_grids: bool = False) -> List[np.ndarray]:
    """Fourier chain prices under the risk kernel exp(-gamma x) (reference :420-484; grid on Re = -1/2 - gamma, normalisers and forwards under the
    kernel from two single-point solves per maturity, gamma slice pricer without discounting -- all as in the reference)"""
    return _fourier_chain(model_params, ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms, is_stiff_solver, is_spot_measure, variable_type,
                          vol_scaler, model_params.risk_premia_gamma, return_grids)


def hawkesjd_forwards_under_risk_kernel(model_params: HawkesJDParams, risk_premia_gamma: float, ttms: np.ndarray, forwards: np.ndarray,
                                        is_stiff_solver: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """(normalizers, gamma_forwards) per maturity (reference :487-515)"""
    ttms, forwards = np.asarray(ttms, dtype=float), np.asarray(forwards, dtype=float)
    one = [np.ones(1) for _ in ttms]
    p = HawkesJDParams(**{**model_params.to_dict(), "risk_premia_gamma": float(risk_premia_gamma)})
    _, _, _, norm, gfw = _fourier_chain(p, ttms, forwards, np.ones_like(ttms), one, [np.array(["C"]) for _ in ttms], is_stiff_solver, True,
                                        VariableType.LOG_RETURN, None, float(risk_premia_gamma), return_grids=True)
    return norm, gfw


def compute_hawkes_a_mgf_grid(ttm: float, phi_grid: np.ndarray, model_params: HawkesJDParams, psi_grid: Optional[np.ndarray] = None,
                              a_t0: Optional[np.ndarray] = None, is_stiff_solver: bool = False, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
    """(a_t1 [P, 3], log_mgf [P]) over a caller-supplied transform grid (reference :518-547): A(ttm) from A(0) = ``a_t0`` (zeros by default) and
    log-MGF = a0 + a_p lambda_p + a_m lambda_m"""
    if is_stiff_solver:
        raise NotImplementedError("is_stiff_solver: the Hawkes Fourier route runs SciPy's RK45 control law on the GPU; its BDF variant is not built")
    phi = C.c128(phi_grid)
    psi = None if psi_grid is None else C.c128(psi_grid)
    a = np.zeros((phi.shape[0], 3), dtype=np.complex128) if a_t0 is None else np.ascontiguousarray(a_t0, dtype=np.complex128).copy()
    if a.shape != (phi.shape[0], 3):
        raise ValueError("a_t0 must have shape (len(phi_grid), 3)")
    d = model_params.to_dict()
    d.pop("risk_premia_gamma", None)
    lm = np.empty(phi.shape[0], dtype=np.complex128)
    cp = lambda arr: None if arr is None else arr.ctypes.data_as(C._dp)
    C.call("b200sv_hawkesjd_mgf_grid", cp(phi), cp(psi), phi.shape[0], float(ttm), cp(a), byref(_params_c(**d)), cp(lm))
    return a, lm


def solve_a_ode_grid(phi_grid: np.ndarray, ttm: float, model_params: HawkesJDParams, psi_grid: Optional[np.ndarray] = None,
                     a_t0: Optional[np.ndarray] = None, is_stiff_solver: bool = False) -> np.ndarray:
    """A(ttm) per transform point (reference :550-579)"""
    return compute_hawkes_a_mgf_grid(ttm, phi_grid, model_params, psi_grid=psi_grid, a_t0=a_t0, is_stiff_solver=is_stiff_solver)[0]
