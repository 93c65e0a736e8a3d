"""Affine-expansion log-MGF grid on the GPU: reference pricers/logsv/affine_expansion.py (default RK45 branch)."""
from __future__ import annotations

from enum import Enum
from typing import Optional, Tuple

import numpy as np

from ... import engine
from ... import _capi as C
from ...utils.config import VariableType


class ExpansionOrder(Enum):
    """truncation order (affine_expansion.py:43-54)."""
    ZERO = 0
    FIRST = 1
    SECOND = 2


def get_expansion_n(expansion_order: ExpansionOrder = ExpansionOrder.FIRST) -> int:
    return 3 if expansion_order == ExpansionOrder.FIRST else 5


def _order_code(expansion_order) -> int:
    v = expansion_order.value if hasattr(expansion_order, "value") else int(expansion_order)
    if v not in (1, 2):
        raise NotImplementedError           # affine_expansion.py:680-681
    return v


def func_a_ode_quadratic_terms(theta: float, kappa1: float, kappa2: float, beta: float, volvol: float, phi: complex, psi: complex,
                               is_spot_measure: bool = True, expansion_order: ExpansionOrder = ExpansionOrder.FIRST,
                               vol_backbone_eta: float = 1.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(M, L, H) of the coefficient ODEs A' = A^T M^(k) A + L A + H, Eqs. (4.17)/(4.25) (affine_expansion.py:67-184; same default order
    FIRST as the reference): dense complex tensors [n,n,n], [n,n], [n] written on the GPU from the row tables the ODE kernels integrate."""
    order = _order_code(expansion_order)
    M, L, H = engine.logsv_ode_terms(np.array([phi]), np.array([psi]), engine.logsv_params_c(theta, theta, kappa1, kappa2, beta, volvol),
                                     vol_backbone_eta, is_spot_measure, order)
    return M[0], L[0], H[0]


def func_rhs(t: float, A0: np.ndarray, M, L: np.ndarray, H: np.ndarray) -> np.ndarray:
    """right-hand side of the coefficient ODE system, Eq. (4.14), with caller-supplied (M, L, H) (affine_expansion.py:187-205; ``t`` unused)."""
    return engine.ode_rhs_dense(A0, np.asarray(M), L, H)


def compute_logsv_a_mgf_grid(ttm: float, phi_grid: np.ndarray, psi_grid: np.ndarray, theta_grid: np.ndarray, sigma0: float,
                             theta: float, kappa1: float, kappa2: float, beta: float, volvol: float,
                             variable_type: VariableType = VariableType.LOG_RETURN,
                             expansion_order: ExpansionOrder = ExpansionOrder.SECOND, a_t0: Optional[np.ndarray] = None,
                             is_stiff_solver: bool = False, is_analytic: bool = False, is_spot_measure: bool = True,
                             vol_backbone_eta: float = 1.0, **kwargs) -> Tuple[np.ndarray, np.ndarray]:
    """(a_t1, log_mgf) over the transform grid, one ODE solve per grid point on the GPU (affine_expansion.py:570-685): the default RK45 branch
    with SciPy's control law (:492-529), or with ``is_analytic=True`` the semi-analytic branch (business-day steps, exact linear part,
    10 fixed-point sweeps for the quadratic part, :306-470; ``vol_backbone_eta`` is ignored there, as in the reference), or with
    ``is_stiff_solver=True`` SciPy's BDF control law with the analytic Jacobian (:229-303)."""
    order = _order_code(expansion_order)
    n = get_expansion_n(ExpansionOrder(order))
    if a_t0 is None:
        a_t0 = np.zeros((phi_grid.shape[0], n), dtype=np.complex128)
        if getattr(variable_type, "value", variable_type) == VariableType.SIGMA.value:   # by value: the reference's own enum duck-types
            a_t0[:, 1] = -theta_grid      # affine_expansion.py:562-564
    params = engine.logsv_params_c(sigma0, theta, kappa1, kappa2, beta, volvol)
    if is_analytic:            # takes precedence over is_stiff_solver, as in the reference (:643-654)
        return engine.logsv_mgf_grid_analytic(phi_grid, psi_grid, ttm, a_t0, params, is_spot_measure, order)
    if is_stiff_solver:
        return engine.logsv_mgf_grid_bdf(phi_grid, psi_grid, ttm, a_t0, params, vol_backbone_eta, is_spot_measure, order)
    return engine.logsv_mgf_grid(phi_grid, psi_grid, ttm, a_t0, params, vol_backbone_eta, is_spot_measure, order)


def solve_analytic_ode_grid_phi(phi_grid: np.ndarray, psi_grid: np.ndarray, ttm: float, theta: float, kappa1: float, kappa2: float, beta: float,
                                volvol: float, is_spot_measure: bool = True, a_t0: Optional[np.ndarray] = None,
                                expansion_order: ExpansionOrder = ExpansionOrder.FIRST, year_days: int = 260) -> np.ndarray:
    """A(ttm) over the grid by the semi-analytic scheme (affine_expansion.py:388-470 -> solve_analytic_ode_for_a :306-384)"""
    order = _order_code(expansion_order)
    if a_t0 is None:
        a_t0 = np.zeros((phi_grid.shape[0], get_expansion_n(ExpansionOrder(order))), dtype=np.complex128)
    a_t1, _ = engine.logsv_mgf_grid_analytic(phi_grid, psi_grid, ttm, a_t0, engine.logsv_params_c(theta, theta, kappa1, kappa2, beta, volvol),
                                             is_spot_measure, order, year_days)
    return a_t1


def get_init_conditions_a(phi_grid: np.ndarray, psi_grid: np.ndarray, theta_grid: np.ndarray, n_terms: int,
                          variable_type: VariableType = VariableType.LOG_RETURN) -> np.ndarray:
    """A(0) over the transform grid (affine_expansion.py:532-567): zeros, except that the volatility transform puts -Theta into the
    second coefficient."""
    vt = getattr(variable_type, "value", variable_type)
    if vt == VariableType.LOG_RETURN.value:
        return np.zeros((phi_grid.shape[0], n_terms), dtype=np.complex128)
    if vt == VariableType.Q_VAR.value:
        return np.zeros((psi_grid.shape[0], n_terms), dtype=np.complex128)
    if vt == VariableType.SIGMA.value:
        a_t0 = np.zeros((theta_grid.shape[0], n_terms), dtype=np.complex128)
        a_t0[:, 1] = -theta_grid
        return a_t0
    raise NotImplementedError


def solve_a_ode_grid(phi_grid: np.ndarray, psi_grid: np.ndarray, ttm: float, theta: float, kappa1: float, kappa2: float, beta: float,
                     volvol: float, is_spot_measure: bool = True, a_t0: Optional[np.ndarray] = None, is_stiff_solver: bool = False,
                     expansion_order: ExpansionOrder = ExpansionOrder.FIRST, vol_backbone_eta: float = 1.0) -> np.ndarray:
    """A(ttm) for every grid point from A(0) = ``a_t0`` (affine_expansion.py:492-529: a Python loop of ``solve_ivp`` calls there, one
    GPU launch here; same default ``expansion_order`` = FIRST as the reference function)."""
    order = _order_code(expansion_order)
    if a_t0 is None:
        a_t0 = np.zeros((phi_grid.shape[0], get_expansion_n(ExpansionOrder(order))), dtype=np.complex128)
    # the log-MGF contraction needs sigma0 - theta; it is discarded here, so any finite sigma0 does
    solver = engine.logsv_mgf_grid_bdf if is_stiff_solver else engine.logsv_mgf_grid
    a_t1, _ = solver(phi_grid, psi_grid, ttm, a_t0, engine.logsv_params_c(theta, theta, kappa1, kappa2, beta, volvol), vol_backbone_eta,
                     is_spot_measure, order)
    return a_t1


# ---- single-point entry points of the reference (affine_expansion.py:209-384): one-element grids through the same kernels ---------------
def func_rhs_jac(t: float, A0: np.ndarray, M, L: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Jacobian of func_rhs in A, d/dA (A^T M^(k) A) = 2 M^(k) A for the symmetric M^(k), plus L (affine_expansion.py:209-225; the argument
    order is solve_ivp's, ``t`` and ``H`` are unused).  Interface helper for callers that drive their own stiff solver: n <= 5 complex
    entries on the host -- the BDF kernel evaluates its Jacobian in registers from the row tables."""
    return 2.0 * np.einsum("kij,j->ki", np.asarray(M, dtype=np.complex128), np.asarray(A0, dtype=np.complex128)) + np.asarray(L)


def solve_ode_for_a(ttm: float, theta: float, kappa1: float, kappa2: float, beta: float, volvol: float, phi: complex, psi: complex,
                    is_spot_measure: bool = True, a_t0: Optional[np.ndarray] = None, expansion_order: ExpansionOrder = ExpansionOrder.FIRST,
                    is_stiff_solver: bool = False, dense_output: bool = False, vol_backbone_eta: float = 1.0):
    """coefficient ODEs of Eq. (4.14) for ONE transform point (affine_expansion.py:229-303).  Returns a SciPy ``OdeResult`` holding the two
    ends of the integration (``t = [0, ttm]``, ``y[:, -1] = A(ttm)`` as the reference's callers read it); the accepted intermediate steps stay
    on the device and a continuous extension is not built, so ``dense_output=True`` raises."""
    if dense_output:
        raise NotImplementedError("dense_output: the GPU solver returns A(ttm) only")
    from scipy.integrate._ivp.ivp import OdeResult
    n = get_expansion_n(ExpansionOrder(_order_code(expansion_order)))
    y0 = np.zeros(n, dtype=np.complex128) if a_t0 is None else np.asarray(a_t0, dtype=np.complex128).reshape(n)
    a_t1 = solve_a_ode_grid(np.array([phi], dtype=np.complex128), np.array([psi], dtype=np.complex128), ttm, theta, kappa1, kappa2, beta, volvol,
                            is_spot_measure=is_spot_measure, a_t0=y0[None, :], is_stiff_solver=is_stiff_solver, expansion_order=expansion_order,
                            vol_backbone_eta=vol_backbone_eta)[0]
    return OdeResult(t=np.array([0.0, float(ttm)]), y=np.stack([y0, a_t1], axis=1), sol=None, t_events=None, y_events=None, nfev=0, njev=0, nlu=0,
                     status=0, message="The solver successfully reached the end of the integration interval.", success=True)


def solve_analytic_ode_for_a(ttm: float, theta: float, kappa1: float, kappa2: float, beta: float, volvol: float, phi: complex, psi: complex,
                             is_spot_measure: bool, a_t0: Optional[np.ndarray] = None, expansion_order: ExpansionOrder = ExpansionOrder.FIRST,
                             year_days: int = 260) -> np.ndarray:
    """A(ttm) for ONE transform point by the semi-analytic scheme (affine_expansion.py:306-384)"""
    a0 = None if a_t0 is None else np.asarray(a_t0, dtype=np.complex128).reshape(1, -1)
    return solve_analytic_ode_grid_phi(np.array([phi], dtype=np.complex128), np.array([psi], dtype=np.complex128), ttm, theta, kappa1, kappa2, beta,
                                       volvol, is_spot_measure=is_spot_measure, a_t0=a0, expansion_order=expansion_order, year_days=year_days)[0]
