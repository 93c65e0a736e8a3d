"""oracle.mgf -- numpy restatement of the Fourier / MGF hot path.  TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``); pinned against ``tests/golden/logsv_fourier_*.npz``,
``heston_fourier_*.npz``, ``mlh.npz``, ``grids.npz``, ``fourier_sum_*.npz``.

All ``file:line`` citations are relative to ``/root/reference/src/stochvolmodels`` unless they
start with ``scipy/``.
"""
from __future__ import annotations

import numpy as np

TYPE_CODES = {"C": 0, "P": 1, "IC": 2, "IP": 3}


# --------------------------------------------------------------------------------------------
# transform grid and quadrature weights
# --------------------------------------------------------------------------------------------
def logsv_vol_scaler(sigma0: float, ttms) -> float:
    """``set_vol_scaler`` (pricers/logsv_pricer.py:664-666): sigma0*sqrt(min(min ttm, 0.5/12))."""
    return sigma0 * np.sqrt(np.minimum(np.min(ttms), 0.5 / 12.0))


def heston_vol_scaler(v0: float, ttms) -> float:
    """pricers/heston_pricer.py:234-235: min(0.3, sqrt(v0*ttms[0]))."""
    return np.minimum(0.3, np.sqrt(v0 * ttms[0]))


def phi_grid(vol_scaler: float, is_spot_measure: bool = True, max_phi: int = 1000) -> np.ndarray:
    """``get_phi_grid`` (utils/mgf_pricer.py:11-34): Re = -1/2 (MMA) | +1/2 (inverse); Im = linspace(0, 5.6/vs, P)."""
    p = np.linspace(0, 5.6 / vol_scaler, max_phi)
    return (-0.5 if is_spot_measure else 0.5) + 1j * p


def legacy_simpson_weights(phi: np.ndarray) -> np.ndarray:
    """``_compute_legacy_pricer_weights`` (utils/mgf_pricer.py:157-171).

    [1, 4, 2, 4, ...]*h/3 with first and last set to 1 and THEN every odd index to 4, so that on the
    even-length (P=1000) production grid the last weight is 4*h/3.
    """
    p = np.imag(phi)
    w = np.full(p.shape[0], 2.0)
    w[0] = 1.0
    w[-1] = 1.0
    w[1::2] = 4.0
    return ((p[1] - p[0]) / 3.0) * w


# --------------------------------------------------------------------------------------------
# affine expansion: coefficient terms and right-hand side
# --------------------------------------------------------------------------------------------
def expansion_n(order: int) -> int:
    """``get_expansion_n`` (pricers/logsv/affine_expansion.py:57-64): 3 for FIRST (1), else 5."""
    return 3 if order == 1 else 5


def logsv_mlh(theta, kappa1, kappa2, beta, volvol, phi, psi, is_spot_measure=True, order=2, eta=1.0):
    """Dense M[n,n,n], L[n,n], H[n] of ``func_a_ode_quadratic_terms`` (affine_expansion.py:67-184).

    Written from the sparse term list (the same list the CUDA kernel hard-codes): every off-diagonal
    M entry appears twice (M is symmetric in its last two indices).
    """
    n = expansion_n(order)
    th2 = theta * theta
    v2 = beta * beta + volvol * volvol
    qv = theta * v2
    qv2 = th2 * v2
    eta2 = eta * eta
    if is_spot_measure:          # :127-134
        lam, k2p, kp = 0.0, kappa2, kappa1 + kappa2 * theta
    else:
        lam = beta * th2 * eta
        k2p = kappa2 - beta * eta
        kp = kappa1 + kappa2 * theta - 2 * beta * theta * eta
    M = np.zeros((n, n, n), dtype=np.complex128)
    sym = []   # (k, i, j, value) with i<j stored on both sides
    diag = [(0, 1, 0.5 * qv2), (1, 1, qv), (2, 1, 0.5 * v2), (2, 2, 2.0 * qv2)]
    sym += [(1, 1, 2, qv2), (2, 1, 2, 2.0 * qv)]
    if order == 2:               # :152-164
        diag += [(3, 2, 4.0 * qv), (4, 2, 2.0 * v2), (4, 3, 4.5 * qv2)]
        sym += [(2, 1, 3, 1.5 * qv2), (3, 1, 2, v2), (3, 1, 3, 3.0 * qv), (3, 1, 4, 2.0 * qv2), (3, 2, 3, 3.0 * qv2),
                (4, 1, 3, 1.5 * v2), (4, 1, 4, 4.0 * qv), (4, 2, 3, 6.0 * qv), (4, 2, 4, 4.0 * qv2)]
    for k, i, val in diag:
        M[k, i, i] = val
    for k, i, j, val in sym:
        M[k, i, j] = val
        M[k, j, i] = val
    bphi = beta * eta * phi
    L = np.zeros((n, n), dtype=np.complex128)
    L[0, 1], L[0, 2] = lam - th2 * bphi, qv2                                             # :168
    L[1, 1], L[1, 2] = -kp - 2.0 * theta * bphi, 2.0 * (lam + qv - th2 * bphi)           # :169
    L[2, 1], L[2, 2] = -k2p - bphi, v2 - 2.0 * kp - 4.0 * theta * bphi                   # :170
    if order == 2:                                                                        # :172-176
        L[1, 3] = 3.0 * qv2
        L[2, 3], L[2, 4] = 3.0 * (2.0 * qv - th2 * bphi), 6.0 * qv2
        L[3, 2], L[3, 3], L[3, 4] = -2.0 * (k2p + bphi), 3.0 * (v2 - kp - 2.0 * theta * bphi), 4.0 * (3.0 * qv - th2 * bphi)
        L[4, 3], L[4, 4] = -3.0 * (k2p + bphi), 2.0 * (v2 - 2.0 * kp - 4.0 * theta * bphi)
    rhs = phi * (phi + 1.0) - 2.0 * psi if is_spot_measure else phi * (phi - 1.0) - 2.0 * psi   # :180-183
    H = np.zeros(n, dtype=np.complex128)
    H[0], H[1], H[2] = 0.5 * th2 * eta2 * rhs, theta * eta2 * rhs, 0.5 * eta2 * rhs
    return M, L, H


def rhs_dense(A, M, L, H):
    """``func_rhs`` (affine_expansion.py:187-205): A' M^(k) A + L A + H for one grid point."""
    quad = np.array([A @ M[k] @ A for k in range(A.shape[0])])
    return quad + L @ A + H


class LogsvRhs:
    """Vectorised (over grid points) right-hand side using the expanded sparse form.

    A has shape (P, n).  Mathematically identical to :func:`rhs_dense`; term order differs, which the
    reference itself tolerates (``fastmath=True`` + BLAS order, tests at 1e-14..1e-15).
    """

    def __init__(self, theta, kappa1, kappa2, beta, volvol, phi, psi, is_spot_measure=True, order=2, eta=1.0):
        self.n = expansion_n(order)
        self.order = order
        th2 = theta * theta
        self.v2 = beta * beta + volvol * volvol
        self.qv = theta * self.v2
        self.qv2 = th2 * self.v2
        eta2 = eta * eta
        if is_spot_measure:
            lam, k2p, kp = 0.0, kappa2, kappa1 + kappa2 * theta
        else:
            lam = beta * th2 * eta
            k2p = kappa2 - beta * eta
            kp = kappa1 + kappa2 * theta - 2 * beta * theta * eta
        b = beta * eta * phi
        v2, qv, qv2 = self.v2, self.qv, self.qv2
        self.l01 = lam - th2 * b
        self.l11 = -kp - 2.0 * theta * b
        self.l12 = 2.0 * (lam + qv - th2 * b)
        self.l21 = -k2p - b
        self.l22 = v2 - 2.0 * kp - 4.0 * theta * b
        self.l23 = 3.0 * (2.0 * qv - th2 * b)
        self.l32 = -2.0 * (k2p + b)
        self.l33 = 3.0 * (v2 - kp - 2.0 * theta * b)
        self.l34 = 4.0 * (3.0 * qv - th2 * b)
        self.l43 = -3.0 * (k2p + b)
        self.l44 = 2.0 * (v2 - 2.0 * kp - 4.0 * theta * b)
        r = phi * (phi + 1.0) - 2.0 * psi if is_spot_measure else phi * (phi - 1.0) - 2.0 * psi
        self.h0, self.h1, self.h2 = 0.5 * th2 * eta2 * r, theta * eta2 * r, 0.5 * eta2 * r

    def __call__(self, A, idx=slice(None)):
        v2, qv, qv2 = self.v2, self.qv, self.qv2
        g = lambda a: a[idx] if isinstance(a, np.ndarray) else a
        a1, a2 = A[:, 1], A[:, 2]
        out = np.empty_like(A)
        a11 = a1 * a1
        a12 = a1 * a2
        a22 = a2 * a2
        if self.order == 2:
            a3, a4 = A[:, 3], A[:, 4]
            a13, a14, a23, a24, a33 = a1 * a3, a1 * a4, a2 * a3, a2 * a4, a3 * a3
            out[:, 0] = 0.5 * qv2 * a11 + g(self.l01) * a1 + qv2 * a2 + g(self.h0)
            out[:, 1] = qv * a11 + 2.0 * qv2 * a12 + g(self.l11) * a1 + g(self.l12) * a2 + 3.0 * qv2 * a3 + g(self.h1)
            out[:, 2] = (0.5 * v2 * a11 + 2.0 * qv2 * a22 + 4.0 * qv * a12 + 3.0 * qv2 * a13
                         + g(self.l21) * a1 + g(self.l22) * a2 + g(self.l23) * a3 + 6.0 * qv2 * a4 + g(self.h2))
            out[:, 3] = (4.0 * qv * a22 + 2.0 * v2 * a12 + 6.0 * qv * a13 + 4.0 * qv2 * a14 + 6.0 * qv2 * a23
                         + g(self.l32) * a2 + g(self.l33) * a3 + g(self.l34) * a4)
            out[:, 4] = (2.0 * v2 * a22 + 4.5 * qv2 * a33 + 3.0 * v2 * a13 + 8.0 * qv * a14 + 12.0 * qv * a23 + 8.0 * qv2 * a24
                         + g(self.l43) * a3 + g(self.l44) * a4)
        else:
            out[:, 0] = 0.5 * qv2 * a11 + g(self.l01) * a1 + qv2 * a2 + g(self.h0)
            out[:, 1] = qv * a11 + 2.0 * qv2 * a12 + g(self.l11) * a1 + g(self.l12) * a2 + g(self.h1)
            out[:, 2] = 0.5 * v2 * a11 + 2.0 * qv2 * a22 + 4.0 * qv * a12 + g(self.l21) * a1 + g(self.l22) * a2 + g(self.h2)
        return out


# --------------------------------------------------------------------------------------------
# SciPy RK45 (Dormand-Prince 5(4)) controller clone, vectorised over independent grid points
# --------------------------------------------------------------------------------------------
_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
_A = [np.array([]),
      np.array([1 / 5]),
      np.array([3 / 40, 9 / 40]),
      np.array([44 / 45, -56 / 15, 32 / 9]),
      np.array([19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729]),
      np.array([9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656])]
_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
RTOL, ATOL = 1e-3, 1e-6          # solve_ivp defaults used by affine_expansion.py:300-301
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0   # scipy/integrate/_ivp/rk.py:8-11


def _rms(v):
    """scipy/integrate/_ivp/common.py:63-65  norm(x)=||x||_2/sqrt(n), per grid point (rows)."""
    return np.sqrt(np.sum(np.abs(v) ** 2, axis=1)) / np.sqrt(v.shape[1])


def rk45_grid(rhs: LogsvRhs, y0: np.ndarray, t_bound: float, return_stats: bool = False):
    """Integrate y' = rhs(y) from 0 to ``t_bound`` independently for every row of ``y0`` with the exact
    control law of ``scipy.integrate.solve_ivp(method='RK45', rtol=1e-3, atol=1e-6)``:

    * initial step: ``select_initial_step`` (scipy/integrate/_ivp/common.py:109-134);
    * step loop, acceptance, step-size factors: ``RungeKutta._step_impl`` (scipy/.../rk.py:111-170);
    * stages / FSAL: ``rk_step`` (rk.py:14-69), tableau ``RK45`` (rk.py:538-552).

    Rows advance in lock-step *attempts* under an ``active`` mask; each row sees exactly the sequence
    of attempts SciPy would make for it alone.
    """
    P, n = y0.shape
    y = y0.astype(np.complex128).copy()
    t = np.zeros(P)
    f = rhs(y)
    # ---- select_initial_step (order = error_estimator_order = 4)
    scale = ATOL + np.abs(y) * RTOL
    d0 = _rms(y / scale)
    d1 = _rms(f / scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        h0 = np.where((d0 < 1e-5) | (d1 < 1e-5), 1e-6, 0.01 * d0 / d1)
    h0 = np.minimum(h0, t_bound)
    y1 = y + h0[:, None] * f
    f1 = rhs(y1)
    d2 = _rms((f1 - f) / scale) / h0
    with np.errstate(divide="ignore"):
        h1 = np.where((d1 <= 1e-15) & (d2 <= 1e-15), np.maximum(1e-6, h0 * 1e-3), (0.01 / np.maximum(d1, d2)) ** (1 / 5))
    h_abs = np.minimum(np.minimum(100 * h0, h1), t_bound)
    nfev = np.full(P, 2)
    nsteps = np.zeros(P, dtype=int)
    nrej = np.zeros(P, dtype=int)

    active = t < t_bound
    rejected = np.zeros(P, dtype=bool)        # "step_rejected" of the CURRENT step, reset on acceptance
    new_step = np.ones(P, dtype=bool)         # first attempt of a step: apply the min_step clamp
    K = np.zeros((7, P, n), dtype=np.complex128)
    while np.any(active):
        ia = np.nonzero(active)[0]
        ta, ya, fa = t[ia], y[ia], f[ia]
        min_step = 10 * np.abs(np.nextafter(ta, np.inf) - ta)
        ha = h_abs[ia]
        ha = np.where(new_step[ia] & (ha < min_step), min_step, ha)     # max_step = inf
        # (a mid-step h_abs < min_step would return TOO_SMALL_STEP in SciPy; never happens on this path)
        t_new = ta + ha
        over = (t_new - t_bound) > 0
        t_new = np.where(over, t_bound, t_new)
        h = t_new - ta
        ha = np.abs(h)
        Ka = np.zeros((7, ia.shape[0], n), dtype=np.complex128)
        Ka[0] = fa
        for s in range(1, 6):
            dy = sum(Ka[j] * _A[s][j] for j in range(s)) * h[:, None]
            Ka[s] = rhs(ya + dy, ia)
        y_new = ya + h[:, None] * sum(Ka[j] * _B[j] for j in range(6))
        f_new = rhs(y_new, ia)
        Ka[6] = f_new
        nfev[ia] += 6
        scale = ATOL + np.maximum(np.abs(ya), np.abs(y_new)) * RTOL
        err = _rms(sum(Ka[j] * _E[j] for j in range(7)) * h[:, None] / scale)
        acc = err < 1
        with np.errstate(divide="ignore"):
            fac_acc = np.where(err == 0, MAX_FACTOR, np.minimum(MAX_FACTOR, SAFETY * err ** -0.2))
            fac_acc = np.where(rejected[ia], np.minimum(1.0, fac_acc), fac_acc)
            fac_rej = np.maximum(MIN_FACTOR, SAFETY * err ** -0.2)
        h_abs[ia] = ha * np.where(acc, fac_acc, fac_rej)
        # commit accepted rows
        ja = ia[acc]
        t[ja] = t_new[acc]
        y[ja] = y_new[acc]
        f[ja] = f_new[acc]
        nsteps[ja] += 1
        rejected[ja] = False
        new_step[ja] = True
        jr = ia[~acc]
        rejected[jr] = True
        new_step[jr] = False
        nrej[jr] += 1
        active = t < t_bound
    if return_stats:
        return y, dict(nfev=nfev, nsteps=nsteps, nrej=nrej)
    return y


# --------------------------------------------------------------------------------------------
# log-MGF grids and Fourier sums
# --------------------------------------------------------------------------------------------
def logsv_a_mgf_grid(dtau, phi, psi, a_t0, sigma0, theta, kappa1, kappa2, beta, volvol, is_spot_measure=True, order=2, eta=1.0):
    """``compute_logsv_a_mgf_grid`` non-analytic branch (affine_expansion.py:570-685 -> :492-529):
    RK45 per grid point from ``a_t0`` over ``dtau``; log_mgf = sum_k A_k y^k, y = sigma0 - theta."""
    rhs = LogsvRhs(theta, kappa1, kappa2, beta, volvol, phi, psi, is_spot_measure, order, eta)
    a_t1 = rk45_grid(rhs, a_t0, dtau)
    yv = sigma0 - theta
    ys = np.array([1.0, yv, yv * yv, yv * yv * yv, (yv * yv) * (yv * yv)])[: expansion_n(order)]
    return a_t1, a_t1 @ (ys + 0j)


def heston_mgf_grid(v0, theta, kappa, volvol, rho, dtau, phi, psi, a_t0=None, b_t0=None):
    """``compute_heston_mgf_grid`` (pricers/heston_pricer.py:183-214), principal-branch sqrt/log."""
    vv2 = volvol * volvol
    b1 = kappa + rho * volvol * phi
    b0 = 0.5 * phi * (phi + 1.0) - psi
    zeta = np.sqrt(b1 * b1 - 2.0 * b0 * vv2)
    ez = np.exp(-zeta * dtau)
    psi_p, psi_m = -b1 + zeta, b1 + zeta
    if b_t0 is None:
        c_p, c_m = psi_p / (2.0 * zeta), psi_m / (2.0 * zeta)
    else:
        c_p, c_m = (psi_p + vv2 * b_t0) / (2.0 * zeta), (psi_m - vv2 * b_t0) / (2.0 * zeta)
    b_t1 = -(-psi_m * c_p * ez + psi_p * c_m) / (vv2 * (c_p * ez + c_m))
    a_t1 = -(theta * kappa / vv2) * (psi_p * dtau + 2.0 * np.log(c_p * ez + c_m))
    if a_t0 is not None:
        a_t1 = a_t1 + a_t0
    return a_t1 + b_t1 * v0, a_t1, b_t1


def vanilla_slice_prices(log_mgf, phi, forward, strikes, types, discfactor=1.0, is_spot_measure=True):
    """``vanilla_slice_pricer_with_mgf_grid`` (utils/mgf_pricer.py:174-221), |Re phi| = 1/2 branch and general branch."""
    p = np.imag(phi)
    dp = legacy_simpson_weights(phi)
    if np.all(np.abs(np.real(phi)) == 0.5):
        w = (dp / np.pi) / (p * p + 0.25) + 0j
    elif is_spot_measure:
        w = -(dp / np.pi) / ((phi + 1.0) * phi)
    else:
        w = -(dp / np.pi) / ((phi - 1.0) * phi)
    out = np.zeros(len(strikes))
    for j, (k, ty) in enumerate(zip(strikes, types)):
        x = np.log(forward / k)
        capped = np.nansum(np.real(w * np.exp(-x * phi + log_mgf)))
        ty = str(ty)
        if is_spot_measure:
            if ty == "C":
                out[j] = discfactor * (forward - k * capped)
            elif ty == "P":
                out[j] = discfactor * (k - k * capped)
            else:
                raise ValueError("not implemented")
        else:
            if ty in ("IC", "C"):
                out[j] = forward * discfactor * (1.0 - capped)
            elif ty in ("IP", "P"):
                out[j] = forward * discfactor * (np.exp(-x) - capped)
            else:
                raise ValueError("not implemented")
    return out


def logsv_chain_prices(params6, ttms, forwards, discfactors, strikes_ttms, types_ttms, is_spot_measure=True, order=2,
                       etas=None, vol_scaler=None, return_grids=False):
    """``logsv_chain_pricer`` LOG_RETURN branch (pricers/logsv_pricer.py:669-739)."""
    sigma0, theta, kappa1, kappa2, beta, volvol = params6
    if vol_scaler is None:
        vol_scaler = logsv_vol_scaler(sigma0, ttms)
    phi = phi_grid(vol_scaler, is_spot_measure)
    psi = np.zeros_like(phi)
    a = np.zeros((phi.shape[0], expansion_n(order)), dtype=np.complex128)
    etas = np.ones(len(ttms)) if etas is None else etas
    t0 = 0.0
    prices, grids = [], []
    for m, ttm in enumerate(ttms):
        a, log_mgf = logsv_a_mgf_grid(ttm - t0, phi, psi, a, sigma0, theta, kappa1, kappa2, beta, volvol, is_spot_measure, order, etas[m])
        prices.append(vanilla_slice_prices(log_mgf, phi, forwards[m], strikes_ttms[m], types_ttms[m], discfactors[m], is_spot_measure))
        grids.append((a.copy(), log_mgf))
        t0 = ttm
    return (prices, grids) if return_grids else prices


def heston_chain_prices(params5, ttms, forwards, discfactors, strikes_ttms, types_ttms, vol_scaler=None, return_grids=False):
    """``heston_chain_pricer`` LOG_RETURN branch (pricers/heston_pricer.py:217-282); params5 = (v0, theta, kappa, rho, volvol)."""
    v0, theta, kappa, rho, volvol = params5
    if vol_scaler is None:
        vol_scaler = heston_vol_scaler(v0, ttms)
    phi = phi_grid(vol_scaler, True)
    psi = np.zeros_like(phi)
    a = np.zeros_like(phi)
    b = np.zeros_like(phi)
    t0 = 0.0
    prices, grids = [], []
    for m, ttm in enumerate(ttms):
        log_mgf, a, b = heston_mgf_grid(v0, theta, kappa, volvol, rho, ttm - t0, phi, psi, a, b)
        prices.append(vanilla_slice_prices(log_mgf, phi, forwards[m], strikes_ttms[m], types_ttms[m], discfactors[m], True))
        grids.append((log_mgf, a, b))
        t0 = ttm
    return (prices, grids) if return_grids else prices


# --------------------------------------------------------------------------------------------
# quadratic-variance options, densities, digitals
# --------------------------------------------------------------------------------------------
def psi_grid():
    """``get_psi_grid`` (utils/mgf_pricer.py:37-47)."""
    return -0.5 + 1j * np.linspace(0, 4000, 40000)


def theta_grid():
    """``get_theta_grid`` (utils/mgf_pricer.py:50-58)."""
    return 0.0 + 1j * np.linspace(0, 600, 5000)


def transform_var_grid(variable_type, is_spot_measure=True, vol_scaler=0.28):
    """``get_transform_var_grid`` (utils/mgf_pricer.py:61-94); variable_type 1 LOG_RETURN, 2 Q_VAR, 3 SIGMA."""
    if variable_type == 1:
        phi = phi_grid(vol_scaler, is_spot_measure)
        return phi, np.zeros_like(phi), np.zeros_like(phi)
    if variable_type == 2:
        psi = psi_grid()
        return (np.zeros_like(psi) if is_spot_measure else np.ones_like(psi)), psi, np.zeros_like(psi)
    if variable_type == 3:
        th = theta_grid()
        return np.zeros_like(th), np.zeros_like(th), th
    raise NotImplementedError


def qvar_slice_prices(log_mgf, psi, ttm, strikes, types, discfactor=1.0):
    """``slice_qvar_pricer_with_a_grid`` (utils/mgf_pricer.py:323-358)."""
    dp = legacy_simpson_weights(psi)
    w = (dp / np.pi) / (psi * psi)
    out = np.zeros(len(strikes))
    for j, (k, ty) in enumerate(zip(strikes, types)):
        if str(ty) != "C":
            raise ValueError("not implemented")
        s = np.nansum(np.real(w * np.exp((k * ttm) * psi + log_mgf)))
        out[j] = np.maximum(discfactor * s / ttm, 1e-10)
    return out


def pdf_from_mgf(log_mgf, grid, space_grid, shift=0.0, scale=1.0):
    """``pdf_with_mgf_grid`` (utils/mgf_pricer.py:361-384)."""
    dp = legacy_simpson_weights(grid) / np.pi
    z = (space_grid - shift) / scale
    pdf = np.array([np.nansum(np.real(dp * np.exp(x * grid + log_mgf))) for x in z])
    return (space_grid[1] - space_grid[0]) * pdf


def digital_slice_prices(log_mgf, phi, forward, strikes, types, discfactor=1.0):
    """``digital_slice_pricer_with_mgf_grid`` (utils/mgf_pricer.py:224-269)."""
    dp = legacy_simpson_weights(phi)
    all_calls = bool(np.all(np.real(phi) < 0.0))
    w = -(dp / np.pi) / phi if all_calls else (dp / np.pi) / phi
    out = np.zeros(len(strikes))
    for j, (k, ty) in enumerate(zip(strikes, types)):
        x = np.log(forward / k)
        s = np.nansum(np.real(w * np.exp(-x * phi + log_mgf)))
        ty = str(ty)
        if ty == "C":
            price = s if all_calls else 1.0 - s
        elif ty == "P":
            price = 1.0 - s if all_calls else s
        else:
            raise ValueError("not implemented")
        out[j] = discfactor * price
    return out


def logsv_qvar_chain_prices(params6, ttms, discfactors, strikes_ttms, types_ttms, is_spot_measure=True, order=2, return_grids=False):
    """``logsv_chain_pricer`` Q_VAR branch (pricers/logsv_pricer.py:669-731)."""
    sigma0, theta, kappa1, kappa2, beta, volvol = params6
    phi, psi, _ = transform_var_grid(2, is_spot_measure)
    a = np.zeros((psi.shape[0], expansion_n(order)), dtype=np.complex128)
    t0, prices, grids = 0.0, [], []
    for m, ttm in enumerate(ttms):
        a, lm = logsv_a_mgf_grid(ttm - t0, phi, psi, a, sigma0, theta, kappa1, kappa2, beta, volvol, is_spot_measure, order, 1.0)
        prices.append(qvar_slice_prices(lm, psi, ttm, strikes_ttms[m], types_ttms[m], discfactors[m]))
        grids.append((a.copy(), lm))
        t0 = ttm
    return (prices, grids) if return_grids else prices


def logsv_pdf(params6, ttm, space_grid, variable_type=1, is_spot_measure=True, order=2):
    """``logsv_pdfs`` (pricers/logsv_pricer.py:742-803)."""
    sigma0, theta, kappa1, kappa2, beta, volvol = params6
    vs = logsv_vol_scaler(sigma0, np.array([ttm]))
    phi, psi, th = transform_var_grid(variable_type, is_spot_measure, vs)
    a0 = np.zeros((phi.shape[0], expansion_n(order)), dtype=np.complex128)
    if variable_type == 3:
        a0[:, 1] = -th                     # get_init_conditions_a, affine_expansion.py:562-564
    _, lm = logsv_a_mgf_grid(ttm, phi, psi, a0, sigma0, theta, kappa1, kappa2, beta, volvol, is_spot_measure, order, 1.0)
    grid, shift, scale = ((phi, 0.0, 1.0), (psi, 0.0, 1.0 / ttm), (th, theta, 1.0))[variable_type - 1]
    return pdf_from_mgf(lm, grid, space_grid, shift, scale) / scale


def heston_qvar_chain_prices(params5, ttms, discfactors, strikes_ttms, types_ttms):
    """``heston_chain_pricer`` Q_VAR branch (pricers/heston_pricer.py:217-282)."""
    v0, theta, kappa, rho, volvol = params5
    phi, psi, _ = transform_var_grid(2, True)
    a = np.zeros_like(psi)
    b = np.zeros_like(psi)
    t0, prices = 0.0, []
    for m, ttm in enumerate(ttms):
        lm, a, b = heston_mgf_grid(v0, theta, kappa, volvol, rho, ttm - t0, phi, psi, a, b)
        prices.append(qvar_slice_prices(lm, psi, ttm, strikes_ttms[m], types_ttms[m], discfactors[m]))
        t0 = ttm
    return prices


# ---------------------------------------------------------------------------------------------------------------------------------
# semi-analytic branch (is_analytic=True)
# ---------------------------------------------------------------------------------------------------------------------------------
def logsv_analytic_a_grid(dtau, phi, psi, a_t0, theta, kappa1, kappa2, beta, volvol, is_spot_measure=True, order=2, year_days=260):
    """``solve_analytic_ode_grid_phi`` / ``solve_analytic_ode_for_a`` (pricers/logsv/affine_expansion.py:306-470), restated WITHOUT the
    eigendecomposition, the way the CUDA kernel does it: ``v diag(exp(w dt)) v^-1 = expm(L dt)``; column 0 of L is zero, so the reference's
    ``v diag((exp(w dt) - 1)/w, with the zero eigenvalue's entry := 0) v^-1`` equals ``int_0^dt exp(L s) ds`` in rows 1.. (its row 0 is
    overwritten by (H_0 + quad_0) dt).  Business-day steps, 10 fixed-point sweeps per step, eta = 1.  Pinned by
    tests/golden/logsv_analytic_branch.npz (outputs of the reference's LAPACK route)."""
    from scipy.linalg import expm
    n = expansion_n(order)
    nb_steps = int(np.ceil(year_days * dtau))
    dt = dtau / nb_steps
    out = np.empty((phi.shape[0], n), dtype=np.complex128)
    for p in range(phi.shape[0]):
        M, L, H = logsv_mlh(theta, kappa1, kappa2, beta, volvol, complex(phi[p]), complex(psi[p]), is_spot_measure, order, 1.0)
        aug = np.zeros((2 * n, 2 * n), dtype=np.complex128)          # expm([[L, I], [0, 0]] dt) = [[E, Psi], [0, I]]
        aug[:n, :n] = L * dt
        aug[:n, n:] = np.eye(n) * dt
        big = expm(aug)
        E, Psi = big[:n, :n], big[:n, n:]
        g = Psi @ H
        a = np.array(a_t0[p], dtype=np.complex128)
        with np.errstate(all="ignore"):
            for _ in range(nb_steps):
                Ea = E @ a
                fp = a
                for _ in range(10):
                    quad = np.array([fp @ M[k] @ fp for k in range(n)])
                    rhs = g + quad * dt
                    rhs[0] = (H[0] + quad[0]) * dt
                    fp = Ea + rhs
                a = fp
        out[p] = a
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# stiff branch (is_stiff_solver=True)
# ---------------------------------------------------------------------------------------------------------------------------------
def logsv_bdf_a_grid(dtau, phi, psi, a_t0, theta, kappa1, kappa2, beta, volvol, is_spot_measure=True, order=2, eta=1.0, return_stats=False):
    """``solve_ode_for_a(is_stiff_solver=True)`` over a grid (pricers/logsv/affine_expansion.py:229-303, 492-529): SciPy's own
    ``solve_ivp(method='BDF', jac=func_rhs_jac)`` (default rtol 1e-3 / atol 1e-6) per grid point on the restated M, L, H.  The third-party
    arithmetic here IS the reference's (scipy 1.18.1, scipy/integrate/_ivp/bdf.py): the oracle calls it, the CUDA kernel clones its control law."""
    from scipy.integrate import solve_ivp
    n = expansion_n(order)
    out = np.empty((phi.shape[0], n), dtype=np.complex128)
    stats = []
    for p in range(phi.shape[0]):
        M, L, H = logsv_mlh(theta, kappa1, kappa2, beta, volvol, complex(phi[p]), complex(psi[p]), is_spot_measure, order, eta)
        fun = lambda t, A: np.array([A @ M[k] @ A for k in range(n)]) + L @ A + H
        jac = lambda t, A: np.array([2.0 * M[k] @ A for k in range(n)]) + L
        sol = solve_ivp(fun=fun, t_span=(0.0, dtau), y0=np.array(a_t0[p], dtype=np.complex128), method="BDF", jac=jac)
        out[p] = sol.y[:, -1]
        stats.append((sol.nfev, sol.njev, sol.nlu, sol.t.size - 1))
    return (out, np.array(stats)) if return_stats else out
