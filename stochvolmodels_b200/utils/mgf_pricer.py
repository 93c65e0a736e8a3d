"""Transform grid (host, trivial) and the Fourier vanilla sum (GPU): reference utils/mgf_pricer.py."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .. import engine
from .config import VariableType


def get_phi_grid(is_spot_measure: bool = True, max_phi: int = 1000, vol_scaler: float = 0.28, real_phi: float = None) -> np.ndarray:
    """phi = -1/2 (MMA) | +1/2 (inverse) + i*linspace(0, 5.6/vol_scaler, max_phi) (utils/mgf_pricer.py:11-34)."""
    p = np.linspace(0, 5.6 / vol_scaler, max_phi)
    if real_phi is None:
        real_phi = -0.5 if is_spot_measure else 0.5
    return real_phi + 1j * p


def get_psi_grid() -> np.ndarray:
    """transform grid for the quadratic variance: -1/2 + i*linspace(0, 4000, 40000) (utils/mgf_pricer.py:37-47)."""
    return -0.5 + 1j * np.linspace(0, 4000, 40000)


def get_theta_grid() -> np.ndarray:
    """transform grid for the volatility: i*linspace(0, 600, 5000) (utils/mgf_pricer.py:50-58)."""
    return 0.0 + 1j * np.linspace(0, 600, 5000)


def get_transform_var_grid(variable_type: VariableType = VariableType.LOG_RETURN, is_spot_measure: bool = True, max_phi: int = 1000,
                           vol_scaler: float = 0.28, real_phi: float = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(phi, psi, theta) grids for Fourier inversion in each state variable (utils/mgf_pricer.py:61-94)."""
    vt = getattr(variable_type, "value", variable_type)
    if vt == 1:
        phi_grid = get_phi_grid(is_spot_measure=is_spot_measure, max_phi=max_phi, vol_scaler=vol_scaler, real_phi=real_phi)
        return phi_grid, np.zeros_like(phi_grid), np.zeros_like(phi_grid)
    if vt == 2:
        psi_grid = get_psi_grid()
        phi_grid = np.zeros_like(psi_grid) if is_spot_measure else np.ones_like(psi_grid)
        return phi_grid, psi_grid, np.zeros_like(psi_grid)
    if vt == 3:
        theta_grid = get_theta_grid()
        return np.zeros_like(theta_grid), np.zeros_like(theta_grid), theta_grid
    raise NotImplementedError


def vanilla_slice_pricer_with_mgf_grid(log_mgf_grid: np.ndarray, phi_grid: np.ndarray, forward: float, strikes: np.ndarray,
                                       optiontypes: np.ndarray, discfactor: float = 1.0, is_spot_measure: bool = True) -> np.ndarray:
    """Simpson-weighted Fourier inversion for vanilla options on the GPU (utils/mgf_pricer.py:174-221): legacy weights on the
    even-length grid, nansum, MMA measure rejects 'IC'/'IP' with ``ValueError("not implemented")``."""
    return engine.fourier_vanilla(log_mgf_grid, phi_grid, forward, strikes, optiontypes, discfactor, is_spot_measure)


def slice_pricer_with_mgf_grid_with_gamma(log_mgf_grid: np.ndarray, phi_grid: np.ndarray, risk_premia_gamma: float, ttm: float, forward: float,
                                          normalizer: float, gamma_forward: float, strikes: np.ndarray, optiontypes: np.ndarray,
                                          discfactor: float = 1.0, is_spot_measure: bool = True, is_simpson: bool = True) -> np.ndarray:
    """vanilla prices under the risk kernel exp(-gamma x) from a log-MGF grid on Re(phi) = -1/2 - gamma (utils/mgf_pricer.py:273-320):
    calls ``gamma_forward - normalizer K^(1+gamma) S``, puts ``K - normalizer K^(1+gamma) S``; ``ttm`` and ``discfactor`` are accepted and
    unused, as in the reference; anything but 'C' / 'P' under the spot measure is ``ValueError("not implemented")``."""
    if not is_simpson:
        raise NotImplementedError("is_simpson=False: the CUDA sums build the legacy Simpson weights")
    return engine.fourier_gamma(log_mgf_grid, phi_grid, risk_premia_gamma, forward, normalizer, gamma_forward, strikes, optiontypes, is_spot_measure)


def slice_qvar_pricer_with_a_grid(log_mgf_grid: np.ndarray, psi_grid: np.ndarray, ttm: float, strikes: np.ndarray, optiontypes: np.ndarray,
                                  forward: float, discfactor: float = 1.0, is_spot_measure: bool = True) -> np.ndarray:
    """calls on the annualised quadratic variance from the log-MGF on the psi grid, on the GPU (utils/mgf_pricer.py:323-358);
    anything but 'C' raises ``ValueError("not implemented")``; ``forward`` is accepted and unused as in the reference."""
    return engine.fourier_qvar(log_mgf_grid, psi_grid, ttm, strikes, optiontypes, discfactor)


def digital_slice_pricer_with_mgf_grid(log_mgf_grid: np.ndarray, phi_grid: np.ndarray, forward: float, strikes: np.ndarray,
                                       optiontypes: np.ndarray, discfactor: float = 1.0) -> np.ndarray:
    """digital calls / puts on the spot from the log-MGF grid, on the GPU (utils/mgf_pricer.py:224-269)."""
    return engine.fourier_digital(log_mgf_grid, phi_grid, forward, strikes, optiontypes, discfactor)


def pdf_with_mgf_grid(log_mgf_grid: np.ndarray, transform_var_grid: np.ndarray, space_grid: np.ndarray, shift: float = 0.0,
                      scale: float = 1.0) -> np.ndarray:
    """density on ``space_grid`` by Fourier inversion of the log-MGF, sums on the GPU (utils/mgf_pricer.py:361-384)."""
    space_grid = np.asarray(space_grid, dtype=np.float64)
    z = (space_grid - shift) / scale
    dx = space_grid[1] - space_grid[0]
    return dx * engine.fourier_pdf_sums(log_mgf_grid, transform_var_grid, z)


def compute_integration_weights(var_grid: np.ndarray, is_simpson: bool = True) -> np.ndarray:
    """composite Simpson (odd point count) or trapezoidal weights on the imaginary part of a transform grid, with the validation of the
    reference's public helper (utils/mgf_pricer.py:98-154: finite, strictly increasing, uniformly spaced to 1e-12 relative; the error texts
    are part of its contract).  Host numpy: the pricing kernels build the reference's *legacy* weights (:158-171, even-grid quirk included)
    in registers and never materialise a weight array."""
    p = np.imag(np.asarray(var_grid))
    if p.size < (3 if is_simpson else 2):
        raise ValueError("integration grid is too short for the selected rule")
    if not np.isfinite(p).all():
        raise ValueError("integration grid must contain only finite values")
    h = np.diff(p)
    if (h <= 0.0).any():
        raise ValueError("integration grid must be strictly increasing")
    if (np.abs(h - h[0]) > 1.0e-12 * max(1.0, abs(h[0]))).any():
        raise ValueError("integration grid must be uniformly spaced")
    if is_simpson:
        if p.size % 2 == 0:
            raise ValueError("Simpson integration requires an odd number of grid points")
        pattern = np.where(np.arange(p.size) % 2 == 1, 4.0, 2.0)
        pattern[0] = pattern[-1] = 1.0
        return (p[1] - p[0]) / 3.0 * pattern
    w = np.full(p.size, h[0])
    w[0] = w[-1] = 0.5 * h[0]
    return w


def _compute_legacy_pricer_weights(var_grid: np.ndarray, is_simpson: bool = True) -> np.ndarray:
    """the weights the pricing formulas actually use (reference utils/mgf_pricer.py:158-171; the CUDA sum kernels build the same pattern in
    registers): Simpson 1,4,2,...,4,1 scaled by h/3 WITHOUT the odd-size check -- on an even grid the last weight is 4h/3 -- or, without
    Simpson, the first half step followed by the forward differences."""
    p = np.imag(np.asarray(var_grid))
    if is_simpson:
        pattern = np.where(np.arange(p.size) % 2 == 1, 4.0, 2.0)
        pattern[0] = 1.0
        if p.size % 2 == 1:
            pattern[-1] = 1.0
        return (p[1] - p[0]) / 3.0 * pattern
    return np.concatenate(([0.5 * (p[1] - p[0])], np.diff(p)))
