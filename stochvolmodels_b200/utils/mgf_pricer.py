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


def get_transform_var_grid(variable_type: VariableType = VariableType.LOG_RETURN, is_spot_measure: bool = True, max_phi: int = 1000,
                           vol_scaler: float = 0.28, real_phi: float = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(phi, psi, theta) grids for LOG_RETURN (utils/mgf_pricer.py:61-77); Q_VAR / SIGMA grids are SURVEY.md §8f "next"."""
    if variable_type == VariableType.LOG_RETURN:
        phi_grid = get_phi_grid(is_spot_measure=is_spot_measure, max_phi=max_phi, vol_scaler=vol_scaler, real_phi=real_phi)
        return phi_grid, np.zeros_like(phi_grid), np.zeros_like(phi_grid)
    raise NotImplementedError


def vanilla_slice_pricer_with_mgf_grid(log_mgf_grid: np.ndarray, phi_grid: np.ndarray, forward: float, strikes: np.ndarray,
                                       optiontypes: np.ndarray, discfactor: float = 1.0, is_spot_measure: bool = True) -> np.ndarray:
    """Simpson-weighted Fourier inversion for vanilla options on the GPU (utils/mgf_pricer.py:174-221): legacy weights on the
    even-length grid, nansum, MMA measure rejects 'IC'/'IP' with ``ValueError("not implemented")``."""
    return engine.fourier_vanilla(log_mgf_grid, phi_grid, forward, strikes, optiontypes, discfactor, is_spot_measure)
