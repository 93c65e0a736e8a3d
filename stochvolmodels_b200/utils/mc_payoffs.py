"""``compute_mc_vars_payoff`` on the GPU: same signature as the reference (utils/mc_payoffs.py:10-88)."""
from __future__ import annotations

import numpy as np

from .. import engine
from .config import VariableType


def compute_mc_vars_payoff(x0: np.ndarray, sigma0: np.ndarray, qvar0: np.ndarray, ttm: float, forward: float,
                           strikes_ttm: np.ndarray, optiontypes_ttm: np.ndarray, discfactor: float = 1.0,
                           variable_type: VariableType = VariableType.LOG_RETURN):
    """discounted per-strike mean payoff and its standard error; spots re-centred on ``forward`` with the nan-mean over all
    paths; ``sigma0`` accepted for signature symmetry and unused, as in the reference.  Raises ``ValueError`` for an unknown
    payoff code and ``NotImplementedError`` for ``VariableType.SIGMA``."""
    return engine.mc_payoffs(x0, qvar0, ttm, forward, strikes_ttm, optiontypes_ttm, discfactor, variable_type)
