"""Host-side helpers of the hot path: time grid rule, timer decorator, flattening (reference utils/funcs.py)."""
from __future__ import annotations

import functools
import logging
import time
from typing import List, Tuple

import numpy as np


def to_flat_np_array(input_list: List[np.ndarray]) -> np.ndarray:
    """concatenate per-maturity arrays into one flat array (utils/funcs.py:19-21)."""
    return np.concatenate(input_list).ravel()


def set_time_grid(ttm: float, nb_steps_per_year: int = 360) -> Tuple[int, float, np.ndarray]:
    """``nb_steps = int(ttm*n) + 1``, ``dt = grid[1] - grid[0]``, grid = linspace(0, ttm, nb_steps+1)
    (utils/funcs.py:24-47).  This rule DEFINES the step count of every MC slice."""
    nb_steps = int(ttm * nb_steps_per_year) + 1
    grid_t = np.linspace(0.0, ttm, nb_steps + 1)
    dt = grid_t[1] - grid_t[0]
    return nb_steps, float(dt), grid_t


def timer(func):
    """log the wall-clock runtime of the wrapped call at DEBUG level (utils/funcs.py:63-78)."""
    @functools.wraps(func)
    def wrapper_timer(*args, **kwargs):
        start = time.perf_counter()
        value = func(*args, **kwargs)
        logging.getLogger(func.__module__).debug("Finished %r in %.4f secs", func.__name__, time.perf_counter() - start)
        return value
    return wrapper_timer


def set_seed(value) -> None:
    """reference utils/funcs.py:51-60: make the seedless Monte Carlo calls that follow reproducible (see engine.set_seed)."""
    from .. import engine
    engine.set_seed(value)
