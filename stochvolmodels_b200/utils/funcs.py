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


# ---- small host utilities of the reference's utils/funcs.py (:79-175), kept for callers that import them from there -------------------
def update_kwargs(kwargs, new_kwargs):
    """a copy of ``kwargs`` updated with ``new_kwargs`` (None or empty: unchanged copy)"""
    return {**kwargs, **(new_kwargs or {})}


def erfcc(x):
    """complementary error function; the reference carries a 1.2e-7 rational approximation for Numba, here SciPy's full-precision erfc"""
    from scipy.special import erfc
    return erfc(x)


def ncdf(x):
    """standard normal distribution function"""
    from scipy.special import ndtr
    return ndtr(x)


def npdf(x, mu: float = 0.0, vol: float = 1.0):
    """normal density with mean ``mu`` and standard deviation ``vol``"""
    z = (np.asarray(x, dtype=float) - mu) / vol
    return np.exp(-0.5 * z * z) / (vol * np.sqrt(2.0 * np.pi))


def find_nearest(a: np.ndarray, value: float, is_sorted: bool = True, is_equal_or_largest: bool = False) -> float:
    """the element of ``a`` closest to ``value``; with ``is_equal_or_largest`` (sorted input) the first element at or above it -- the maturity
    lookup of the vol backbone (reference :131-175; ties go to the upper neighbour as there)"""
    a = np.asarray(a)
    if not is_sorted:
        return a[np.abs(a - value).argmin()]
    idx = int(np.searchsorted(a, value, side="left"))
    if is_equal_or_largest:
        return a[idx]
    if idx > 0 and (idx == len(a) or abs(value - a[idx - 1]) < abs(value - a[idx])):
        return a[idx - 1]
    return a[idx]


def compute_histogram_data(data: np.ndarray, x_grid: np.ndarray, name: str = "Histogram"):
    """frequencies of ``data`` on the bins of ``x_grid`` as a pandas Series indexed by the bin edges (the first entry carries x_grid[0] / n, as
    in the reference :79-92)"""
    import pandas as pd
    counts, edges = np.histogram(a=data, bins=len(x_grid) - 1, range=(x_grid[0], x_grid[-1]))
    return pd.Series(np.append(np.array(x_grid[0]), counts) / len(data), index=edges, name=name)
