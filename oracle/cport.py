"""ctypes binding of the C port (oracle/csrc/oracle_mc.c).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle_mc.so")
_TYPE_CODES = {"C": 0, "P": 1, "IC": 2, "IP": 3}
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "oracle_mc.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_mc_chain.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(load().oracle_num_threads())


def mc_chain(model: str, params, ttms, forwards, discfactors, etas, strikes_ttms, types_ttms, nb_path, nb_steps_per_year,
             is_spot_measure=True, variable_type=1, seed=10, gauss="f64", nthreads=0, path_offset=0, return_states=False):
    """C port of logsv_mc_chain_pricer / heston_mc_chain_pricer with the device Philox stream; returns (prices, stderrs[, states])."""
    lib = load()
    dp = ctypes.POINTER(ctypes.c_double)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    params, ttms, forwards, discfactors = f64(params), f64(ttms), f64(forwards), f64(discfactors)
    M = ttms.shape[0]
    etas = f64(np.ones(M) if etas is None else etas)
    sizes = [len(s) for s in strikes_ttms]
    offsets = np.zeros(M + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(sizes)
    strikes = f64(np.concatenate([np.asarray(s, dtype=float) for s in strikes_ttms]))
    types = np.ascontiguousarray([_TYPE_CODES[str(t)] for tt in types_ttms for t in tt], dtype=np.int8)
    prices, stds = np.empty(strikes.shape[0]), np.empty(strikes.shape[0])
    states = np.empty((3, nb_path)) if return_states else None
    rc = lib.oracle_mc_chain(ctypes.c_int(0 if model == "logsv" else 1), params.ctypes.data_as(dp), ctypes.c_int(M),
                             ttms.ctypes.data_as(dp), forwards.ctypes.data_as(dp), discfactors.ctypes.data_as(dp),
                             etas.ctypes.data_as(dp), offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                             strikes.ctypes.data_as(dp), types.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
                             ctypes.c_longlong(nb_path), ctypes.c_longlong(path_offset), ctypes.c_int(nb_steps_per_year),
                             ctypes.c_int(int(bool(is_spot_measure))), ctypes.c_int(int(variable_type)), ctypes.c_uint64(seed),
                             ctypes.c_int(1 if gauss == "f64" else 0), ctypes.c_int(nthreads), prices.ctypes.data_as(dp),
                             stds.ctypes.data_as(dp), states.ctypes.data_as(dp) if return_states else None)
    if rc != 0:
        raise MemoryError("oracle_mc_chain allocation failed")
    split = lambda a: [a[offsets[m]:offsets[m + 1]].copy() for m in range(M)]
    return (split(prices), split(stds), states) if return_states else (split(prices), split(stds))
