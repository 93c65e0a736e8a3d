"""oracle.bsm -- Black-76 prices and implied-volatility inversion on the host (numpy).  TEST INFRASTRUCTURE ONLY: the checker of the
CUDA kernel ``black_ivol_kernel`` (stochvolmodels_b200/csrc/ivol_kernels.cu), same bracketed bisection.

In the reference this step is third-party: ``vanilla_option_pricers.infer_bsm_ivols_from_model_chain_prices`` called from
``data/option_chain.py:327-346``; that package is not part of the reference tree and is absent here, so bit-level parity
with it is UNPINNED.  What the reference pins is satisfied by any correct inversion: the quickstart
implied vols 0.999577 / 0.995757 (examples/getting_started/quickstart.py:44,46, rtol 5e-6) and flat-vol round trips.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
from scipy.special import ndtr


def compute_bsm_vanilla_price(forward, strike, ttm, vol, optiontype="C", discfactor=1.0):
    """undiscounted-forward Black-76 price times discfactor; 'IC'/'IP' are quoted in the same units as 'C'/'P'."""
    forward, strike, vol = np.asarray(forward, float), np.asarray(strike, float), np.asarray(vol, float)
    sdev = vol * np.sqrt(ttm)
    with np.errstate(divide="ignore", invalid="ignore"):
        d1 = np.log(forward / strike) / sdev + 0.5 * sdev
    d2 = d1 - sdev
    call = forward * ndtr(d1) - strike * ndtr(d2)
    is_call = np.isin(np.asarray(optiontype), ("C", "IC"))
    return discfactor * np.where(is_call, call, call - (forward - strike))


def infer_bsm_implied_vol(forward, ttm, strikes, given_prices, optiontypes, discfactor=1.0, lo=1e-8, hi=10.0):
    """implied vols of one slice by bracketed bisection (80 halvings => ~1e-23 bracket) -- robust for deep OTM."""
    strikes = np.asarray(strikes, float)
    prices = np.asarray(given_prices, float) / discfactor
    is_call = np.isin(np.asarray(optiontypes), ("C", "IC"))
    # work with the OTM-equivalent call price via parity for stability
    call_prices = np.where(is_call, prices, prices + (forward - strikes))
    intrinsic = np.maximum(forward - strikes, 0.0)
    ok = (call_prices > intrinsic) & (call_prices < forward) & np.isfinite(call_prices)
    a = np.full(strikes.shape, lo)
    b = np.full(strikes.shape, hi)
    for _ in range(80):
        mid = 0.5 * (a + b)
        pm = compute_bsm_vanilla_price(forward, strikes, ttm, mid, "C")
        up = pm < call_prices
        a = np.where(up, mid, a)
        b = np.where(up, b, mid)
    vol = 0.5 * (a + b)
    return np.where(ok, vol, np.nan)


def infer_bsm_ivols_from_model_chain_prices(ttms, forwards, discfactors, strikes_ttms: Sequence[np.ndarray],
                                            optiontypes_ttms: Sequence[np.ndarray], model_prices_ttms: Sequence[np.ndarray]
                                            ) -> List[np.ndarray]:
    """same call shape as the third-party function used at data/option_chain.py:340-345."""
    out = []
    for ttm, forward, df, strikes, types, prices in zip(ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms, model_prices_ttms):
        out.append(infer_bsm_implied_vol(forward, ttm, strikes, prices, types, df))
    return out
