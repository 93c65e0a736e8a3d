"""OptionChain: the ragged chain container that crosses the Pricer boundary.

Host-side data type only ("boundary input type").  It keeps the attributes the hot path reads --
``ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms`` (reference pricers/logsv_pricer.py:358-365, 412-427) -- with the
reference's validation rules (data/option_chain.py:147-215) and constructors ``slice_to_chain`` (:230-245) and
``get_uniform_chain`` (:462-492).  The pricers are duck-typed: a reference ``stochvolmodels.OptionChain`` works as well.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_VALID_TYPES = ("C", "P", "IC", "IP")


def _validate_option_slice_data(strikes, optiontypes) -> int:
    strikes = np.asarray(strikes)
    optiontypes = np.asarray(optiontypes)
    if strikes.ndim != 1 or strikes.size == 0:
        raise ValueError("strikes must be a non-empty one-dimensional array")
    if not np.all(np.isfinite(strikes)) or np.any(strikes <= 0.0):
        raise ValueError("strikes must contain only finite positive values")
    if optiontypes.shape != strikes.shape:
        raise ValueError("strikes and optiontypes must have the same length")
    if not np.all(np.isin(optiontypes, _VALID_TYPES)):
        raise ValueError("optiontypes must be one of 'C', 'P', 'IC', 'IP'")
    return strikes.size


@dataclass
class OptionChain:
    ttms: np.ndarray
    forwards: np.ndarray
    strikes_ttms: Sequence[np.ndarray]
    optiontypes_ttms: Sequence[np.ndarray]
    ids: Optional[np.ndarray] = None
    discfactors: Optional[np.ndarray] = None
    discount_rates: Optional[np.ndarray] = None
    ticker: Optional[str] = None
    bid_ivs: Optional[Sequence[np.ndarray]] = None
    ask_ivs: Optional[Sequence[np.ndarray]] = None

    def __post_init__(self):
        self.ttms = np.asarray(self.ttms, dtype=float)
        self.forwards = np.asarray(self.forwards, dtype=float)
        ttms, forwards = self.ttms, self.forwards
        if ttms.ndim != 1 or ttms.size == 0:
            raise ValueError("ttms must be a non-empty one-dimensional array")
        if not np.all(np.isfinite(ttms)) or np.any(ttms <= 0.0):
            raise ValueError("ttms must contain only finite positive values")
        if np.any(np.diff(ttms) <= 0.0):
            raise ValueError("ttms must be strictly increasing")
        if forwards.ndim != 1 or forwards.size != ttms.size:
            raise ValueError("ttms and forwards must have the same one-dimensional length")
        if not np.all(np.isfinite(forwards)) or np.any(forwards <= 0.0):
            raise ValueError("forwards must contain only finite positive values")
        if len(self.strikes_ttms) != ttms.size or len(self.optiontypes_ttms) != ttms.size:
            raise ValueError("ttms, strikes_ttms, and optiontypes_ttms must have the same length")
        if self.ids is not None and len(self.ids) != ttms.size:
            raise ValueError("ids and ttms must have the same length")
        if self.discfactors is not None:
            self.discfactors = np.asarray(self.discfactors, dtype=float)
            if self.discfactors.shape != ttms.shape or np.any(self.discfactors <= 0.0) or not np.all(np.isfinite(self.discfactors)):
                raise ValueError("discfactors must contain only finite positive values")
            self.discount_rates = -np.log(self.discfactors) / ttms
        elif self.discount_rates is not None:
            self.discount_rates = np.asarray(self.discount_rates, dtype=float)
            self.discfactors = np.exp(-self.discount_rates * ttms)
        else:
            self.discfactors = np.ones_like(ttms)
            self.discount_rates = np.zeros_like(ttms)
        for strikes, optiontypes in zip(self.strikes_ttms, self.optiontypes_ttms):
            _validate_option_slice_data(strikes, optiontypes)

    @classmethod
    def slice_to_chain(cls, ttm: float, forward: float, strikes: np.ndarray, optiontypes: np.ndarray,
                       discfactor: float = 1.0, id: Optional[str] = None) -> "OptionChain":
        return cls(ttms=np.array([ttm]), forwards=np.array([forward]), strikes_ttms=(np.asarray(strikes, dtype=float),),
                   optiontypes_ttms=(np.asarray(optiontypes),), discfactors=np.array([discfactor]),
                   ids=np.array([id]) if id is not None else np.array([f"{ttm:0.2f}"]))

    @classmethod
    def get_uniform_chain(cls, ttms=np.array([0.083, 0.25]), ids=np.array(["1m", "3m"]), forwards=np.array([1.0, 1.0]),
                          strikes=np.linspace(0.9, 1.1, 3), flat_vol: float = 0.2) -> "OptionChain":
        ttms = np.asarray(ttms, dtype=float)
        forwards = np.asarray(forwards, dtype=float)
        if forwards.ndim == 1 and forwards.size != ttms.size and forwards.size > 0 and np.all(forwards == forwards[0]):
            forwards = np.full(ttms.size, forwards[0], dtype=float)
        strikes = np.asarray(strikes, dtype=float)
        return cls(ttms=ttms, ids=ids, forwards=forwards, strikes_ttms=[strikes for _ in ttms],
                   bid_ivs=[flat_vol * np.ones_like(strikes) for _ in ttms], ask_ivs=[flat_vol * np.ones_like(strikes) for _ in ttms],
                   optiontypes_ttms=[np.where(strikes >= forward, "C", "P") for forward in forwards])

    def get_mid_vols(self) -> Optional[List[np.ndarray]]:
        if self.bid_ivs is not None and self.ask_ivs is not None:
            return [0.5 * (b + a) for b, a in zip(self.bid_ivs, self.ask_ivs)]
        return None

    def get_chain_data_as_xy(self) -> Tuple[tuple, List[np.ndarray]]:
        """(x, y) for model calibration: chain inputs and mid implied vols (reference data/option_chain.py:318-325)."""
        mid_vols = [0.5 * (b + a) for b, a in zip(self.bid_ivs, self.ask_ivs)]
        return (self.ttms, self.forwards, self.discfactors, self.strikes_ttms, self.optiontypes_ttms), mid_vols

    def get_chain_atm_vols(self) -> np.ndarray:
        """mid vol of each slice interpolated to the forward (reference :281-286)."""
        return np.array([np.interp(x=f, xp=k, fp=y) for f, k, y in zip(self.forwards, self.strikes_ttms, self.get_mid_vols())])

    def get_chain_vegas(self, is_unit_ttm_vega: bool = False) -> List[np.ndarray]:
        """Black-76 vegas F*n(d1)*sqrt(T) at the mid vols, one array per slice -- the calibration weights of Eq. (6.3) (reference
        :263-279 -> third-party ``vanilla_option_pricers.bsm.compute_bsm_vegas_ttms``, absent from the reference tree; host-side
        arithmetic on a handful of quotes, normalised per slice by the caller)."""
        ttms = np.ones_like(self.ttms) if is_unit_ttm_vega else self.ttms
        out = []
        for ttm, forward, strikes, vols in zip(ttms, self.forwards, self.strikes_ttms, self.get_mid_vols()):
            sdev = vols * np.sqrt(ttm)
            d1 = np.log(forward / strikes) / sdev + 0.5 * sdev
            out.append(forward * np.exp(-0.5 * d1 * d1) / np.sqrt(2.0 * np.pi) * np.sqrt(ttm))
        return out

    def get_slice_varswap_strikes(self, floor_with_atm_vols: bool = True):
        """variance-swap VOL per maturity by static replication from the quoted strip (reference data/option_chain.py:402-426 with
        utils/var_swap_pricer.py:8-56): K_var = (2/T) sum_i dk_i O(K_i)/K_i^2 - (F/K_atm - 1)^2 / T with Black mid prices O (puts below
        the forward, calls above), centred strike spacings, floored with the ATM vol.  Host arithmetic; returns a pandas Series."""
        import pandas as pd
        from scipy.special import ndtr
        out = np.zeros_like(self.ttms)
        for m, (ttm, forward, strikes, vols, types) in enumerate(zip(self.ttms, self.forwards, self.strikes_ttms, self.get_mid_vols(),
                                                                      self.optiontypes_ttms)):
            sdev = vols * np.sqrt(ttm)
            d1 = np.log(forward / strikes) / sdev + 0.5 * sdev
            calls = forward * ndtr(d1) - strikes * ndtr(d1 - sdev)
            prices = np.where(np.asarray(types) == "P", calls - (forward - strikes), calls)
            is_put = np.asarray(types) == "P"
            # the reference joins the put and call strips on strike and picks puts below / calls at-or-above the forward
            table = pd.concat([pd.Series(prices[is_put], index=strikes[is_put], name="puts"),
                               pd.Series(prices[~is_put], index=strikes[~is_put], name="calls")], axis=1).sort_index()
            k = table.index.to_numpy()
            dk = np.empty_like(k)
            dk[0], dk[-1] = k[1] - k[0], k[-1] - k[-2]
            dk[1:-1] = 0.5 * (k[2:] - k[:-2])
            below = k < forward
            strip = np.where(below, table["puts"].to_numpy(), table["calls"].to_numpy())
            k_atm = k[~below][0]
            out[m] = np.sqrt((2.0 * np.nansum(dk * strip / np.square(k)) - np.square(forward / k_atm - 1.0)) / ttm)
        if floor_with_atm_vols:
            out = np.maximum(self.get_chain_atm_vols(), out)
        return pd.Series(out, index=self.ttms)

    def compute_model_ivols_from_chain_data(self, model_prices: Sequence[np.ndarray], forwards: np.ndarray = None) -> List[np.ndarray]:
        """invert model prices to Black implied vols for the whole chain in one GPU kernel launch (reference
        data/option_chain.py:327-346 -> third-party ``vanilla_option_pricers``)."""
        from .. import engine
        if forwards is None:
            forwards = self.forwards
        return engine.bsm_implied_vols(self.ttms, forwards, self.discfactors, self.strikes_ttms, self.optiontypes_ttms, model_prices)


def get_btc_test_chain_data() -> OptionChain:
    """BTC option chain of 21 Oct 2021 bundled with the reference (data/sample_option_chains.py:80-132): market data
    (maturities, forwards, listed strikes, put/call flags) used by BASELINE.json config "full BTC-style option chain"."""
    ttms = np.array([0.04289242541152263, 0.10122575874485597, 0.1984479809670782, 0.4317813143004115])
    forwards = np.array([67106.44399999999, 67843.219, 68689.48000000001, 70617.77892857141])
    strikes_ttms = (
        np.array([52000., 56000., 58000., 60000., 64000., 66000., 70000., 72000., 74000., 75000., 90000., 95000.]),
        np.array([45000., 48000., 55000., 58000., 64000., 65000., 70000., 75000., 80000., 85000., 90000., 100000., 120000.]),
        np.array([38000., 42000., 46000., 52000., 56000., 60000., 64000., 70000., 75000., 80000., 85000., 90000., 100000., 140000., 160000.]),
        np.array([35000., 40000., 60000., 80000., 100000., 120000., 150000., 250000., 300000.]))
    optiontypes_ttms = (np.array(["P"] * 6 + ["C"] * 6), np.array(["P"] * 6 + ["C"] * 7), np.array(["P"] * 7 + ["C"] * 8),
                        np.array(["P"] * 3 + ["C"] * 6))
    return OptionChain(ids=np.array(["2w", "1m", "2m", "3m"]), ttms=ttms, ticker="BTC", forwards=forwards,
                       strikes_ttms=strikes_ttms, optiontypes_ttms=optiontypes_ttms, discfactors=np.ones(4))
