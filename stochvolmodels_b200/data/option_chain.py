"""OptionChain: the ragged chain container that crosses the Pricer boundary.

Host-side data type only ("boundary input type").  It keeps the attributes the hot path reads --
``ttms, forwards, discfactors, strikes_ttms, optiontypes_ttms`` (reference pricers/logsv_pricer.py:358-365, 412-427) -- with the
reference's validation rules (data/option_chain.py:147-215) and constructors ``slice_to_chain`` (:230-245) and
``get_uniform_chain`` (:462-492), the slice views ``OptionSlice`` / ``get_slice`` / ``get_slices_as_chain`` (:72-124, 387-459) and the
strike transforms ``to_forward_normalised_strikes`` / ``to_uniform_strikes`` (:348-385).  The pricers are duck-typed: a reference
``stochvolmodels.OptionChain`` works as well.
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


def _check_quotes(name: str, values, size: int, strictly_positive: bool) -> np.ndarray:
    """per-slice bid / ask arrays: finite, aligned with the strikes, positive (vols) or non-negative (prices) -- option_chain.py:33-51"""
    v = np.asarray(values, dtype=float)
    if v.shape != (size,):
        raise ValueError(f"{name} must have the same length as strikes")
    if not np.isfinite(v).all():
        raise ValueError(f"{name} must contain only finite values")
    if strictly_positive and (v <= 0.0).any():
        raise ValueError(f"{name} must contain only positive values")
    if not strictly_positive and (v < 0.0).any():
        raise ValueError(f"{name} must contain only non-negative values")
    return v


@dataclass
class OptionSlice:
    """one maturity of a chain (reference data/option_chain.py:72-124): discount factor and rate are kept consistent, quotes validated"""
    ttm: float
    forward: float
    strikes: np.ndarray
    optiontypes: np.ndarray
    id: str
    discfactor: Optional[float] = None
    discount_rate: Optional[float] = None
    bid_ivs: Optional[np.ndarray] = None
    ask_ivs: Optional[np.ndarray] = None
    bid_prices: Optional[np.ndarray] = None
    ask_prices: Optional[np.ndarray] = None

    def __post_init__(self):
        for name, value in (("ttm", self.ttm), ("forward", self.forward)):
            if not np.isscalar(value) or not np.isfinite(value) or value <= 0.0:
                raise ValueError(f"{name} must be a finite positive scalar")
        size = _validate_option_slice_data(self.strikes, self.optiontypes)
        if self.discfactor is not None:
            if not np.isfinite(self.discfactor) or self.discfactor <= 0.0:
                raise ValueError("discfactor must be a finite positive scalar")
            self.discount_rate = -np.log(self.discfactor) / self.ttm
        elif self.discount_rate is not None:
            if not np.isfinite(self.discount_rate):
                raise ValueError("discount_rate must be a finite scalar")
            self.discfactor = np.exp(-self.discount_rate * self.ttm)
        else:
            self.discfactor, self.discount_rate = 1.0, 0.0
        quotes = {name: _check_quotes(name, getattr(self, name), size, name.endswith("ivs"))
                  for name in ("bid_ivs", "ask_ivs", "bid_prices", "ask_prices") if getattr(self, name) is not None}
        for bid, ask in (("bid_ivs", "ask_ivs"), ("bid_prices", "ask_prices")):
            if bid in quotes and ask in quotes and (quotes[bid] > quotes[ask]).any():
                raise ValueError(f"{bid} must not exceed {ask}")


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
    bid_prices: Optional[Sequence[np.ndarray]] = None
    ask_prices: Optional[Sequence[np.ndarray]] = None
    forwards0: Optional[np.ndarray] = None          # the forwards a forward-normalised chain was divided by

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
        if self.forwards0 is not None:
            _check_quotes("forwards0", self.forwards0, ttms.size, True)
        quoted = [q for q in ("bid_ivs", "ask_ivs", "bid_prices", "ask_prices") if getattr(self, q) is not None]
        for q in quoted:
            if len(getattr(self, q)) != ttms.size:
                raise ValueError(f"{q} and ttms must have the same length")
        for m, (strikes, optiontypes) in enumerate(zip(self.strikes_ttms, self.optiontypes_ttms)):
            size = _validate_option_slice_data(strikes, optiontypes)
            ok = {q: _check_quotes(f"{q}[{m}]", getattr(self, q)[m], size, q.endswith("ivs")) for q in quoted}     # reference :187-215
            for bid, ask in (("bid_ivs", "ask_ivs"), ("bid_prices", "ask_prices")):
                if bid in ok and ask in ok and (ok[bid] > ok[ask]).any():
                    raise ValueError(f"{bid}[{m}] must not exceed {ask}[{m}]")

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

    # ---- slice views and strike transforms (reference :217-227, 348-459) -------------------------------------------------------
    def _per_slice(self, name: str, idx: int):
        values = getattr(self, name)
        return None if values is None else values[idx]

    def _index_of(self, id) -> int:
        if self.ids is None:
            raise ValueError("the chain has no ids")
        return list(self.ids).index(id)

    def print(self) -> None:
        for name in ("ttms", "forwards", "strikes_ttms", "optiontypes_ttms", "ids", "bid_ivs", "ask_ivs"):
            print(f"{name}:\n{getattr(self, name)}")

    def get_slice(self, id: str) -> OptionSlice:
        k = self._index_of(id)
        return OptionSlice(id=self.ids[k], ttm=float(self.ttms[k]), forward=float(self.forwards[k]), strikes=self.strikes_ttms[k],
                           optiontypes=self.optiontypes_ttms[k], discfactor=float(self.discfactors[k]),
                           **{q: self._per_slice(q, k) for q in ("bid_ivs", "ask_ivs", "bid_prices", "ask_prices")})

    @classmethod
    def get_slices_as_chain(cls, option_chain: "OptionChain", ids: Sequence[str]) -> "OptionChain":
        """the sub-chain of the maturities named in ``ids`` -- in CHAIN order when several are asked for (``np.isin`` semantics of the
        reference), labelled with ``ids`` as given"""
        if len(ids) == 1:
            keep = [option_chain._index_of(ids[0])]
        else:
            keep = list(np.flatnonzero(np.isin(option_chain.ids, ids)))
        pick = lambda name: None if getattr(option_chain, name) is None else [getattr(option_chain, name)[k] for k in keep]
        return cls(ids=ids, ttms=option_chain.ttms[keep], ticker=option_chain.ticker, forwards=option_chain.forwards[keep],
                   strikes_ttms=pick("strikes_ttms"), optiontypes_ttms=pick("optiontypes_ttms"), discfactors=option_chain.discfactors[keep],
                   bid_ivs=pick("bid_ivs"), ask_ivs=pick("ask_ivs"), bid_prices=pick("bid_prices"), ask_prices=pick("ask_prices"))

    @classmethod
    def to_forward_normalised_strikes(cls, obj: "OptionChain") -> "OptionChain":
        """strikes divided by their forward, forwards set to 1 (the originals kept in ``forwards0``); quotes in vols are unchanged"""
        return cls(ttms=obj.ttms, forwards=np.ones_like(obj.forwards), strikes_ttms=[k / f for k, f in zip(obj.strikes_ttms, obj.forwards)],
                   optiontypes_ttms=obj.optiontypes_ttms, discfactors=obj.discfactors, ticker=obj.ticker, ids=obj.ids, bid_ivs=obj.bid_ivs,
                   ask_ivs=obj.ask_ivs, forwards0=obj.forwards)

    @classmethod
    def to_uniform_strikes(cls, obj: "OptionChain", num_strikes: int = 21) -> "OptionChain":
        """``num_strikes`` equally spaced strikes between the first and last strike of every slice, calls at and above the forward, puts
        below; the quotes no longer apply and are dropped"""
        grids = [np.linspace(k[0], k[-1], num_strikes) for k in obj.strikes_ttms]
        return cls(ttms=obj.ttms, forwards=obj.forwards, strikes_ttms=grids,
                   optiontypes_ttms=[np.where(g >= f, "C", "P") for g, f in zip(grids, obj.forwards)], discfactors=obj.discfactors,
                   ticker=obj.ticker, ids=obj.ids, bid_ivs=None, ask_ivs=None)

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

    def get_chain_deltas(self) -> List[np.ndarray]:
        """Black-76 forward deltas at the mid vols, N(d1) for calls and N(d1) - 1 for puts (inverse options take the vanilla delta of their
        payoff side) -- reference :254-261 -> third-party ``vanilla_option_pricers.bsm.compute_bsm_vanilla_deltas_ttms``, absent from the
        reference tree: the convention is the textbook one, its bit-level parity is unpinned like the Black inversion's."""
        from scipy.special import ndtr
        out = []
        for ttm, forward, strikes, types, vols in zip(self.ttms, self.forwards, self.strikes_ttms, self.optiontypes_ttms, self.get_mid_vols()):
            sdev = vols * np.sqrt(ttm)
            n_d1 = ndtr(np.log(forward / strikes) / sdev + 0.5 * sdev)
            out.append(np.where(np.isin(types, ("C", "IC")), n_d1, n_d1 - 1.0))
        return out

    def get_chain_skews(self, delta: float = 0.25) -> np.ndarray:
        """(vol of the ``delta`` put - vol of the ``delta`` call) / ATM vol per slice, each read off the slice's own quotes by linear
        interpolation in delta (reference :288-316); a slice needs quotes on both sides"""
        skews = np.zeros(self.ttms.size)
        for m, (deltas, vols, types, atm) in enumerate(zip(self.get_chain_deltas(), self.get_mid_vols(), self.optiontypes_ttms,
                                                           self.get_chain_atm_vols())):
            sides = []
            for flag, target in (("P", -delta), ("C", delta)):
                pick = np.asarray(types) == flag
                if not pick.any():
                    raise ValueError("skew interpolation requires both put and call quotes")
                order = np.argsort(deltas[pick])
                sides.append(np.interp(target, deltas[pick][order], vols[pick][order]))
            skews[m] = (sides[0] - sides[1]) / atm
        return skews

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
    # quoted implied vols of the same snapshot (market data): what the calibration examples of the reference fit to
    bid_ivs = (
        np.array([0.9231, 0.8835, 0.8695, 0.8621, 0.855, 0.8589, 0.8822, 0.8856, 0.8944, 0.8996, 0.9607, 0.9718]),
        np.array([0.9475, 0.9211, 0.8917, 0.8863, 0.8873, 0.8913, 0.9012, 0.9102, 0.9244, 0.9377, 0.9494, 0.9755, 1.0317]),
        np.array([0.9882, 0.9595, 0.94, 0.9247, 0.914, 0.9131, 0.9109, 0.9168, 0.9244, 0.9305, 0.941, 0.9544, 0.9682, 1.047, 1.0887]),
        np.array([1.0052, 0.981, 0.9593, 0.9722, 0.9924, 1.013, 1.0419, 1.1222, 1.1489]))
    ask_ivs = (
        np.array([0.9399, 0.8944, 0.8967, 0.8856, 0.8744, 0.8774, 0.9006, 0.9047, 0.9144, 0.9202, 0.9762, 1.0151]),
        np.array([0.972, 0.9401, 0.9092, 0.9014, 0.9041, 0.9079, 0.917, 0.9264, 0.9417, 0.952, 0.9602, 0.9899, 1.0585]),
        np.array([1.0167, 0.9739, 0.9568, 0.9372, 0.9285, 0.9261, 0.9261, 0.9225, 0.9358, 0.9362, 0.9558, 0.9637, 0.9823, 1.0664, 1.1144]),
        np.array([1.0204, 0.9968, 0.9683, 0.976, 0.9963, 1.0173, 1.047, 1.1411, 1.1736]))
    return OptionChain(ids=np.array(["2w", "1m", "2m", "3m"]), ttms=ttms, ticker="BTC", forwards=forwards,
                       strikes_ttms=strikes_ttms, optiontypes_ttms=optiontypes_ttms, discfactors=np.ones(4), bid_ivs=bid_ivs, ask_ivs=ask_ivs)
