"""INTEGRATION.md claims that the reference's OWN objects (stochvolmodels.OptionChain, LogSvParams, HestonParams, VariableType) can be handed
to the B200 pricers unchanged.  Checked here without a GPU: the C entry points are replaced by a recorder and the marshalled arguments of
calls made with reference objects must equal those made with this package's objects.  Needs /root/reference (absent on the GPU box ->
skipped there); the third-party packages the reference imports at module level are stubbed exactly as in tests/golden/make_golden.py."""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import pytest

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_pdf", "matplotlib.lines",
                 "matplotlib.ticker", "matplotlib.figure", "matplotlib.axes", "matplotlib.dates", "matplotlib.colors", "seaborn",
                 "vanilla_option_pricers", "vanilla_option_pricers.bsm", "vanilla_option_pricers.bachelier"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REF_SRC)
    try:
        from stochvolmodels.data.option_chain import OptionChain
        from stochvolmodels.pricers.heston_pricer import HestonParams
        from stochvolmodels.pricers.logsv.logsv_params import LogSvParams
        from stochvolmodels.utils.config import VariableType
    finally:
        sys.path.remove(REF_SRC)
    return dict(OptionChain=OptionChain, LogSvParams=LogSvParams, HestonParams=HestonParams, VariableType=VariableType)


class Recorder:
    """stands in for _capi.call: records (name, decoded args) and fills nothing (outputs stay uninitialised numpy buffers)."""
    def __init__(self):
        self.calls = []

    def __call__(self, name, *args):
        import ctypes
        decoded = []
        for a in args:
            if isinstance(a, (int, float, type(None))):
                decoded.append(a)
            elif isinstance(a, ctypes.Structure):
                decoded.append(tuple(getattr(a, f) for f, _ in a._fields_))
            elif hasattr(a, "_obj") and isinstance(a._obj, ctypes.Structure):          # byref(struct)
                decoded.append(tuple(getattr(a._obj, f) for f, _ in a._obj._fields_))
            else:
                decoded.append("ptr")
        self.calls.append((name, tuple(decoded)))


def _chains(ref):
    import stochvolmodels_b200 as svm
    kw = dict(ttms=np.array([0.1, 0.3]), forwards=np.array([1.0, 1.02]), discfactors=np.array([1.0, 0.99]), ids=np.array(["a", "b"]),
              strikes_ttms=[np.array([0.9, 1.0, 1.1]), np.array([0.8, 1.2])], optiontypes_ttms=[np.array(["P", "C", "C"]), np.array(["P", "C"])])
    return ref["OptionChain"](**kw), svm.OptionChain(**kw)


def test_reference_objects_marshal_like_our_own(ref, monkeypatch):
    import stochvolmodels_b200 as svm
    from stochvolmodels_b200 import _capi as C, engine
    flat = []
    real_flatten = C.flatten_chain
    monkeypatch.setattr(C, "flatten_chain", lambda s, t: flat.append(real_flatten(s, t)) or flat[-1])
    rec = Recorder()
    monkeypatch.setattr(C, "call", rec)
    ref_chain, our_chain = _chains(ref)
    p6 = dict(sigma0=0.9, theta=1.0, kappa1=4.0, kappa2=3.0, beta=0.3, volvol=1.5)
    ref_params, our_params = ref["LogSvParams"](**p6), svm.LogSvParams(**p6)
    pricer = svm.LogSVPricer()
    for chain, params, vt in ((ref_chain, ref_params, ref["VariableType"].LOG_RETURN), (our_chain, our_params, svm.VariableType.LOG_RETURN)):
        pricer.price_chain(chain, params)
        pricer.model_mc_price_chain(chain, params, nb_path=1000, nb_steps=360, seed=3, variable_type=vt, distributed=False)
    names = [c[0] for c in rec.calls]
    assert names == ["b200sv_logsv_price_chain", "b200sv_logsv_mc_chain"] * 2
    assert rec.calls[0] == rec.calls[2] and rec.calls[1] == rec.calls[3]            # identical scalars / structs for both object families
    assert rec.calls[0][1][0] == (0.9, 1.0, 4.0, 3.0, 0.3, 1.5)
    for a, b in zip(flat[:2], flat[2:]):                                            # identical flattened strikes / type codes / offsets
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)
    # kappa2=None of the reference's PARAMS5 convention and its vol backbone lookup go through unchanged as well
    q = ref["LogSvParams"](sigma0=0.9, theta=0.8, kappa1=4.0, kappa2=None, beta=0.3, volvol=1.5)
    rec.calls.clear()
    pricer.price_chain(ref_chain, q)
    assert rec.calls[0][1][0][3] == 4.0 / 0.8


def test_reference_heston_params_and_enum_values(ref, monkeypatch):
    import stochvolmodels_b200 as svm
    from stochvolmodels_b200 import _capi as C, engine
    rec = Recorder()
    monkeypatch.setattr(C, "call", rec)
    ref_chain, _ = _chains(ref)
    hp = ref["HestonParams"](v0=0.05, theta=0.06, kappa=3.0, rho=-0.4, volvol=0.7)
    svm.HestonPricer().price_chain(ref_chain, hp)
    svm.HestonPricer().model_mc_price_chain(ref_chain, hp, nb_path=1000, seed=1, variable_type=ref["VariableType"].Q_VAR, distributed=False)
    assert rec.calls[0][0] == "b200sv_heston_price_chain" and rec.calls[0][1][0] == (0.05, 0.06, 3.0, -0.4, 0.7)
    assert rec.calls[1][0] == "b200sv_heston_mc_chain"
    assert engine.variable_code(ref["VariableType"].Q_VAR) == C.Q_VAR and engine.variable_code(ref["VariableType"].LOG_RETURN) == C.LOG_RETURN
    with pytest.raises(NotImplementedError):
        engine.variable_code(ref["VariableType"].SIGMA)


def test_reference_hawkes_params_marshal_like_our_own(ref, monkeypatch):
    """the reference's own HawkesJDParams (and OptionChain) through HawkesJDPricer: Monte Carlo and Fourier routes, with and without the risk
    kernel -- same C calls, same sixteen floats, same gamma"""
    import stochvolmodels_b200 as svm
    from stochvolmodels_b200 import _capi as C
    sys.path.insert(0, REF_SRC)
    try:
        from stochvolmodels.pricers.hawkes_jd_pricer import HawkesJDParams as RefHawkes
    finally:
        sys.path.remove(REF_SRC)
    rec = Recorder()
    monkeypatch.setattr(C, "call", rec)
    ref_chain, our_chain = _chains(ref)
    kw = dict(sigma=0.4, mean_p=0.04, mean_m=-0.05, theta_p=6.0, lambda_p=7.0, kappa_p=20.0, beta1_p=30.0, beta2_p=-25.0)
    pricer = svm.HawkesJDPricer()
    for chain, cls in ((ref_chain, RefHawkes), (our_chain, svm.HawkesJDParams)):
        pricer.model_mc_price_chain(chain, cls(**kw), nb_path=1000, seed=4, distributed=False)
        pricer.price_chain(chain, cls(**kw))
        pricer.price_chain(chain, cls(risk_premia_gamma=0.3, **kw))
    names = [c[0] for c in rec.calls]
    assert names == ["b200sv_hawkesjd_mc_chain", "b200sv_hawkesjd_price_chain", "b200sv_hawkesjd_price_chain"] * 2
    for a, b in zip(rec.calls[:3], rec.calls[3:]):
        assert a == b or all(x == y or (x != x and y != y) for x, y in zip(a[1], b[1]))        # NaN = "no risk kernel" compares by identity
    sixteen = rec.calls[0][1][0]
    assert len(sixteen) == 16 and sixteen[1] == 0.4 and sixteen[6] == 7.0
    gammas = [c[1][11] for c in rec.calls if c[0] == "b200sv_hawkesjd_price_chain"]
    assert gammas[0] != gammas[0] and gammas[1] == 0.3 and gammas[3] == 0.3
