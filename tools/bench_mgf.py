"""Wall-clock of the Fourier route through the public API (host buffers in/out) on BASELINE config 3 (5 maturities x 21 strikes)
and on the BTC chain, LogSV and Heston, with parity against the golden reference prices (reference CPU timings: SURVEY.md §6 and
profiles/r01_mgf_bench.txt).  usage: python tools/bench_mgf.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stochvolmodels_b200 import HestonParams, HestonPricer, LogSvParams, LogSVPricer, LOGSV_BTC_PARAMS, OptionChain, get_btc_test_chain_data, _capi

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
K21 = np.linspace(0.5, 1.5, 21); T21 = np.where(K21 >= 1.0, "C", "P")
ttms5 = np.array([1.0 / 12.0, 0.25, 0.5, 0.75, 1.0])
c3 = OptionChain(ttms=ttms5, forwards=np.ones(5), strikes_ttms=[K21] * 5, optiontypes_ttms=[T21] * 5)
btc = get_btc_test_chain_data()
Q = LogSvParams(1.0, 1.0, 5.0, 5.0, 0.2, 2.0)
lib = _capi.load_library()

def timed(f, reps=20):
    f(); f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); out = f(); ts.append(time.perf_counter() - t)
    return out, 1e3 * float(np.median(ts)), 1e3 * float(np.min(ts))

for name, pricer, params, chain, gold in (("logsv c3 5x21", LogSVPricer(), Q, c3, "logsv_fourier_c3_5x21.npz"),
                                          ("logsv btc 4x49", LogSVPricer(), LOGSV_BTC_PARAMS, btc, "logsv_fourier_btc.npz"),
                                          ("heston c3 5x21", HestonPricer(), HestonParams(), c3, "heston_fourier_c3_5x21.npz")):
    g = np.load(os.path.join(G, gold))
    lib.b200sv_reset_launch_count()
    prices, med, mn = timed(lambda: pricer.price_chain(chain, params))
    M = len(chain.ttms)
    rel = max(np.max(np.abs(prices[m] - g[f"prices_{m}"]) / np.maximum(np.abs(g[f"prices_{m}"]), 1e-12 * chain.forwards[m])) for m in range(M))
    mask_rel = max(np.max(np.abs(prices[m] / g[f"prices_{m}"] - 1)[g[f"prices_{m}"] > 1e-6 * chain.forwards[m]]) for m in range(M))
    print(f"{name:16s} GPU e2e median {med:8.3f} ms (min {mn:.3f})  max rel err vs reference golden {mask_rel:.2e}", flush=True)

# ---- options on quadratic variance, psi grid P = 40000 (reference: ~80 s per maturity on the CPU)
from stochvolmodels_b200 import VariableType
from stochvolmodels_b200.pricers.logsv_pricer import logsv_chain_pricer
g = np.load(os.path.join(G, "logsv_fourier_qvar_mma.npz"))
Kq, Tq = g["strikes"], np.array(["C"] * len(g["strikes"]))
prices, med, mn = timed(lambda: logsv_chain_pricer(Q, g["ttms"], np.ones(2), g["discfactors"], [Kq, Kq], [Tq, Tq], variable_type=VariableType.Q_VAR), reps=10)
rel = max(np.max(np.abs(prices[m] / g[f"prices_{m}"] - 1)) for m in range(2))
print(f"{'logsv Q_VAR 2x5':16s} GPU e2e median {med:8.3f} ms (min {mn:.3f})  max rel err vs reference golden {rel:.2e}  [2 maturities x 40000 RK45 solves]")

# ---- Hawkes jump-diffusion Fourier route (3 maturities x 500-point grid; reference: 3 x 500 solve_ivp calls, 7.9 s on this container's CPU)
from stochvolmodels_b200 import HawkesJDParams
from stochvolmodels_b200.pricers.hawkes_jd_pricer import hawkesjd_chain_pricer, hawkesjd_chain_pricer_with_risk_premia
g = np.load(os.path.join(G, "hawkes_fourier.npz"), allow_pickle=True)
Kh, Th = [g["strikes"] * f for f in g["forwards"]], [g["types"]] * 3
hp = HawkesJDParams(**dict(zip([str(k) for k in g["keys"]], g["dflt_params"])))
prices, med, mn = timed(lambda: hawkesjd_chain_pricer(hp, g["ttms"], g["forwards"], g["discfactors"], Kh, Th))
rel = max(np.max(np.abs(prices[m] / g["dflt_prices"][m] - 1)) for m in range(3))
print(f"{'hawkes 3x5':16s} GPU e2e median {med:8.3f} ms (min {mn:.3f})  max rel err vs reference golden {rel:.2e}")
hg = HawkesJDParams(**dict(zip([str(k) for k in g["keys"]], g["dflt_params"])), risk_premia_gamma=float(g["gamma"]))
prices, med, mn = timed(lambda: hawkesjd_chain_pricer_with_risk_premia(hg, g["ttms"], g["forwards"], g["discfactors"], Kh, Th))
rel = max(np.max(np.abs(prices[m] / g["gamma_prices"][m] - 1)) for m in range(3))
print(f"{'hawkes gamma 3x5':16s} GPU e2e median {med:8.3f} ms (min {mn:.3f})  max rel err vs reference golden {rel:.2e}  [risk kernel]")
