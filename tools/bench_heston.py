"""BASELINE.json configs[1]: Heston MC, 1e6 paths x 252 steps, single maturity, floor-Euler (reference scheme) and QE (opt-in), plus 1e8 paths.
   python tools/bench_heston.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stochvolmodels_b200 import HestonParams, HestonPricer, OptionChain

params = HestonParams(v0=0.04, theta=0.04, kappa=4.0, rho=-0.5, volvol=0.4)
K = np.array([0.8, 0.9, 1.0, 1.1, 1.2])
chain = OptionChain(ttms=np.array([1.0]), forwards=np.array([1.0]), strikes_ttms=[K], optiontypes_ttms=[np.array(["P", "P", "C", "C", "C"])],
                    ids=np.array(["1y"]))
pricer = HestonPricer()
fourier = pricer.price_chain(chain, params)[0]
print("# Heston MC, single 1y maturity, nb_steps_per_year=251 -> 252 steps; public API, host buffers; median of 10 calls")
for n in (1_000_000, 100_000_000):
    for scheme in ("euler_floor", "qe"):
        f = lambda seed: pricer.model_mc_price_chain(chain, params, nb_path=n, nb_steps_per_year=251, seed=seed, scheme=scheme)
        f(1); f(2)
        ts = []
        for r in range(10 if n < 1e8 else 3):
            t = time.perf_counter(); p, se = f(10 + r); ts.append(time.perf_counter() - t)
        med = float(np.median(ts))
        z = (p[0] - fourier) / se[0]
        print(f"{n:>10d} paths, {scheme:11s}: {1e3 * med:8.3f} ms  {n * 252 / med:.3e} path-steps/s   (MC - Fourier)/SE {np.round(z, 2)}  max|err| {np.max(np.abs(p[0] - fourier)):.1e}")
