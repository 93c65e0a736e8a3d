"""Latency of the public API for calibration-sized calls (host buffers in/out): MC chain at small path counts, Fourier chain."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stochvolmodels_b200 import LOGSV_BTC_PARAMS, LogSVPricer, get_btc_test_chain_data, _capi

chain = get_btc_test_chain_data()
pricer = LogSVPricer()
lib = _capi.load_library()
for n in (10_000, 100_000, 1_000_000, 10_000_000):
    f = lambda: pricer.model_mc_price_chain(chain, LOGSV_BTC_PARAMS, nb_path=n, nb_steps=360, seed=1)
    f(); f()
    lib.b200sv_reset_launch_count()
    ts = []
    for _ in range(10):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    steps = n * sum(int((b - a) * 360) + 1 for a, b in zip(np.r_[0.0, chain.ttms[:-1]], chain.ttms))
    print(f"MC BTC chain nb_path={n:>9d} nb_steps=360: median {1e3 * np.median(ts):8.3f} ms  ({steps / np.median(ts):.3e} path-steps/s, {lib.b200sv_launch_count() // 10} launches/call)")
