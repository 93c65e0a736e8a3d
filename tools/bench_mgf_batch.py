"""Batched LogSV Fourier chain pricer (b200sv_logsv_price_chain_batch): ms per call against batch size and perturbation width.
python tools/bench_mgf_batch.py       (B200SV_MGF_SET_MAJOR=0 switches the set-major thread mapping off for an A/B)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stochvolmodels_b200 import LOGSV_BTC_PARAMS as P, engine, get_btc_test_chain_data

chain = get_btc_test_chain_data()
base = np.array([P.sigma0, P.theta, P.kappa1, P.kappa2, P.beta, P.volvol])
print(f"# lib {os.environ.get('B200SV_LIB', 'default')}  set_major {os.environ.get('B200SV_MGF_SET_MAJOR', '1')}")
print("width  |  B sets | ms per call | us per set")
for width in (0.05, 1e-6):            # 5 % = a global search population; 1e-6 = the finite-difference gradient of one SLSQP iteration
    for B in (7, 32, 128, 512, 2048):
        rng = np.random.RandomState(0)
        sets = [engine.logsv_params_c(*(base * (1 + width * rng.uniform(-1, 1, 6)))) for _ in range(B)]
        call = lambda: engine.logsv_price_chain_batch(sets, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms,
                                                      chain.optiontypes_ttms, vol_scaler=0.17)
        for _ in range(3):
            call()
        ts = []
        for _ in range(10 if B <= 512 else 4):
            torch.cuda.synchronize(); t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
        ms = 1e3 * float(np.median(ts))
        print(f"{width:6.0e} | {B:6d} | {ms:9.3f} | {1e3 * ms / B:9.1f}")
