"""one batched LogSV Fourier chain call (B sets) for ncu:  ncu ... python tools/profile_mgf_batch.py 512"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stochvolmodels_b200 import LOGSV_BTC_PARAMS as P, engine, get_btc_test_chain_data
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chain = get_btc_test_chain_data()
base = np.array([P.sigma0, P.theta, P.kappa1, P.kappa2, P.beta, P.volvol])
rng = np.random.RandomState(0)
sets = [engine.logsv_params_c(*(base * (1 + 0.05 * rng.uniform(-1, 1, 6)))) for _ in range(B)]
for _ in range(3):
    engine.logsv_price_chain_batch(sets, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms, chain.optiontypes_ttms, vol_scaler=0.17)
