import numpy as np, sys
sys.path.insert(0, '/root/repo')
from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, LogSvParams, LogSVPricer, get_btc_test_chain_data
chain = get_btc_test_chain_data()
p = LogSvParams(sigma0=0.377, theta=0.347, kappa1=1.29, kappa2=1.93, beta=2.45, volvol=1.81, H=0.1,
                weights=np.array([0.77718798, 1.5545139, 8.51550426]), nodes=np.array([7.71995736e-02, 5.19154220e+00, 1.08459557e+02]))
for _ in range(2):
    LogSVPricer().model_mc_price_chain(chain, p, nb_path=4_000_000, nb_steps=360, use_rough_mc=True, seed=1, gauss="fp32")
    HawkesJDPricer().model_mc_price_chain(chain, HawkesJDParams(), nb_path=4_000_000, seed=1)
