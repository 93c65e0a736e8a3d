import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from stochvolmodels_b200 import LOGSV_BTC_PARAMS, LogSvParams, LogSVPricer, LogsvModelCalibrationType, get_btc_test_chain_data
chain = get_btc_test_chain_data()
pricer = LogSVPricer()
for ctype in (LogsvModelCalibrationType.PARAMS4, LogsvModelCalibrationType.PARAMS5):
    t = time.perf_counter()
    fit, info = pricer.calibrate_model_params_to_chain(chain, LOGSV_BTC_PARAMS, model_calibration_type=ctype, return_info=True)
    print(ctype.name, f"{(time.perf_counter()-t)*1e3:.1f} ms", fit.to_str(), info["fun"], info["nit"])
    start = LogSvParams(sigma0=0.7, theta=0.9, kappa1=LOGSV_BTC_PARAMS.kappa1, kappa2=LOGSV_BTC_PARAMS.kappa2, beta=0.0, volvol=1.4)
    fit2, info2 = pricer.calibrate_model_params_to_chain(chain, start, model_calibration_type=ctype, return_info=True)
    print("  from a perturbed start:", fit2.to_str(), info2["fun"], info2["nit"])
ivs = pricer.compute_model_ivols_for_chain(chain, LOGSV_BTC_PARAMS)
mid = chain.get_mid_vols()
print("rms (model at LOGSV_BTC_PARAMS - mid):", np.sqrt(np.mean(np.concatenate([a - b for a, b in zip(ivs, mid)]) ** 2)))
