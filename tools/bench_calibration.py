"""Batched Fourier chain pricer + calibration timings on the BTC chain.   python tools/bench_calibration.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stochvolmodels_b200 import (LOGSV_BTC_PARAMS, LogSvParams, LogSVPricer, LogsvModelCalibrationType, OptionChain, engine,
                                 get_btc_test_chain_data)
from stochvolmodels_b200.pricers.logsv_pricer import _params_c

chain = get_btc_test_chain_data()
P = LOGSV_BTC_PARAMS
base = np.array([P.sigma0, P.theta, P.kappa1, P.kappa2, P.beta, P.volvol])
rng = np.random.RandomState(0)
J = sum(len(k) for k in chain.strikes_ttms)
print(f"# BTC chain: {len(chain.ttms)} maturities, {J} strikes; P=1000 grid points; times = median of 20 host-to-host calls")
print("B sets | ms per call | us per set | chain pricings / s")
for B in (1, 6, 32, 128, 512, 2048):
    sets = [engine.logsv_params_c(*(base * (1 + 0.05 * rng.uniform(-1, 1, 6)))) for _ in range(B)]
    call = lambda: engine.logsv_price_chain_batch(sets, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms,
                                                  chain.optiontypes_ttms, vol_scaler=0.17)
    for _ in range(10 if B <= 128 else 3):
        call()
    ts = []
    for _ in range(20 if B <= 512 else 5):
        torch.cuda.synchronize(); t = time.perf_counter(); call(); ts.append(time.perf_counter() - t)
    ms = 1e3 * float(np.median(ts))
    print(f"{B:6d} | {ms:9.3f} | {1e3 * ms / B:9.1f} | {B / ms * 1e3:12.0f}")

# whole calibration: synthetic market from the model itself, perturbed start
pricer = LogSVPricer()
flat_vols = [0.8 * np.ones_like(k) for k in chain.strikes_ttms]
mk = lambda vols: OptionChain(ttms=chain.ttms, forwards=chain.forwards, strikes_ttms=chain.strikes_ttms, optiontypes_ttms=chain.optiontypes_ttms,
                              discfactors=chain.discfactors, ids=chain.ids, bid_ivs=vols, ask_ivs=[v.copy() for v in vols])
c0 = mk(flat_vols)
market = pricer.compute_model_ivols_for_chain(c0, P, vol_scaler=pricer.set_vol_scaler(c0))
cm = mk(market)
start = LogSvParams(sigma0=0.7, theta=0.9, kappa1=P.kappa1, kappa2=P.kappa2, beta=0.0, volvol=1.4)
for ctype in (LogsvModelCalibrationType.PARAMS4, LogsvModelCalibrationType.PARAMS5):
    pricer.calibrate_model_params_to_chain(cm, start, model_calibration_type=ctype)          # warm-up
    t = time.perf_counter()
    fit, info = pricer.calibrate_model_params_to_chain(cm, start, model_calibration_type=ctype, return_info=True)
    dt = time.perf_counter() - t
    print(f"calibration {ctype.name}: {dt * 1e3:.1f} ms, {info['nit']} SLSQP iterations, {info['nb_batches']} batched GPU evaluations, "
          f"objective {info['fun']:.2e}; fit {fit.to_str()}")
# MC engine: numpy normals resident in HBM vs Philox common random numbers (one batched call per evaluation)
from stochvolmodels_b200 import CalibrationEngine
for rnd in ("numpy", "philox"):
    kw = dict(model_calibration_type=LogsvModelCalibrationType.PARAMS4, calibration_engine=CalibrationEngine.MC, nb_path=100000, nb_steps=360,
              seed=10, mc_randoms=rnd)
    pricer.calibrate_model_params_to_chain(cm, start, **kw)
    t = time.perf_counter()
    fit, info = pricer.calibrate_model_params_to_chain(cm, start, return_info=True, **kw)
    dt = time.perf_counter() - t
    print(f"calibration PARAMS4, MC engine 1e5 paths, {rnd} normals: {dt * 1e3:.1f} ms ({info['nit']} iterations, {info['nb_batches']} evaluations of 5 "
          f"chains, objective {info['fun']:.2e}); fit {fit.to_str()}")
print("# reference: one objective evaluation = one CPU chain pricing = 7.6 s on this chain (profiles/r01_mgf_bench.txt); SLSQP needs "
      "(n+1) per iteration")
