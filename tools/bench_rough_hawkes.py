"""Throughput of the two 'next' Monte Carlo routes through the public API (host buffers in and out; median of 5 calls):
rough-LogSV multi-factor MC (3 factors, BTC chain, in-kernel Philox draws and uploaded RandomState normals) and Hawkes jump-diffusion MC.
    python tools/bench_rough_hawkes.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, LogSvParams, LogSVPricer, get_btc_test_chain_data
from stochvolmodels_b200.pricers.logsv_pricer import get_randoms_for_rough_vol_chain_valuation, rough_logsv_mc_chain_pricer_fixed_randoms
from stochvolmodels_b200.utils.funcs import set_time_grid


def med(f, n=5):
    f()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
    return float(np.median(ts))


chain = get_btc_test_chain_data()
p = LogSvParams(sigma0=0.377, theta=0.347, kappa1=1.29, kappa2=1.93, beta=2.45, volvol=1.81, H=0.1,
                weights=np.array([0.77718798, 1.5545139, 8.51550426]), nodes=np.array([7.71995736e-02, 5.19154220e+00, 1.08459557e+02]))
npy = 360
steps = sum(set_time_grid(t, npy)[0] for t in chain.ttms)          # every maturity restarts at t = 0
pricer = LogSVPricer()
print(f"# rough-LogSV, 3 factors, BTC chain 4 x 49, {npy} steps/yr: {steps} path-steps per path (maturities restart at t = 0)")
for n in (100_000, 1_000_000, 10_000_000):
    s = med(lambda: pricer.model_mc_price_chain(chain, p, nb_path=n, nb_steps=npy, use_rough_mc=True, seed=1, gauss="fp32"))
    print(f"  in-kernel Philox draws   {n:>9d} paths: {1e3 * s:9.3f} ms  {n * steps / s:.3e} path-steps/s")
for n in (100_000, 1_000_000):
    Z0, Z1, grids = get_randoms_for_rough_vol_chain_valuation(chain.ttms, n, npy, 10)
    f = lambda: rough_logsv_mc_chain_pricer_fixed_randoms(ttms=chain.ttms, forwards=chain.forwards, discfactors=chain.discfactors, strikes_ttms=chain.strikes_ttms,
                                                          optiontypes_ttms=chain.optiontypes_ttms, Z0=Z0, Z1=Z1, sigma0=p.sigma0, theta=p.theta, kappa1=p.kappa1,
                                                          kappa2=p.kappa2, beta=p.beta, orthog_vol=p.volvol, weights=p.weights, nodes=p.nodes, timegrids=grids)
    s = med(f, 3)
    print(f"  uploaded normals ({Z0.nbytes * 2 / 1e6:7.0f} MB) {n:>9d} paths: {1e3 * s:9.3f} ms  {n * steps / s:.3e} path-steps/s")
hp, hpr = HawkesJDParams(), HawkesJDPricer()
hsteps = sum(set_time_grid(t - t0, 1800)[0] for t, t0 in zip(chain.ttms, np.concatenate([[0.0], chain.ttms[:-1]])))
print(f"# Hawkes jump-diffusion, BTC chain 4 x 49, 1800 steps/yr: {hsteps} steps per path")
for n in (100_000, 1_000_000, 10_000_000):
    s = med(lambda: hpr.model_mc_price_chain(chain, hp, nb_path=n, seed=1))
    print(f"  {n:>9d} paths: {1e3 * s:9.3f} ms  {n * hsteps / s:.3e} path-steps/s")
