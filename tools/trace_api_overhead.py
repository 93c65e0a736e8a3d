#!/usr/bin/env python
"""Where does an API call of the sharded MC path spend its host time?  (N>1 e2e lost 6-13 ms per call.)

Runs on ONE GPU (the distributed driver works unsharded without a process group): the host-level C call, then
multi_gpu.mc_chain_distributed with the engine cache off / on, 1e8 paths x BTC chain, wall ms per call (median of 5)."""
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
    return statistics.median(ts)


def main():
    import torch
    from stochvolmodels_b200 import LOGSV_BTC_PARAMS, LogSVPricer, engine, get_btc_test_chain_data, multi_gpu
    from stochvolmodels_b200 import _capi as C
    from stochvolmodels_b200.pricers.logsv_pricer import _params_c
    chain, p = get_btc_test_chain_data(), LOGSV_BTC_PARAMS
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    flags = engine.mc_flags("fp64", "fp32")
    host = lambda: LogSVPricer().model_mc_price_chain(chain, p, nb_path=n, nb_steps=582, seed=10)
    dist = lambda: multi_gpu.mc_chain_distributed("logsv", _params_c(p), chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms,
                                                  chain.optiontypes_ttms, n, 582, True, C.LOG_RETURN, 10, flags)
    print(f"paths {n:.0e}")
    print(f"host-level C call (b200sv_logsv_mc_chain)        {med(host):9.3f} ms")
    os.environ["B200SV_ENGINE_CACHE"] = "0"
    print(f"mc_chain_distributed, engine cache OFF            {med(dist):9.3f} ms")
    torch.cuda.empty_cache()
    os.environ["B200SV_ENGINE_CACHE"] = "1"
    print(f"mc_chain_distributed, engine cache ON             {med(dist):9.3f} ms")
    a, _ = host()
    b, _ = dist()
    print("max rel diff host-level vs distributed driver:", float(np.max(np.abs(np.concatenate(a) / np.concatenate(b) - 1))))


if __name__ == "__main__":
    main()
