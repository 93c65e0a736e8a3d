"""Do the float (SFU) Gaussian draws of the default mode bias the prices?  (the evidence was prose.)

Three checks on the bench workload (BTC chain, 49 strikes, 252 steps/path, LOGSV_BTC_PARAMS):

1. PAIRED: default mode vs ``gauss="fp64_paired"`` -- the SAME Philox words pushed through the fp64 Box-Muller (gauss64.cuh).  The two
   runs share every uniform, so their price difference has a tiny variance; it isolates what the SFU approximations (lg2 / sqrt /
   sin / cos .approx) and the float rounding of the uniforms do.  At 1e8 paths every strike must satisfy
   |delta| < 3 paired SE + 2 % of one MC standard error (the second term says: whatever systematic shift exists is below 1/50 of the
   statistical error of a 1e8-path price -- it would take 2.5e11 paths to see it).  Measured on B200 (r02): the paired difference IS
   resolved (6e3 paired SE: the SFU draws do shift prices systematically) and amounts to 1.3e-3 of one MC standard error at 1e8 paths,
   i.e. ~1e-7 relative to the price -- 6e13 paths would be needed to detect it statistically.
2. UNPAIRED: default mode vs ``gauss="fp64"`` (52-bit uniforms, one Philox call per step: different draws, no 6.7-sigma tail cut) at
   1e8 paths: every strike within 3 combined standard errors -- and the z-scores as a group look standard normal.
3. GPU default mode at 1e8 paths vs the C port of the reference arithmetic with fp64 libm draws (oracle/csrc/oracle_mc.c) on its own
   stream: within 3 combined standard errors.
All arithmetic under test runs through the C ABI; the oracle is only the checker.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NPY = 582
N_BIG = 100_000_000


def _chain():
    from stochvolmodels_b200 import LOGSV_BTC_PARAMS, get_btc_test_chain_data
    return get_btc_test_chain_data(), LOGSV_BTC_PARAMS


def _price(nb_path, seed, gauss):
    from stochvolmodels_b200 import LogSVPricer
    chain, params = _chain()
    p, e = LogSVPricer().model_mc_price_chain(chain, params, nb_path=nb_path, nb_steps=NPY, seed=seed, gauss=gauss)
    return np.concatenate(p), np.concatenate(e)


def test_paired_float_draws_vs_fp64_arithmetic_on_the_same_uniforms(cuda_lib):
    # paired SE of the difference, estimated from 24 independent seeds at 1e6 paths and scaled to 1e8 paths
    n_small, seeds = 1_000_000, range(1000, 1024)
    d_small = np.array([_price(n_small, s, "fp32")[0] - _price(n_small, s, "fp64_paired")[0] for s in seeds])
    paired_se_big = d_small.std(axis=0, ddof=1) * np.sqrt(n_small / N_BIG)
    p32, e32 = _price(N_BIG, 10, "fp32")
    p64, e64 = _price(N_BIG, 10, "fp64_paired")
    delta = p32 - p64
    bound = 3.0 * paired_se_big + 0.02 * e32
    worst = np.max(np.abs(delta) / bound)
    print(f"paired: max |delta| / MC SE = {np.max(np.abs(delta) / e32):.2e}, max |delta| / paired SE = {np.max(np.abs(delta) / paired_se_big):.2f}, "
          f"mean delta over seeds at 1e6 paths / MC SE(1e8) = {np.max(np.abs(d_small.mean(axis=0)) / e32):.2e}")
    assert worst < 1.0, (delta, paired_se_big, e32)
    # and the per-path perturbation is what an SFU-level error predicts: the two modes' prices differ by far less than one MC SE
    assert np.max(np.abs(delta) / e32) < 0.05


def test_float_draws_vs_fp64_draws_unpaired(cuda_lib):
    p32, e32 = _price(N_BIG, 10, "fp32")
    p64, e64 = _price(N_BIG, 11, "fp64")
    z = (p32 - p64) / np.sqrt(e32 ** 2 + e64 ** 2)
    print(f"unpaired f32 draws vs f64 draws at 1e8 paths: max |z| = {np.max(np.abs(z)):.2f}, mean z = {z.mean():.2f}")
    assert np.max(np.abs(z)) < 3.0, z
    # strikes of one maturity share paths (strongly correlated z), so only a loose group statement: no common shift beyond 2.5 SE
    assert abs(z.mean()) < 2.5


def test_gpu_default_vs_cport_fp64_libm_draws(cuda_lib):
    from oracle import cport
    chain, params = _chain()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n_cpu = 20_000_000 if cores >= 32 else 2_000_000          # ~10 s of host time either way
    p6 = (params.sigma0, params.theta, params.kappa1, params.kappa2, params.beta, params.volvol)
    pc, ec = cport.mc_chain("logsv", p6, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms, chain.optiontypes_ttms,
                            n_cpu, NPY, True, 1, 777, "f64", nthreads=cores)
    pc, ec = np.concatenate(pc), np.concatenate(ec)
    pg, eg = _price(N_BIG, 10, "fp32")
    z = (pg - pc) / np.sqrt(eg ** 2 + ec ** 2)
    print(f"GPU default (1e8 paths) vs C port fp64 draws ({n_cpu:.0e} paths): max |z| = {np.max(np.abs(z)):.2f}, mean z = {z.mean():.2f}")
    assert np.max(np.abs(z)) < 3.0, z
