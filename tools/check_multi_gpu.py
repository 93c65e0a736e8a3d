"""torchrun helper: sharded (NCCL) chain MC == single-GPU chain MC for the same TOTAL path count and seed.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from stochvolmodels_b200 import LOGSV_BTC_PARAMS, HestonParams, HestonPricer, LogSVPricer, get_btc_test_chain_data

chain = get_btc_test_chain_data()
N = 4_000_001
ok = True
for name, pricer, params, kw in (("logsv", LogSVPricer(), LOGSV_BTC_PARAMS, dict(nb_steps=252)), ("heston", HestonPricer(), HestonParams(v0=0.8, theta=1.0, kappa=2.0, rho=0.0, volvol=2.0), {})):
    p_d, e_d = pricer.model_mc_price_chain(chain, params, nb_path=N, seed=123, **kw)                       # sharded, peer-memory exchange
    p_c, e_c = pricer.model_mc_price_chain(chain, params, nb_path=N, seed=123, exchange="collective", **kw)  # sharded, NCCL all-reduce
    same = all(np.array_equal(a, b) for a, b in zip(p_d, p_c)) and all(np.array_equal(a, b) for a, b in zip(e_d, e_c))
    if rank == 0:
        print(f"{name}: p2p exchange == NCCL exchange bitwise: {same}" + ("" if world == 2 else "  (only expected at 2 ranks: NCCL sums in its own order)"), flush=True)
    gathered = [None] * world
    dist.all_gather_object(gathered, np.concatenate(p_d).tobytes())
    ok &= all(g == gathered[0] for g in gathered)            # every rank holds the same bits
    p_s, e_s = pricer.model_mc_price_chain(chain, params, nb_path=N, seed=123, distributed=False, **kw)    # this rank alone
    rel_p = max(np.max(np.abs(a / b - 1)) for a, b in zip(p_d, p_s))
    rel_e = max(np.max(np.abs(a / b - 1)) for a, b in zip(e_d, e_s))
    if rank == 0:
        print(f"{name}: world={world} N={N} max rel diff prices {rel_p:.2e} stderr {rel_e:.2e}", flush=True)
    ok &= rel_p < 1e-12 and rel_e < 1e-10

# the two neighbouring Monte Carlo routes shard the same way (Philox draws; global path ids): Hawkes jump-diffusion and rough LogSV
from stochvolmodels_b200 import HawkesJDParams, HawkesJDPricer, LogSvParams
rough_params = LogSvParams(sigma0=0.8, theta=1.0, kappa1=2.2, kappa2=2.2, beta=0.2, volvol=1.6, H=0.3, weights=np.array([0.7, 0.5, 0.3]),
                           nodes=np.array([0.05, 1.5, 20.0]))
for name, pricer, params, kw, n in (("hawkes", HawkesJDPricer(), HawkesJDParams(), {}, 400_001),
                                    ("rough", LogSVPricer(), rough_params, dict(use_rough_mc=True, gauss="fp32", nb_steps=360), 400_001)):
    p_d, e_d = pricer.model_mc_price_chain(chain, params, nb_path=n, seed=321, exchange="p2p", **kw)
    p_c, e_c = pricer.model_mc_price_chain(chain, params, nb_path=n, seed=321, exchange="collective", **kw)
    p_s, e_s = pricer.model_mc_price_chain(chain, params, nb_path=n, seed=321, distributed=False, **kw)
    rel_p = max(np.max(np.abs(a / b - 1)) for a, b in zip(p_d, p_s))
    rel_c = max(np.max(np.abs(a / b - 1)) for a, b in zip(p_c, p_s))
    rel_e = max(np.max(np.abs(a / b - 1)) for a, b in zip(e_d, e_s))
    if rank == 0:
        print(f"{name}: world={world} N={n} max rel diff vs single GPU: prices p2p {rel_p:.2e} NCCL {rel_c:.2e} stderr {rel_e:.2e}", flush=True)
    ok &= rel_p < 1e-12 and rel_c < 1e-12 and rel_e < 1e-10

# small calls are priced REPLICATED (every rank runs all paths through the single-launch kernel): bit-identical to this rank alone, the same on
# every rank -- also without a seed (rank 0's fresh seed is broadcast)
lp = LogSVPricer()
r_seed = lp.model_mc_price_chain(chain, LOGSV_BTC_PARAMS, nb_path=20_000, nb_steps=252, seed=5)
r_single = lp.model_mc_price_chain(chain, LOGSV_BTC_PARAMS, nb_path=20_000, nb_steps=252, seed=5, distributed=False)
same_single = all(np.array_equal(a, b) for a, b in zip(r_seed[0], r_single[0]))
r_free = lp.model_mc_price_chain(chain, LOGSV_BTC_PARAMS, nb_path=20_000, nb_steps=252)
gathered = [None] * world
dist.all_gather_object(gathered, np.concatenate(r_free[0]).tobytes())
same_ranks = all(g_ == gathered[0] for g_ in gathered)
if rank == 0:
    print(f"replicated small call: == single GPU bitwise {same_single}; seedless call identical on all ranks {same_ranks}", flush=True)
ok &= same_single and same_ranks

# ragged corner cases through the peer-memory exchange: a maturity without strikes, fewer paths than ranks (ranks without paths),
# IC/IP payoffs (general payoff kernel), and many repeated calls (epoch wrap of the double-buffered mailbox)
from stochvolmodels_b200.pricers.logsv_pricer import logsv_mc_chain_pricer
P = LOGSV_BTC_PARAMS
rag = dict(ttms=np.array([0.1, 0.2, 0.4]), forwards=np.array([1.0, 1.0, 1.0]), discfactors=np.array([1.0, 0.99, 0.98]),
           strikes_ttms=[np.array([0.9, 1.0, 1.1]), np.zeros(0), np.array([0.8, 1.2])],
           optiontypes_ttms=[np.array(["C", "IP", "P"]), np.array([], dtype="<U2"), np.array(["IC", "C"])],
           v0=P.sigma0, theta=P.theta, kappa1=P.kappa1, kappa2=P.kappa2, beta=P.beta, volvol=P.volvol, vol_backbone_etas=np.ones(3),
           nb_steps_per_year=100)
for n in (world - 1, 3 * world + 1, 200_003):
    if n < 1:
        continue
    for rep in range(3):
        a, ea = logsv_mc_chain_pricer(nb_path=n, seed=7 + rep, exchange="p2p", **rag)       # named exchange: sharded even for a few paths
        b, eb = logsv_mc_chain_pricer(nb_path=n, seed=7 + rep, distributed=False, **rag)
        d = max(np.max(np.abs(x - y)) for x, y in zip(a, b) if x.size)
        good = d < 1e-12 and all(np.all(np.isfinite(x)) for x in a)
        ok &= good
        if rank == 0 and (rep == 0 or not good):
            print(f"ragged chain N={n}: max abs diff vs single GPU {d:.2e} finite={good}", flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
ok = bool(flag.item() > 0.5)
from stochvolmodels_b200.multi_gpu import release_p2p
dist.barrier()
release_p2p()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

