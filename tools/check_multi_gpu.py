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
    p_d, e_d = pricer.model_mc_price_chain(chain, params, nb_path=N, seed=123, **kw)                       # sharded over `world` GPUs
    p_s, e_s = pricer.model_mc_price_chain(chain, params, nb_path=N, seed=123, distributed=False, **kw)    # this rank alone
    rel_p = max(np.max(np.abs(a / b - 1)) for a, b in zip(p_d, p_s))
    rel_e = max(np.max(np.abs(a / b - 1)) for a, b in zip(e_d, e_s))
    if rank == 0:
        print(f"{name}: world={world} N={N} max rel diff prices {rel_p:.2e} stderr {rel_e:.2e}", flush=True)
    ok &= rel_p < 1e-12 and rel_e < 1e-10
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
