"""BASELINE.json configs[4]: LogSV MC 1e8 paths x 1024 steps sharded over the GPUs of one box, payoff moments exchanged per maturity.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/config4_sharded.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from stochvolmodels_b200 import LogSvParams, LogSVPricer, OptionChain

params = LogSvParams(sigma0=1.0, theta=1.0, kappa1=5.0, kappa2=5.0, beta=0.2, volvol=2.0)
K = np.array([0.8, 0.9, 1.0, 1.1, 1.2])
chain = OptionChain(ttms=np.array([1.0]), forwards=np.array([1.0]), strikes_ttms=[K], optiontypes_ttms=[np.array(["P", "P", "C", "C", "C"])],
                    ids=np.array(["1y"]))
pricer = LogSVPricer()
N, NPY = 100_000_000, 1023                       # set_time_grid: int(1.0 * 1023) + 1 = 1024 steps
fourier = pricer.price_chain(chain, params)[0]
for exchange in ("p2p", "collective"):
    f = lambda seed: pricer.model_mc_price_chain(chain, params, nb_path=N, nb_steps=NPY, seed=seed, exchange=exchange)
    f(1)
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    prices, se = f(2)
    dist.barrier(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        z = (prices[0] - fourier) / se[0]
        print(f"config4 {world} GPUs, {exchange}: 1e8 paths x 1024 steps in {1e3 * dt.item():.1f} ms = {N * 1024 / dt.item():.3e} path-steps/s; "
              f"prices {np.round(prices[0], 6)} vs Fourier {np.round(fourier, 6)}; max |err|/F {np.max(np.abs(prices[0] - fourier)):.2e}, "
              f"(MC - Fourier)/SE {np.round(z, 2)}", flush=True)
from stochvolmodels_b200.multi_gpu import release_p2p
dist.barrier()
release_p2p()
dist.destroy_process_group()
