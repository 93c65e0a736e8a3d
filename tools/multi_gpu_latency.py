"""torchrun helper: latency of the sharded MC chain call (public API, host buffers) with the two per-maturity exchanges done
(a) inside the kernels over NVLink peer memory (exchange="p2p") and (b) by NCCL all-reduces (exchange="collective").
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from stochvolmodels_b200 import LOGSV_BTC_PARAMS, LogSVPricer, get_btc_test_chain_data

chain = get_btc_test_chain_data()
pricer = LogSVPricer()
if rank == 0:
    print(f"# BTC chain (4 maturities, 49 strikes), nb_steps=360, {world} GPUs; median of 20 calls, max over ranks; ms per call", flush=True)
    print("total paths | p2p (fused) | NCCL all-reduce | single GPU (distributed=False) | default (replicated below 5e5 paths, else p2p)", flush=True)
for n in (10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
    row = []
    for kw in (dict(exchange="p2p"), dict(exchange="collective"), dict(distributed=False), dict()):
        f = lambda: pricer.model_mc_price_chain(chain, LOGSV_BTC_PARAMS, nb_path=n, nb_steps=360, seed=1, **kw)
        for _ in range(3):
            f()
        ts = []
        for _ in range(20 if n <= 10_000_000 else 5):
            dist.barrier(); torch.cuda.synchronize()
            t = time.perf_counter(); f(); ts.append(time.perf_counter() - t)
        med = torch.tensor([float(np.median(ts))], device="cuda")
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
        row.append(1e3 * float(med.item()))
    if rank == 0:
        print(f"{n:>11d} | {row[0]:10.3f} | {row[1]:10.3f} | {row[2]:10.3f} | {row[3]:10.3f}", flush=True)
from stochvolmodels_b200.multi_gpu import release_p2p
dist.barrier()
release_p2p()
dist.destroy_process_group()
