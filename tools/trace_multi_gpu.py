#!/usr/bin/env python
"""Timeline of the sharded MC API call on N GPUs without nsys (not in the image): per call and per rank, host wall-clock stamps
(enter, all kernels enqueued, stream synchronised -- time.time() is one clock for all ranks of a box) and the CUDA-event duration of
the chain on that rank's stream.  Attributes the e2e-vs-device gap of bench.py to (a) host work before the first
launch, (b) rank skew at entry (a late rank makes every peer's gather spin), (c) the device work itself.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/trace_multi_gpu.py [p2p|collective] [paths_per_gpu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200SV_TRACE"] = "1"


def main():
    import torch
    import torch.distributed as dist
    from stochvolmodels_b200 import LOGSV_BTC_PARAMS, LogSVPricer, get_btc_test_chain_data, multi_gpu
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    exchange = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    n_local = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
    chain, p = get_btc_test_chain_data(), LOGSV_BTC_PARAMS
    pricer = LogSVPricer()
    call = lambda s: pricer.model_mc_price_chain(chain, p, nb_path=n_local * world, nb_steps=582, seed=s, exchange=exchange)
    for w in range(3):
        call(w)
    multi_gpu.TRACE_LOG.clear()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    K = int(os.environ.get("TRACE_CALLS", "10"))
    import gc
    if os.environ.get("TRACE_NO_GC"):
        gc.disable()
    def cpu_stat():
        try:
            return dict(l.split() for l in open("/sys/fs/cgroup/cpu.stat"))
        except Exception:
            return {}
    st0, gc0 = cpu_stat(), [g["collections"] for g in gc.get_stats()]
    for k in range(K):
        call(100 + k)
    dist.barrier(); torch.cuda.synchronize()
    total = time.time() - t0
    st1, gc1 = cpu_stat(), [g["collections"] for g in gc.get_stats()]
    if rank == 0:
        print("cgroup cpu.stat deltas:", {k: int(st1[k]) - int(st0[k]) for k in st1 if k in ("nr_periods", "nr_throttled", "throttled_usec", "usage_usec")},
              "gc collections (gen0,1,2):", [b - a for a, b in zip(gc0, gc1)])
    logs = [None] * world
    dist.all_gather_object(logs, multi_gpu.TRACE_LOG)
    if rank == 0:
        print(f"exchange={exchange} world={world} paths/gpu={n_local:.0e}: {1e3 * total / K:.2f} ms per call (wall, incl. barriers)")
        print("call | per rank: enter offset vs earliest rank (ms), host launch phase (ms), gpu chain (ms), enter->synced (ms)")
        slow = [k for k in range(K) if max(1e3 * (logs[r][k]["t_synced"] - logs[r][k]["t_enter"]) for r in range(world)) > 1.08 * np.median([l["gpu_ms"] for l in logs[0]])]
        print("slow calls:", slow, "at t(s) =", [round(logs[0][k]["t_enter"] - t0, 2) for k in slow])
        for k in (range(K) if K <= 12 else slow[:12]):
            rows = [logs[r][k] for r in range(world)]
            e0 = min(r["t_enter"] for r in rows)
            print(f"{k:3d}  | " + " | ".join(f"r{r['rank']}: +{1e3 * (r['t_enter'] - e0):6.2f} {1e3 * (r['t_launched'] - r['t_enter']):6.2f} {r['gpu_ms']:7.2f} {1e3 * (r['t_synced'] - r['t_enter']):7.2f}" for r in rows))
        gaps = [1e3 * (logs[r][k + 1]["t_enter"] - logs[r][k]["t_synced"]) for r in range(world) for k in range(K - 1)]
        print(f"host gap between calls (synced -> next enter): median {np.median(gaps):.3f} ms, max {np.max(gaps):.3f} ms")
        print(f"gpu chain ms: median {np.median([l['gpu_ms'] for lr in logs for l in lr]):.2f}, max {np.max([l['gpu_ms'] for lr in logs for l in lr]):.2f}")
    dist.barrier()
    multi_gpu.release_p2p()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
