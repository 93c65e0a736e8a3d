#!/usr/bin/env python
"""bench.py -- LogSV Monte Carlo path-steps/sec on B200 (BASELINE.json `metric`), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--paths P] [--precision fp64|fp32] [--gauss fp32|fp64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of synthetic input: price the whole option chain by Monte Carlo once
(simulate every maturity slice with the fused kernel, re-centre, per-strike payoff moments, prices + std errors).

Workload (config.workload): LogSV MC, BTC-style option chain (4 maturities, 49 strikes, reference sample chain and
LOGSV_BTC_PARAMS), nb_steps_per_year=582 => 25+34+57+136 = 252 steps per path, 1e8 paths PER GPU (weak scaling: per-GPU work
is fixed, the two per-maturity exchange steps are fp64 all-reduces of <= 360 bytes).

`value`  : device-resident arm -- chain inputs already in HBM, CUDA-event timed on the launching stream, max over ranks.
`e2e`    : the public API call a user makes (LogSVPricer.model_mc_price_chain -> C ABI) with HOST buffers in and out, host<->device
           copies inside the timed region.
`roofline`: dominant kernel (mc_slice_kernel) -- algorithmic bytes = 64 B per path-step (the reference's streaming
           dataflow, 2 fp64 normals read + 3 fp64 state read + 3 written per path-step) over its CUDA-event duration, against
           the measured HBM copy bandwidth in MEASURED_PEAKS.json.  The fused kernel is physically issue-bound.
`cpu_baseline` / `--impl reference`: the oracle's C port of the reference algorithm (oracle/csrc/oracle_mc.c) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NB_STEPS_PER_YEAR = 582          # BTC chain maturity gaps -> 25 + 34 + 57 + 136 = 252 steps
ALGO_BYTES_PER_PATH_STEP = 64.0
SLICE_DRAM_BYTES_PER_PATH = 45.1  # measured: ncu --set full, profiles/r01_slice_kernel_ncu.txt (state read + write per launch)
SEED = 10


def workload():
    from stochvolmodels_b200 import LOGSV_BTC_PARAMS, get_btc_test_chain_data
    from stochvolmodels_b200.utils.funcs import set_time_grid
    chain = get_btc_test_chain_data()
    steps, t0 = [], 0.0
    for ttm in chain.ttms:
        steps.append(set_time_grid(ttm - t0, NB_STEPS_PER_YEAR)[0])
        t0 = ttm
    return chain, LOGSV_BTC_PARAMS, steps


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        clocks = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        maxc = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        power = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": max(maxc) if maxc else None,
                "power_w_max": max(power) if power else None, "samples": len(clocks), "reasons": reasons}


def cpu_port_rate(chain, params, budget_s: float, nthreads: int = 0):
    """time the C port (all host threads) on a bounded sample of the SAME workload; returns (path_steps/s, paths, seconds, threads)."""
    from oracle import cport
    p6 = (params.sigma0, params.theta, params.kappa1, params.kappa2, params.beta, params.volvol)
    # all host threads this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the N>1 reference arm into a single-core run (measured: 1.9e7 instead of 1.5e8 path-steps/s)
    threads = (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)) if nthreads <= 0 else nthreads
    run = lambda n: cport.mc_chain("logsv", p6, chain.ttms, chain.forwards, chain.discfactors, None, chain.strikes_ttms,
                                   chain.optiontypes_ttms, n, NB_STEPS_PER_YEAR, True, 1, SEED, "f64", nthreads=threads)
    n0 = 20_000 * threads
    run(n0 // 4)                                   # warm (page faults, thread pool)
    t = time.perf_counter(); run(n0); dt0 = time.perf_counter() - t
    n = int(min(max(n0 * budget_s / max(dt0, 1e-3), n0), 4e8))
    t = time.perf_counter(); run(n); dt = time.perf_counter() - t
    return n * 252 / dt, n, dt, threads


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (C port of the oracle; the Python/Numba reference cannot travel to the
    GPU box) on all host cores, same workload / metric / unit; rank 0 only."""
    if rank != 0:
        return
    chain, params, steps = workload()
    total = max(1, args.steps + args.warmup)
    budget = min(20.0, 150.0 / total)
    budget = float(os.environ.get("B200SV_BENCH_CPU_BUDGET_S", budget))     # tests shorten the bounded sample
    rates = []
    n = secs = threads = None
    for i in range(total):
        r, n, secs, threads = cpu_port_rate(chain, params, budget)
        if i >= args.warmup:
            rates.append(r)
    value = statistics.median(rates)
    line = {"impl": "reference", "metric": "LogSV MC path-steps/sec", "value": value, "unit": "path-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "LogSV MC, BTC chain 4 maturities x 49 strikes, 252 steps/path (nb_steps_per_year=582), "
                                   f"bounded sample of {n} paths per step", "paths_per_step": n},
            "cpu_baseline": {"value": value, "unit": "path-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{n} paths x 252 steps, oracle/csrc/oracle_mc.c (OpenMP, fp64, Philox stream), {secs:.1f} s"},
            "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--paths", type=float, default=1e8, help="paths per GPU")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--gauss", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "collective"],
                    help="N>1: per-maturity moment exchange through NVLink peer memory inside the kernels (p2p) or NCCL all-reduce (collective)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from stochvolmodels_b200 import LogSVPricer, engine
    from stochvolmodels_b200 import _capi as C
    from stochvolmodels_b200.multi_gpu import CudaMcEngine
    from stochvolmodels_b200.pricers.logsv_pricer import _params_c
    from stochvolmodels_b200.utils.funcs import set_time_grid

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: stochvolmodels_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = C.load_library()
    chain, params, steps = workload()
    n_local = int(args.paths)
    total_paths = n_local * world
    flags = engine.mc_flags(args.precision, args.gauss)
    offsets, strikes, types = C.flatten_chain(chain.strikes_ttms, chain.optiontypes_ttms)
    sizes = np.diff(offsets)
    eng = CudaMcEngine("logsv", _params_c(params), n_local, rank * n_local, flags, int(sizes.max()))
    use_p2p = world > 1 and args.exchange == "p2p"     # the two per-maturity messages ride inside the kernels over NVLink peer memory
    if use_p2p:
        from stochvolmodels_b200.multi_gpu import P2pUnavailable
        try:
            eng.enable_p2p()
        except P2pUnavailable as e:          # raised on every rank together; the NCCL exchange is the same arithmetic
            if rank == 0:
                print(f"bench.py: peer-memory exchange unavailable ({e}); using NCCL all-reduces", file=sys.stderr)
            use_p2p = False
            args.exchange = "collective"
    strikes_dev = eng.to_device(strikes, torch.float64)
    types_dev = eng.to_device(types, torch.int8)
    out_dev = torch.zeros((2, strikes.shape[0]), dtype=torch.float64, device=eng.device)
    grid = []
    t0 = 0.0
    for ttm in chain.ttms:
        s, dt, _ = set_time_grid(ttm - t0, NB_STEPS_PER_YEAR)
        grid.append((s, dt))
        t0 = ttm
    path_steps_per_gpu = n_local * sum(s for s, _ in grid)
    slice_events = []

    def one_chain(seed, record):
        for m, (S, dt) in enumerate(grid):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            mom = eng.simulate_slice(m, m == 0, S, dt, 1.0, True, float(chain.forwards[m]), seed)
            if record:
                e1.record()
                slice_events.append((e0, e1, S))
            if world > 1 and not use_p2p:
                dist.all_reduce(mom)
            J, jo = int(sizes[m]), int(offsets[m])
            sums = eng.payoff_sums(float(chain.ttms[m]), float(chain.forwards[m]), strikes_dev[jo: jo + J], types_dev[jo: jo + J], J, C.LOG_RETURN, 1)
            if world > 1 and not use_p2p:
                dist.all_reduce(sums)
            p, s = eng.finalize(sums, J, float(chain.discfactors[m]), total_paths)
            out_dev[0, jo: jo + J].copy_(p)
            out_dev[1, jo: jo + J].copy_(s)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    for w in range(args.warmup):
        one_chain(SEED + w, False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.b200sv_reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        one_chain(SEED + 100 + k, True)
    ev1.record()
    barrier()
    launches = int(lib.b200sv_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed_ms.item()) / 1e3
    value = path_steps_per_gpu * world * args.steps / elapsed_s
    slice_ms = sum(a.elapsed_time(b) for a, b, _ in slice_events)
    slice_path_steps = sum(n_local * S for _, _, S in slice_events)
    prices_dev = out_dev.cpu().numpy()

    # ---------------- end-to-end arm: public API, host buffers ----------------
    pricer = LogSVPricer()
    api = lambda seed: pricer.model_mc_price_chain(chain, params, nb_path=total_paths, nb_steps=NB_STEPS_PER_YEAR, seed=seed,
                                                   precision=args.precision, gauss=args.gauss, exchange=args.exchange)
    del eng
    torch.cuda.empty_cache()
    for w in range(min(args.warmup, 2)):
        api(SEED + w)
    barrier()
    t = time.perf_counter()
    for k in range(args.steps):
        prices_api, _ = api(SEED + 100 + k)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = path_steps_per_gpu * world * args.steps / float(e2e_s.item())
    h2d = strikes.nbytes + types.nbytes + 6 * 8 + 4 * 8 * 4 + offsets.nbytes      # strikes, types, params, ttms/forwards/dfs/etas, offsets
    d2h = 2 * strikes.shape[0] * 8

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        # per launch: algorithmic bytes = 64 B x (n_local x S_m); achieved = sum of bytes / sum of CUDA-event durations of the slice launches
        achieved = ALGO_BYTES_PER_PATH_STEP * slice_path_steps / (slice_ms / 1e3) / 1e9
        same = bool(np.allclose(np.concatenate(prices_api), prices_dev[0], rtol=1e-12)) if world == 1 else None
        line = {"metric": "LogSV MC path-steps/sec", "value": value, "unit": "path-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * elapsed_s / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
                "config": {"workload": "LogSV MC, BTC chain 4 maturities x 49 strikes, 252 steps/path (nb_steps_per_year=582, "
                                       "slices 25+34+57+136), LOGSV_BTC_PARAMS", "paths_per_gpu": n_local, "paths_total": total_paths,
                           "state": args.precision, "gaussians": f"Philox4x32-10 + Box-Muller {args.gauss}",
                           "parallelism": f"paths sharded over {world} GPU(s); 2 fp64 exchanges (16 B, 24*J B) per maturity, "
                                          + ("none needed at N=1" if world == 1 else ("fused into the reduction/payoff kernels over NVLink peer memory"
                                                                                      if use_p2p else "NCCL all-reduce")),
                           "l2": "per-GPU state 3 x paths x 8 B >> 126 MB L2 (inputs larger than L2; no flush needed)"},
                "e2e": {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "api": "LogSVPricer.model_mc_price_chain (host numpy in/out through the C ABI)", "same_prices_as_device_arm": same},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": SLICE_DRAM_BYTES_PER_PATH * n_local * len(slice_events) / max(args.steps, 1) / len(grid),
                             "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum = 45.1 B per path per launch "
                                               "(profiles/r01_slice_kernel_ncu.txt), scaled to this launch size",
                             "kernel": "mc_slice_kernel<LogsvPath>", "peak_source": peak_src,
                             "algorithmic_bytes_per_path_step": ALGO_BYTES_PER_PATH_STEP,
                             "kernel_share_of_step": slice_ms / (1e3 * elapsed_s),
                             "note": "effective bandwidth of the reference's streaming dataflow; the fused kernel keeps state in registers, so "
                                     "frac > 1 means the HBM roofline is not what binds it: see compute_bound",
                             # what does bind it (one ncu --set full capture of this kernel per change, not measured in this run)
                             "compute_bound": {"source": "profiles/r01_slice_kernel_ncu.txt (capture r1n) + profiles/r01_pipe_overlap.txt",
                                               "issue_slots_busy": 0.586, "fp64_pipe": 0.326, "xu_pipe": 0.408, "fma_heavy_pipe": 0.419,
                                               "alu_pipe": 0.378, "dram_throughput": 0.013,
                                               "warp_instructions_per_path_step": 70.2 / 32,
                                               "clk_per_warp_step_per_smsp": {"measured": 120}}},
                "clocks": clocks}
        # second half of BASELINE.json's metric: price error of the timed MC chain against the Fourier reference (our GPU Fourier route,
        # itself within 1e-10 of the reference CPU path, tests/test_gpu_mgf.py); every one of the 49 strikes
        try:
            fourier = np.concatenate(pricer.price_chain(chain, params))
            fwd = np.concatenate([np.full(int(sizes[m]), chain.forwards[m]) for m in range(len(sizes))])
            err = np.abs(prices_dev[0] - fourier)
            line["price_err_vs_fourier"] = {"max_abs_over_forward": float(np.max(err / fwd)), "max_in_se": float(np.max(err / prices_dev[1])),
                                            "note": "explicit-Euler bias at 25..136 steps per slice dominates at 1e8 paths (DESIGN.md §6); "
                                                    "north_star abs tolerance 1e-3"}
        except Exception as exc:      # never lose the throughput line over the accuracy annotation
            line["price_err_vs_fourier"] = {"error": str(exc)}
        if not args.no_cpu_baseline and world == 1:
            r, n, secs, threads = cpu_port_rate(chain, params, 12.0)
            line["cpu_baseline"] = {"value": r, "unit": "path-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{n} paths x 252 steps of the same chain, oracle/csrc/oracle_mc.c (OpenMP, fp64), {secs:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
