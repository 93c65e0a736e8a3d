#!/usr/bin/env python
"""tools/ncu_summary.py -- parse ONE kernel of an ncu report into the JSON that bench.py's `roofline.traffic` / `compute_bound` read.

    python tools/ncu_summary.py gpurun_out/prof_slice_rXX.ncu-rep --paths 2e7 --steps 136 --out profiles/slice_kernel_metrics.json

`--paths` / `--steps` = the paths and time steps of the captured launch (the capture command is recorded in the JSON's `source`).
Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU); nothing is estimated -- every number is a counter of that
capture or a ratio of two of them.  Run it after every `ncu --set full` capture of a changed kernel and commit the JSON.
"""
import argparse
import csv
import io
import json
import subprocess

WANT = {
    "gpu__time_duration.sum": "duration_ms",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "dram__bytes_read.sum": "dram_read_mbyte",
    "dram__bytes_write.sum": "dram_write_mbyte",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_slots_busy_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_heavy_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.max": "cycles_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio": "stall_dispatch_per_issue",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle_per_issue",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected_per_issue",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default="mc_slice_kernel", help="substring of the kernel name (first match is used)")
    ap.add_argument("--paths", type=float, required=True)
    ap.add_argument("--steps", type=int, required=True)
    ap.add_argument("--command", default="", help="the capture command, for the record")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if a.kernel in r[hdr.index("Kernel Name")]]
    if not body:
        raise SystemExit(f"no kernel matching {a.kernel!r} in {a.report}")
    row = body[0]
    col = {h: i for i, h in enumerate(hdr)}
    m = {}
    for name, key in WANT.items():
        if name in col and row[col[name]] not in ("", "n/a"):
            v = float(row[col[name]].replace(",", ""))
            u = units[col[name]]
            if key.endswith("_mbyte") and u.lower().startswith("gbyte"):
                v *= 1e3
            if key == "duration_ms" and u == "us":
                v /= 1e3
            m[key] = v
    paths, steps = a.paths, a.steps
    warps_steps = paths / 32.0 * steps
    out = {"source": f"ncu --set full --clock-control none capture {a.report.split('/')[-1]} ({a.command or 'see profiles/'}); parsed by tools/ncu_summary.py",
           "kernel": row[col["Kernel Name"]], "captured_paths": paths, "captured_steps": steps,
           "dram_bytes_per_path_per_launch": (m["dram_read_mbyte"] + m["dram_write_mbyte"]) * 1e6 / paths,
           "issue_slots_busy": m["issue_slots_busy_pct"] / 100.0, "fp64_pipe": m["fp64_pipe_pct"] / 100.0, "xu_pipe": m["xu_pipe_pct"] / 100.0,
           "fma_heavy_pipe": m["fma_heavy_pipe_pct"] / 100.0, "alu_pipe": m["alu_pipe_pct"] / 100.0, "lsu_pipe": m["lsu_pipe_pct"] / 100.0,
           "tensor_pipe": m.get("tensor_pipe_pct", 0.0) / 100.0, "dram_throughput": m["dram_throughput_pct"] / 100.0,
           "warp_instructions_per_warp_step": m["warp_instructions"] / warps_steps,
           "clk_per_warp_step_per_smsp": m["cycles_elapsed"] * 148 * 4 / warps_steps,
           "kernel_path_steps_per_s_under_ncu": paths * steps / (m["duration_ms"] / 1e3),
           "registers_per_thread": m["registers_per_thread"], "grid": m["grid"], "block": m["block"],
           "stalls_per_issue": {k.replace("stall_", "").replace("_per_issue", ""): m[k] for k in m if k.startswith("stall_")}}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
