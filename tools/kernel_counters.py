"""profiles/<tag>_kernel_counters.json + profiles/<tag>_ncu_full_summary.md from one `ncu --set full` capture of a replay step:
   python tools/kernel_counters.py gpurun_out/x.ncu-rep r2 [pairs n_pt n_ln]
per kernel (first launch of each name): duration under ncu, DRAM bytes, executed fp64 flops (dadd + dmul + 2 dfma), pipe and
issue utilisation, occupancy, stall picture.  bench.py reads the JSON for `roofline.traffic` and the line-cut flop count."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
pairs, n_pt, n_ln = (int(x) for x in sys.argv[3:6]) if len(sys.argv) > 5 else (8192, 300, 100)
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def short(n):
    return n.replace("gfpl::", "").replace("void ", "").split("(")[0]


def val(r, m, default=None):
    if m not in idx or r[idx[m]] in ("", "n/a"):
        return default
    v = float(r[idx[m]].replace(",", ""))
    u = units[idx[m]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "msecond": 1e3, "us": 1.0, "usecond": 1.0, "ns": 1e-3,
                "nsecond": 1e-3}.get(u, 1.0)


show = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
per, seen = {}, set()
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of one resident replay step\n\n"
            f"Workload: {pairs} frame pairs, {n_pt} points + {n_ln} lines per image, line cutting on (`tools/stage_ms.py`, second step). Numbers taken "
            "under the profiler are for diagnosis only (cold caches, serialised kernels); the timings of BENCH / DESIGN come from CUDA events without ncu.\n\n")
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        if k in seen:
            continue
        seen.add(k)
        fl = sum((val(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum", 0.0) or 0.0) * w for op, w in (("dadd", 1), ("dmul", 1), ("dfma", 2)))
        per[k] = {"duration_us_under_ncu": val(r, "gpu__time_duration.sum"), "dram_bytes": (val(r, "dram__bytes_read.sum", 0) or 0) + (val(r, "dram__bytes_write.sum", 0) or 0),
                  "fp64_flops": fl or None, "warp_instructions": val(r, "smsp__inst_executed.sum"),
                  "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "fp64_pipe_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                  "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                  "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
        f.write(f"## `{k}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for m in show:
            if m in idx:
                f.write(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |\n")
        if fl:
            f.write(f"| executed fp64 flops (dadd + dmul + 2 dfma, predicated-on threads) | {fl:.4g} | flop |\n")
        f.write("\n")
json.dump({"workload": "tools/stage_ms.py: one resident replay step", "pairs": pairs, "n_pt": n_pt, "n_ln": n_ln, "per_launch": per},
          open(os.path.join(ROOT, "profiles", f"{tag}_kernel_counters.json"), "w"), indent=1)
for k, v in per.items():
    print("%-40s %8.1f us  dram %7.1f MB  fp64 %s  issue %s%%  warps %s%%" % (k, v["duration_us_under_ncu"], v["dram_bytes"] / 1e6,
          ("%.3g" % v["fp64_flops"]) if v["fp64_flops"] else "-", v["issue_active_pct"], v["warps_active_pct"]))
