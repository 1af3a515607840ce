"""Summaries for profiles/ from the artefacts a gpurun call brought back:
   python tools/profile_summary.py gpurun_out/launches_r1.csv gpurun_out/prof_r1_final.ncu-rep r1
"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]

def short(n):
    n = n.replace("gfpl::", "").replace("void ", "")
    return n.split("(")[0]

# ---- launch list -------------------------------------------------------------------------------------------
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    k = short(r[ki]); t = float(r[vi].replace(",", "")) * (1e-3 if r[ui] in ("ns", "nsecond") else 1.0)  # -> us
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md"), "w") as f:
    f.write(f"# {tag}: kernel launch list of `python bench.py --skip-cpu-baseline` (ncu --metrics gpu__time_duration.sum --clock-control none -c 400)\n\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes. Raw list: "
            f"`profiles/{tag}_launches.csv`.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f} % |\n")
    f.write(f"\nFirst 400 launches of the run: setup (stereo stage of the previous frames, 8 chunks), the pipe "
            f"micro-benchmarks (`popc_peak_kernel`, `lop3_peak_kernel`, `fp64_peak_kernel`), warm-up, timed steps "
            f"(13 kernels per step), the streaming e2e steps (`expand_compact_kernel` + the same 13) and the LBA side note (`lba_*`).\n")
import shutil
shutil.copy(launches, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))

# ---- full capture --------------------------------------------------------------------------------------------
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
traffic = {}
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the hot kernels\n\n"
            "Command: `python bench.py --steps 2 --warmup 1 --skip-cpu-baseline` (default workload: 8192 EuRoC-shaped pairs, "
            "300 points + 100 lines per image, line cutting on), `-k regex:knn2_tc_kernel|knn2_tc_small_kernel|knn2_sym_kernel|f2f_hybrid_points_tc_kernel|"
            "linecut_kernel|pose_optimize_kernel|stereo_accept_kernel -s 48 -c 8`. Numbers taken under the profiler are for diagnosis only; "
            "the timings in BENCH/README come from CUDA events without ncu.\n\n")
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        k = short(name)
        if (k, r[idx["launch__grid_size"]]) in seen:
            continue
        seen.add((k, r[idx["launch__grid_size"]]))
        f.write(f"## `{k}`\n\n| metric | value | unit |\n|---|---|---|\n")
        for m in keys:
            if m in idx:
                f.write(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |\n")
        f.write("\n")
        def num(m):
            v = float(r[idx[m]].replace(",", "")); u = units[idx[m]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        traffic[k + "@" + r[idx["launch__grid_size"]]] = {"dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
                                                         "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else (1e3 if units[idx["gpu__time_duration.sum"]].startswith("ms") else 1.0))}
json.dump({"workload": "bench.py default (8192 pairs, 300 pts + 100 lines)", "per_launch": traffic},
          open(os.path.join(ROOT, "profiles", f"{tag}_kernel_traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md")).read())
print(json.dumps(traffic, indent=1)[:1500])
