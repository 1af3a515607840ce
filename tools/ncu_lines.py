"""Per-source-line summary of one kernel of an ncu report (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]
prints file:line, warp instructions executed, stall samples, for the hottest lines."""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = None
hdr = None
lines = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] == "" : continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def g(name):
        v = r[hdr[name]] if name in hdr and hdr[name] < len(r) else "0"
        try: return float(v)
        except ValueError: return 0.0
    key = (cur_file, ln)
    a = lines.setdefault(key, [r[1].strip()[:90], 0.0, 0.0])
    a[1] += g("Instructions Executed"); a[2] += g("Warp Stall Sampling (All Samples)")
ti = sum(a[1] for a in lines.values()); ts = sum(a[2] for a in lines.values())
print("total warp instructions %.0f, stall samples %.0f" % (ti, ts))
by_file = collections.Counter()
for (f, ln), a in lines.items(): by_file[f] += a[1]
print("by file:", {f: "%.1f%%" % (100 * v / ti) for f, v in by_file.most_common()})
for (f, ln), a in sorted(lines.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-22s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100 * a[1] / ti, 100 * a[2] / ts, a[0]))
# optional: regions "name:file:lo-hi,..." in env NCU_REGIONS -> share of instructions / samples per region
import os
reg = os.environ.get("NCU_REGIONS")
if reg:
    print("regions:")
    for item in reg.split(","):
        name, f, rng = item.split(":"); lo, hi = map(int, rng.split("-"))
        i = sum(a[1] for (ff, ln), a in lines.items() if ff == f and lo <= ln <= hi)
        s = sum(a[2] for (ff, ln), a in lines.items() if ff == f and lo <= ln <= hi)
        print("  %-28s inst %5.1f%%  samples %5.1f%%" % (name, 100 * i / ti, 100 * s / ts))
