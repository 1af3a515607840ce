"""SASS instruction mix of the matching inner loops (evidence for the per-pair pipe counts used by bench.py's
roofline): python tools/sass_mix.py > profiles/r1_sass_mix.md   (needs cuobjdump, no GPU)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gf-pl-slam_b200", "lib", "libgfpl.so")
sass = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
funcs = re.split(r"\n\s*Function : ", sass)[1:]
print("# SASS mix of the Hamming matching kernels (cuobjdump -sass libgfpl.so, sm_100a)\n")
print("Unrolled body = 4 train rows; counts below are per (query, train row) pair = per 256-bit distance.\n")
for f in funcs:
    name = f.split("\n", 1)[0]
    if not re.search(r"knn2_sym_kernelILi6ELi2E|knn2_sym_kernelILi7ELi1E|f2f_hybrid_points_kernelE", name):
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M)
    c = collections.Counter(o.split(".")[0] if not o.startswith(("LDS", "UBLKCP", "SYNCS", "IMAD", "LOP3", "VIMNMX")) else o for o in ops)
    popc = c["POPC"]
    qpt = 2 if "ELi2E" in name else 1
    rows = popc / (5 if "Li6E" in name or "f2f" in name else 8) / qpt
    print("## `%s`\n" % name[:80])
    print("unrolled rows (main + remainder copies): %.1f, queries per thread: %d\n" % (rows, qpt))
    print("| op | total | per pair |\n|---|---|---|")
    pairs = rows * qpt
    for k in sorted(c, key=lambda k: -c[k]):
        if c[k] >= 4 and re.match(r"POPC|LOP3|IADD3|VIMNMX|IMAD|LEA|LDS|SHF|UBLKCP|SYNCS|ISETP|BRA|REDUX|CREDUX|ATOMS|SEL", k):
            print("| %s | %d | %.2f |" % (k, c[k], c[k] / pairs))
    print()
print("Pipes: POPC -> 16 lanes/clk/SM; LOP3/IADD3/VIMNMX/SHF/ISETP -> ALU pipe (64 lanes/clk/SM); IMAD -> FMA pipe; "
      "LDS.128 (broadcast) -> LSU.  `UBLKCP` = cp.async.bulk (1-D TMA), `SYNCS.*` = mbarrier.")

# ---- tensor-core kernels: proof of the tcgen05 / TMEM / bulk-copy path and the selection instruction mix ----------------
print("\n# Tensor-core kernels (tcgen05.mma kind::i8 -> UTCIMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, "
      "tcgen05.alloc/dealloc -> UTCATOMSWS / UTCDEALLOC... as cuobjdump names them)\n")
print("| kernel | UTCIMMA | LDTM | UTCBAR | SYNCS (mbarrier) | VIADDMNMX | VIMNMX3 | VIMNMX | PRMT | STS.128 | total |\n|---|---|---|---|---|---|---|---|---|---|---|")
for f in funcs:
    name = f.split("\n", 1)[0]
    if not re.search(r"knn2_tc_kernel|knn2_tc_small_kernel|knn2_tc_stream_kernel|f2f_hybrid_points_tc_kernel|int8_peak_kernel", name):
        continue
    ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M)
    c = collections.Counter(ops)
    def n(prefix):
        return sum(v for k, v in c.items() if k.startswith(prefix))
    short = re.sub(r"^_ZN4gfpl\d+", "", name)[:48]
    print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (short, n("UTCIMMA"), n("LDTM"), n("UTCBAR"), n("SYNCS"), n("VIADDMNMX"),
          n("VIMNMX3"), n("VIMNMX") - n("VIMNMX3"), n("PRMT"), n("STS.128"), len(ops)))
