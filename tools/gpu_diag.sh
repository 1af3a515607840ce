#!/bin/bash
# Diagnosis call: full ncu capture of the fp64 kernels, the min-eigenvalue diagnostic, compute-sanitizer passes of the smoke step.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:pose_optimize_kernel|linecut_kernel' -s 2 -c 2 \
    -f -o $O/${TAG}_fp64 python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --c4-pairs 0 --lba-windows 0 > $O/${TAG}_ncu_fp64.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/mineig_diag.py > $O/${TAG}_mineig.log 2>&1; echo "mineig rc=$?"; tail -8 $O/${TAG}_mineig.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python __graft_entry__.py smoke > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/${TAG}_racecheck.log
