#!/bin/bash
# Kernel experiments: parity tests on the default build, then per-kernel stage times of every side build under
# gf-pl-slam_b200/lib/variants/*.  Usage (through gpurun): bash tools/gpu_variants.sh tag "pytest -k expression"
TAG=${1:-var}
KEXPR=${2:-"pose or track or replay"}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$KEXPR" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
echo "== default" | tee $O/${TAG}_stage.log
GFPL_SWEEP=1 python tools/stage_ms.py 8192 5 2>&1 | tee -a $O/${TAG}_stage.log
for d in gf-pl-slam_b200/lib/variants/*/; do
  n=$(basename $d)
  echo "== $n" | tee -a $O/${TAG}_stage.log
  GFPL_LIB=$d/libgfpl.so GFPL_SWEEP=1 python tools/stage_ms.py 8192 5 2>&1 | tee -a $O/${TAG}_stage.log
done
