#!/bin/bash
# Full ncu capture of one resident replay step (13 kernels) + the fp64 op counters.  Usage: bash tools/gpu_ncu_step.sh tag
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
GFPL_SWEEP=1 python tools/stage_ms.py 8192 3 > $O/${TAG}_stage_plain.log 2>&1; echo "plain rc=$?"; head -1 $O/${TAG}_stage_plain.log
GFPL_SWEEP=1 timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum \
  --clock-control none --import-source on -s 13 -c 13 -f -o $O/${TAG}_step python tools/stage_ms.py 8192 3 > $O/${TAG}_ncu_step.log 2>&1; echo "ncu rc=$?"
