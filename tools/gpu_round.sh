#!/bin/bash
# One GPU-box call of a development round: GPU tests, bench line, ncu launch list, sanitizer passes of the smoke step.
# Usage (through gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $O/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --c4-pairs 0 --lba-windows 0 > $O/${TAG}_ncu_launch.log 2>&1; echo "ncu launch rc=$?"
