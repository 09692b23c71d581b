#!/bin/bash
# One GPU-box round trip: parity tests, bench + ncu launch list for configs A (B=16, 32x32) and B (B=64, 64x64).
# usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [notests]'
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
fi
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "benchA rc=$?"; cut -c1-330 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --graphs 0 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu.log 2>&1; echo "ncuA rc=$?"
python bench.py --batch 64 --size 64 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/benchB.json 2> gpurun_out/benchB.err; echo "benchB rc=$?"; cut -c1-330 gpurun_out/benchB.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/launchesB.csv \
  python bench.py --batch 64 --size 64 --steps 3 --warmup 3 --graphs 0 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncuB.log 2>&1; echo "ncuB rc=$?"
