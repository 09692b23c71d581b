#!/bin/bash
# One GPU-box round trip: parity tests, the default bench line (configs[1] + the configs[2] 'also' leg + CPU baseline),
# ncu launch lists for configs A (B=16, 32x32) and B (B=64, 64x64), and one `ncu --set full` capture of the
# correlation kernel at each size.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [notests]'
mkdir -p gpurun_out
if [ "$1" != "notests" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
fi
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 4 --warmup 3 --graphs 0 --no-cpu-baseline --e2e-steps 2 --no-also > gpurun_out/ncu.log 2>&1; echo "ncuA rc=$?"
python bench.py --batch 64 --size 64 --steps 50 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/benchB.json 2> gpurun_out/benchB.err; echo "benchB rc=$?"; cut -c1-330 gpurun_out/benchB.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/launchesB.csv \
  python bench.py --batch 64 --size 64 --steps 3 --warmup 3 --graphs 0 --no-cpu-baseline --e2e-steps 2 --no-also > gpurun_out/ncuB.log 2>&1; echo "ncuB rc=$?"
for cfg in "A --batch 16 --size 32" "B --batch 64 --size 64"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:corr_tc -s 2 -c 2 -f -o gpurun_out/prof_corr_$1 \
    python bench.py $2 $3 $4 $5 --steps 1 --warmup 3 --graphs 0 --no-cpu-baseline --e2e-steps 1 --no-also > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu full $1 rc=$?"
done
