set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_t0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t0.log
python bench.py > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
for s in 1234 1235 1236 1237 1238 1239 1240 1241; do
  python bench.py --batch 64 --size 64 --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 4 --no-also --seed $s > gpurun_out/r2_seed_$s.json 2> gpurun_out/r2_seed_$s.err
done
tail -3 gpurun_out/r2_t0.log
