python -m pytest tests/test_gpu_patches.py tests/test_gpu_multi.py -x -q > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2_t13.log
python scripts/patch_config4.py --iters 10
IPSR_WIDE_NO_PREFETCH=1 python scripts/patch_config4.py --iters 10
python scripts/patch_config4.py --iters 10 --batch 4
python bench.py --workload patch3x3 --steps 30 --no-cpu-baseline > gpurun_out/r2_patch2.json 2> gpurun_out/r2_patch2.err; echo "rc=$?"
python bench.py --workload generator --steps 20 > gpurun_out/r2_gen3.json 2> gpurun_out/r2_gen3.err; echo "gen rc=$?"
