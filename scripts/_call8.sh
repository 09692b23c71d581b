python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_tm8.log 2>&1; echo "multi rc=$?"; tail -n 3 gpurun_out/r2_tm8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --workload patch3x3 --shard bank --steps 40 > gpurun_out/r2_patch_g8.json 2> gpurun_out/r2_patch_g8.err; echo "patch8 rc=$?"
$TR --master-port 29522 bench.py --gpus 8 --workload generator --steps 20 > gpurun_out/r2_gen_g8.json 2> gpurun_out/r2_gen_g8.err; echo "gen8 rc=$?"
$TR --master-port 29523 bench.py --gpus 8 --steps 500 > gpurun_out/r2_bench_g8.json 2> gpurun_out/r2_bench_g8.err; echo "bench8 rc=$?"
tail -n 3 gpurun_out/r2_bench_g8.err gpurun_out/r2_gen_g8.err gpurun_out/r2_patch_g8.err | grep -v "^\*\|OMP"
