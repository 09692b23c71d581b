python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_tm2.log 2>&1; echo "multi rc=$?"; tail -n 3 gpurun_out/r2_tm2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29523 bench.py --gpus 2 --steps 300 --no-cpu-baseline > gpurun_out/r2_final_bench_g2.json 2> gpurun_out/r2_final_bench_g2.err; echo "bench2 rc=$?"
$TR --master-port 29521 bench.py --gpus 2 --workload patch3x3 --shard bank --steps 40 > gpurun_out/r2_final_patch_g2.json 2> gpurun_out/r2_final_patch_g2.err; echo "patch2 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2_final_bench_g2.json','gpurun_out/r2_final_patch_g2.json']:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))
    a=d.get('also')
    if a:
        a=a[0] if isinstance(a,list) else a
        print('   also', a['value'], a['ms_per_step'], a.get('per_rank_ms_per_step'))
PY
