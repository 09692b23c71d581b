TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$TR --master-port 29523 bench.py --gpus 4 --steps 500 --no-cpu-baseline > gpurun_out/r2_final_bench_g4.json 2> gpurun_out/r2_final_bench_g4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2_final_bench_g4.json']:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d['e2e']['copies_only_ceiling']['value'])
    a=d.get('also'); a=a[0] if isinstance(a,list) else a
    print('   also', a['value'], a['ms_per_step'], a.get('per_rank_ms_per_step'))
PY
