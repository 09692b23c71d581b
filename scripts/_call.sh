timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t25.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_t25.log
python bench.py > gpurun_out/r2_final_bench_g1.json 2> gpurun_out/r2_final_bench_g1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref_g1.json 2> gpurun_out/r2_final_ref_g1.err; echo "ref rc=$?"
python bench.py --workload patch3x3 --steps 40 > gpurun_out/r2_final_patch_g1.json 2> gpurun_out/r2_final_patch_g1.err; echo "patch rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r2_final_bench_g1.json','gpurun_out/r2_final_patch_g1.json','gpurun_out/r2_final_ref_g1.json']:
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    a=d.get('also')
    if a:
        a=a[0] if isinstance(a,list) else a
        print('   also', a['value'], a['ms_per_step'], a['roofline']['frac'], a['roofline']['kernel_ms'])
PY
