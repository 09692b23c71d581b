timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t31.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_t31.log
python bench.py > gpurun_out/r2_final_bench_g1.json 2> gpurun_out/r2_final_bench_g1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_final_bench_g1.json') if l.startswith('{')][-1])
a=d['also']; a=a[0] if isinstance(a,list) else a
print('A %.0f img/s %.4f ms e2e %.0f frac %.3f | B %.0f img/s %.4f ms frac %.3f corr %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], a['value'], a['ms_per_step'], a['roofline']['frac'], a['roofline']['kernel_ms']))
PY
