timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_t20.log
for p in 1 0 1 0; do
IPSR_PDL=$p python bench.py --steps 500 --e2e-steps 20 --no-cpu-baseline > gpurun_out/r2_bench8_pdl$p.json 2> gpurun_out/r2_bench8_pdl$p.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench8_pdl$p.json') if l.startswith('{')][-1])
a=d['also']; a=a[0] if isinstance(a,list) else a
print('PDL $p: A %.4f ms corr %.4f | B %.4f ms corr %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], a['ms_per_step'], a['roofline']['kernel_ms']))
PY
done
