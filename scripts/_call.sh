timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_patches.py -m gpu -x -q > gpurun_out/r2_t28.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_t28.log
for i in 1 2; do
timeout 300 python bench.py --steps 300 --e2e-steps 20 --no-cpu-baseline > gpurun_out/x.json 2> gpurun_out/x.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x.json') if l.startswith('{')][-1])
a=d.get('also'); a=a[0] if isinstance(a,list) else a
print('A %.4f ms corr %.4f | B %.4f ms corr %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], a['ms_per_step'], a['roofline']['kernel_ms']))
PY
done
