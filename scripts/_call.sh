timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_ragged.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2_t27.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_t27.log
for v in new old new old; do
cp deepinpainting_b200/lib/$v.so.bin deepinpainting_b200/lib/libipsr_sm100.so
timeout 300 python bench.py --steps 500 --e2e-steps 20 --no-cpu-baseline > gpurun_out/x.json 2> gpurun_out/x.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x.json') if l.startswith('{')][-1])
a=d.get('also'); a=a[0] if isinstance(a,list) else a
print('$v: A %.4f ms corr %.4f | B %.4f ms corr %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms'], a['ms_per_step'], a['roofline']['kernel_ms']))
PY
done
python scripts/chaos_bwd.py 64 256 64 3
