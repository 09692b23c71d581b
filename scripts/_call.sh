for v in "IPSR_TC_DBG=0" "IPSR_TC_DBG=1" "IPSR_TC_DBG=2"; do
env $v timeout 300 python bench.py --steps 200 --e2e-steps 20 --no-cpu-baseline > gpurun_out/x.json 2> gpurun_out/x.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x.json') if l.startswith('{')][-1])
a=d.get('also'); a=a[0] if isinstance(a,list) else a
print('$v: A corr %.4f | B corr %.4f' % (d['roofline']['kernel_ms'], a['roofline']['kernel_ms']))
PY
done
