for v in "IPSR_TC_PERSIST=1 IPSR_TC_PERSIST_VERBOSE=1" "IPSR_TC_PERSIST=0" "IPSR_TC_PERSIST=1" "IPSR_TC_PERSIST=0"; do
env $v timeout 300 python bench.py --batch 64 --size 64 --steps 200 --e2e-steps 20 --no-cpu-baseline --no-also > gpurun_out/x.json 2> gpurun_out/x.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/x.json') if l.startswith('{')][-1])
    print('$v: B %.4f ms corr %.4f frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print('$v failed', e)
PY
grep "ipsr: persistent" gpurun_out/x.err | head -1
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "persistent" 2>&1 | tail -2
