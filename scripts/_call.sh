for v in "IPSR_TC_PAIRS=1" "IPSR_TC_PAIRS=0" "IPSR_TC_ARES=0" "IPSR_TC_PAIRS=0 IPSR_TC_ARES=0" "IPSR_TC_PRODUCERS=1"; do
env $v python bench.py --steps 300 --e2e-steps 20 --no-cpu-baseline --no-also > gpurun_out/x.json 2> gpurun_out/x.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/x.json') if l.startswith('{')][-1])
print('$v: A %.4f ms corr %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))
PY
done
