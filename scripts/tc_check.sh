#!/bin/bash
# Correlation-kernel experiment loop: tensor-path parity tests, then bench configs[2] with both MMA shapes and configs[1].
mkdir -p gpurun_out
IPSR_TC_BN256=${1:-1} timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or cascade or tensor_mode or seeded or tcgen05 or golden" 2>&1 | tail -3
for w in 0 1; do
  IPSR_TC_BN256=$w python bench.py --no-cpu-baseline --e2e-steps 10 --no-also --batch 64 --size 64 --steps 50 --warmup 5 > gpurun_out/bench_bn_$w.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/bench_bn_$w.json')); print('B bn256=$w', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
python bench.py --no-cpu-baseline --e2e-steps 10 --no-also > gpurun_out/bench_epi_A.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_epi_A.json')); print('A', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
