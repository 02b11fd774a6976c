set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 100 > gpurun_out/r2_b20.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b20.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['speedup_vs_gpu_reference'])
print([(r['call'], r['us'], r['frac']) for r in d['layers']])
PY
