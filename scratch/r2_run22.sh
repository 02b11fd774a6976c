cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 600 python -m pytest tests/test_wfdb16.py tests/test_loader.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 100 --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_b22.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b22.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])
print([(r['call'], r['us']) for r in d['layers'][:4]])
PY
