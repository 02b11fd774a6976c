cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 900 python -m pytest tests/test_gpu_step_engine.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 200 --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_b31.log 2>&1; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b31.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['last_loss'])
PY
