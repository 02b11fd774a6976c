set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 600 python -m pytest tests/test_wfdb16.py tests/test_loader.py tests/test_gpu_bf16.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 50 --input int16 --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_b10_int16.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b10_int16.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])
print([(r['call'], r['us']) for r in d['layers'][:4]])
PY
