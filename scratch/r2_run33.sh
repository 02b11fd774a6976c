cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 900 python -m pytest tests/test_gpu_infer.py -m gpu -x -q -s 2>&1 | grep -E "fp32x3|passed|failed|Error|error" | head -20
timeout 300 python scratch/bench_infer.py 2>&1 | tail -40
