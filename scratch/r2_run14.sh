set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_step_engine.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/bench_kernels.py 256 1000 cnn 2>&1 | grep -E "wgrad|bn_bwd|sum"
timeout 300 python scratch/ab_step.py 256 1000 2>&1 | head -1
timeout 300 python scratch/timeline.py 256 1000 2>&1 | grep -E "wgrad|bn_bwd|dgrad|span"
