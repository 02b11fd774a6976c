set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_infer.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/ab_step.py 256 1000
timeout 300 python scratch/cta_span.py 2>&1 | grep -v "^\["
timeout 300 python scratch/bench_kernels.py 256 1000 cnn
