cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 300 python scratch/cta_span.py 2>&1 | grep wgrad
timeout 300 python scratch/ab_step.py 256 1000 2>&1 | head -1
