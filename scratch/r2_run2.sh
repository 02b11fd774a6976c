set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000
timeout 300 python scratch/timeline.py 256 1000 > gpurun_out/r2_timeline_b256.log 2>&1; cat gpurun_out/r2_timeline_b256.log
timeout 120 scratch/bin/mma_bench > gpurun_out/r2_mma_bench.log 2>&1; tail -22 gpurun_out/r2_mma_bench.log
