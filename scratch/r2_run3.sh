set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000
timeout 120 scratch/bin/mma_bench > gpurun_out/r2_mma_bench.log 2>&1; tail -13 gpurun_out/r2_mma_bench.log
for ncc in 4 2 1; do
ECGB200_WGRAD_NCC=$ncc timeout 300 python scratch/bench_kernels.py 256 1000 cnn 2>&1 | grep -E "wgrad|sum"
done
ECGB200_WGRAD_NCC=2 timeout 300 python scratch/timeline.py 256 1000 2>&1 | grep -E "wgrad|bn_bwd|dgrad|span"
