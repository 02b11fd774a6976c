set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python scratch/bench_kernels.py 256 1000 cnn > gpurun_out/r2_kernels_b256.log 2>&1; tail -40 gpurun_out/r2_kernels_b256.log
timeout 300 python scratch/trace_conv.py > gpurun_out/r2_trace_conv.log 2>&1; cat gpurun_out/r2_trace_conv.log
timeout 300 python scratch/trace_wgrad.py > gpurun_out/r2_trace_wgrad.log 2>&1; cat gpurun_out/r2_trace_wgrad.log
timeout 120 scratch/bin/mma_bench > gpurun_out/r2_mma_bench.log 2>&1; cat gpurun_out/r2_mma_bench.log
timeout 600 python scratch/gpu_ref.py > gpurun_out/r2_gpu_ref.log 2>&1; cat gpurun_out/r2_gpu_ref.log
