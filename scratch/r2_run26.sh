set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2>gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_reference_arm.json
timeout 400 python bench.py --input fp32 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_n1_fp32_input.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_n1_fp32_input.json
timeout 400 python bench.py --config 2 > gpurun_out/r02_bench_c2_n1.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_c2_n1.json
timeout 400 python bench.py --config 3 > gpurun_out/r02_bench_c3_n1.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_c3_n1.json
timeout 400 python bench.py --config 4 > gpurun_out/r02_bench_c4_n1.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_c4_n1.json
timeout 400 python bench.py --precision fp32 --input fp32 --steps 30 --no-gpu-reference --no-cpu-baseline > gpurun_out/r02_bench_n1_fp32_engine.json 2>gpurun_out/fp32.err; tail -c 300 gpurun_out/r02_bench_n1_fp32_engine.json; tail -3 gpurun_out/fp32.err
