set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 300 python bench.py --steps 50 --input int16 --no-gpu-reference --no-cpu-baseline > gpurun_out/r2_b9_int16.log 2>&1; tail -c 1800 gpurun_out/r2_b9_int16.log | head -c 1500; echo
timeout 400 python bench.py --config 2 --steps 30 > gpurun_out/r2_b9_c2.log 2>&1; tail -c 600 gpurun_out/r2_b9_c2.log; echo
timeout 400 python bench.py --config 3 --steps 10 > gpurun_out/r2_b9_c3.log 2>&1; tail -c 600 gpurun_out/r2_b9_c3.log; echo
timeout 400 python bench.py --config 4 --steps 3 > gpurun_out/r2_b9_c4.log 2>&1; tail -c 1500 gpurun_out/r2_b9_c4.log; echo
timeout 300 python bench.py --impl reference --config 2 --steps 3 --warmup 1 2>&1 | tail -c 700
timeout 300 python bench.py --impl reference --config 4 --steps 3 --warmup 1 2>&1 | tail -c 700
