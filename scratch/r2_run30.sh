set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,60000
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02_bench_n2.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_n2.json
timeout 400 $TR --master-port 29534 bench.py --gpus 2 --config 2 --steps 100 --warmup 5 --no-gpu-reference > gpurun_out/r02_bench_c2_n2.json 2>/dev/null; tail -c 200 gpurun_out/r02_bench_c2_n2.json
timeout 300 $TR --master-port 29551 scratch/timeline_dp.py 256 > gpurun_out/r02_timeline_dp2.log 2>&1; grep -E "us/step|span" gpurun_out/r02_timeline_dp2.log
timeout 300 $TR --master-port 29552 scratch/timeline_dp.py 256 sync 2>&1 | grep -E "us/step|span"
