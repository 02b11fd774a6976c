set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,60000
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s 2>&1 | tail -40
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2_bench6_n2.log 2>&1; tail -c 1500 gpurun_out/r2_bench6_n2.log
