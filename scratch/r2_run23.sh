set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,120000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_n8b_c1.log 2>&1; tail -c 300 gpurun_out/r2_n8b_c1.log; echo
timeout 300 $TR --master-port 29543 bench.py --gpus 8 --config 2 --steps 100 --warmup 5 --no-gpu-reference > gpurun_out/r2_n8b_c2.log 2>&1; tail -c 300 gpurun_out/r2_n8b_c2.log; echo
timeout 300 $TR --master-port 29544 bench.py --gpus 8 --config 3 --steps 50 --warmup 5 --no-gpu-reference > gpurun_out/r2_n8b_c3.log 2>&1; tail -c 300 gpurun_out/r2_n8b_c3.log; echo
timeout 300 $TR --master-port 29545 bench.py --gpus 8 --config 4 --steps 3 --warmup 3 --no-gpu-reference > gpurun_out/r2_n8b_c4.log 2>&1; tail -c 300 gpurun_out/r2_n8b_c4.log; echo
