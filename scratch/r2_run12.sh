set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,60000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 scratch/timeline_dp.py 256 > gpurun_out/r2_tl_dp2.log 2>&1; grep -v "^\*\|OMP\|^$" gpurun_out/r2_tl_dp2.log | tail -45
timeout 300 $TR --master-port 29552 scratch/timeline_dp.py 256 sync 2>&1 | grep -E "us/step|bn_sync|span" | head -30
