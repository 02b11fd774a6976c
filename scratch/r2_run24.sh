cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,120000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for mode in fused barrier nccl fused; do
timeout 200 $TR --master-port 29551 scratch/timeline_dp.py 256 nosync $mode 2>&1 | grep -E "us/step|dp_adamw|span|rror"
done
