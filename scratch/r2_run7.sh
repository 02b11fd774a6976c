set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,60000
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py > gpurun_out/r2_dp_check.log 2>&1
tail -60 gpurun_out/r2_dp_check.log
