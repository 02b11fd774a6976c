set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=5000,30000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29561 scratch/dp_latency.py 2>&1 | grep "us / launch\|rror"
timeout 600 $TR --master-port 29511 tests/dp_check.py 2>&1 | grep -E "^[0-9]\.|^3b|dp_check|rror"
timeout 300 $TR --master-port 29551 scratch/timeline_dp.py 256 2>&1 | grep -E "us/step|dp_adamw|dgrad|span"
timeout 300 $TR --master-port 29552 scratch/timeline_dp.py 256 sync 2>&1 | grep -E "us/step|span"
