set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 50 > gpurun_out/r2_bench5.log 2>&1; tail -c 3000 gpurun_out/r2_bench5.log
