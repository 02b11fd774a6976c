set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scratch/ncu_step.py bf16 > gpurun_out/plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python scratch/ncu_step.py bf16 > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log
ncu --set full --clock-control none --import-source on -k regex:"wfdb16_zscore_pack|wgrad_thin|wgrad_tc_kernel|wgrad_tc_reduce|conv_tc_kernel|bn_bwd|bn_fwd|head_fwd_bwd" -s 58 -c 29 -o gpurun_out/r02_full python scratch/ncu_step.py bf16 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 300 python scratch/timeline.py 256 1000 > gpurun_out/r02_timeline_b256.log 2>&1; tail -3 gpurun_out/r02_timeline_b256.log
