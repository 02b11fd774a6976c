set -x
cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
python scratch/ncu_step.py bf16 > gpurun_out/plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python scratch/ncu_step.py bf16 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
ncu --set full --clock-control none --import-source on -k regex:"wfdb16_zscore_pack|wgrad_thin|wgrad_tc_kernel|wgrad_tc_reduce|conv_tc_kernel|bn_bwd|bn_fwd|head_fwd_bwd" -s 40 -c 30 -o gpurun_out/r02_full python scratch/ncu_step.py bf16 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/r02_full.ncu-rep
