timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; echo tests rc=$?; tail -2 gpurun_out/r2_final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r2_final_smoke.log
timeout 300 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo bench rc=$?
timeout 120 python scratch/ncu_step.py bf16 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c_launches.csv python scratch/ncu_step.py bf16 > gpurun_out/ncu1.log 2>&1; echo ncu1 rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wfdb16_zscore_pack|wgrad_thin|wgrad_pair|wgrad_tc_kernel|wgrad_tc_reduce|conv_tc|bn_bwd|bn_fwd|head_fwd_bwd" -s 56 -c 28 -o gpurun_out/r02c_full python scratch/ncu_step.py bf16 > gpurun_out/ncu2.log 2>&1; echo ncu2 rc=$?
