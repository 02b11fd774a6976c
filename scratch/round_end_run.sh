set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s16_bench.log 2>&1; tail -c 600 gpurun_out/s16_bench.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s16_bench_ref.log 2>&1; tail -c 400 gpurun_out/s16_bench_ref.log
python scratch/ncu_step.py bf16 > gpurun_out/plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_v4.csv python scratch/ncu_step.py bf16 > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|wgrad_tc_reduce|conv_tc_kernel|bn_bwd" -s 42 -c 21 -o gpurun_out/r01_full_v4 python scratch/ncu_step.py bf16 > gpurun_out/ncu_full_v4.log 2>&1; tail -2 gpurun_out/ncu_full_v4.log
python scratch/bench_aux.py > gpurun_out/s16_aux.json 2> gpurun_out/s16_aux.err; tail -3 gpurun_out/s16_aux.err
