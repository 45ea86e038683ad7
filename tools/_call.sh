set -x
export PRE=150 K=6
ncu --set full --clock-control none --import-source on -k regex:k_env_step_ls --launch-skip 152 -c 1 -o gpurun_out/r2_env_step python tools/steady_diag.py acorn 4096 > gpurun_out/r2_ncu_env.log 2>&1
python bench.py --steps 2 --warmup 1 --preroll 150 --no-cpu-baseline --no-vecenv > gpurun_out/r2_bench_short.json 2>gpurun_out/r2_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --preroll 150 --no-cpu-baseline --no-vecenv > gpurun_out/r2_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_layer|k_conv|k_mlp" --launch-skip 15 -c 5 --csv --log-file gpurun_out/r2_policy_launches.csv python tools/policy_bench.py 4096 > gpurun_out/r2_ncu_policy.log 2>&1
tail -2 gpurun_out/r2_ncu_policy.log
