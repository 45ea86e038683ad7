set -x
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_short.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --config 4 --steps 50 --warmup 20 --no-cpu-baseline --no-vecenv > gpurun_out/r2_cfg4_n1.json 2> gpurun_out/r2_cfg4_n1.err
python bench.py --config 5 --steps 30 --warmup 10 --no-cpu-baseline --no-vecenv > gpurun_out/r2_cfg5_n1.json 2> gpurun_out/r2_cfg5_n1.err
export PRE=150 K=6
ncu --set full --clock-control none --import-source on -k regex:k_env_step_ls --launch-skip 153 -c 1 -o gpurun_out/r2_env_step python tools/steady_diag.py acorn 4096 > gpurun_out/r2_ncu_env.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --preroll 150 --no-cpu-baseline --no-vecenv > gpurun_out/r2_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_layer|k_conv|k_mlp" --launch-skip 15 -c 5 --csv --log-file gpurun_out/r2_policy_launches.csv python tools/policy_bench.py 4096 > gpurun_out/r2_ncu_policy.log 2>&1
python tools/stage_timing.py 150 2>&1 | tee gpurun_out/r2_stage_timing.log
tail -c 300 gpurun_out/r2_bench_n1.json
