# 8-GPU evidence runs: /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/_call8.sh'
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
$TR bench.py --gpus 8 --no-cpu-baseline --no-vecenv > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
$TR bench.py --gpus 8 --config 3 --steps 50 --warmup 20 --no-cpu-baseline --no-vecenv > gpurun_out/r2_cfg3_n8.json 2> gpurun_out/r2_cfg3_n8.err
$TR bench.py --gpus 8 --config 5 --steps 30 --warmup 10 --no-cpu-baseline --no-vecenv > gpurun_out/r2_cfg5_n8.json 2> gpurun_out/r2_cfg5_n8.err
$TR tools/ppo_rollout.py --envs 4096 --steps 16 --iters 3 > gpurun_out/r2_cfg4_ppo_n8.jsonl 2> gpurun_out/r2_cfg4_ppo_n8.err
