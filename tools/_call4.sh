set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus 4 --no-cpu-baseline --no-vecenv > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
$TR bench.py --gpus 4 --config 3 --steps 50 --warmup 20 --no-cpu-baseline --no-vecenv > gpurun_out/r2_cfg3_n4.json 2> gpurun_out/r2_cfg3_n4.err
