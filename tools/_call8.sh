set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
$TR bench.py --gpus 8 --no-cpu-baseline --no-vecenv > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
