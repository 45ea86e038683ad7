# round-end check on a GPU box: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/_call.sh'
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_short.json 2> gpurun_out/bench.err; tail -c 200 gpurun_out/bench_short.json
