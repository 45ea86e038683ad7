# development aid: substep throughput by scene and batch size
for cfg in "acorn 8192" "acorn 16384" "sugar_cube 4096" "sand_ball 4096" "bread_crumb 4096" "sugar_cube 16384"; do
  set -- $cfg
  echo -n "$1 N=$2: "
  python bench.py --scene $1 --envs $2 --steps 12 --warmup 6 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', round(d['ms_per_step'],2),'ms', round(d['substeps_per_transition'],1),'sub/tr')"
done
