# BASELINE.json configs[2..4] on one GPU (short runs) -> gpurun_out/cfg_*.json ; development aid, not the driver's bench
set -x
python bench.py --scene sugar_cube --direction 45 --envs 8192 --steps 20 --warmup 8 --no-cpu-baseline > gpurun_out/cfg3_sugar_cube_8192.json 2> gpurun_out/cfg3.err
python bench.py --scene bread_crumb --envs 4096 --policy --steps 20 --warmup 8 --no-cpu-baseline > gpurun_out/cfg4_bread_crumb_policy.json 2> gpurun_out/cfg4.err
python bench.py --scene sand_ball --envs 16384 --steps 12 --warmup 6 --no-cpu-baseline > gpurun_out/cfg5_sand_ball_16384.json 2> gpurun_out/cfg5.err
for f in gpurun_out/cfg3_sugar_cube_8192.json gpurun_out/cfg4_bread_crumb_policy.json gpurun_out/cfg5_sand_ball_16384.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']/1e6,2),'M substeps/s', round(d['transitions_per_s']), 'transitions/s', round(d['ms_per_step'],2),'ms/step', round(d['substeps_per_transition'],1),'sub/tr kernel_ms', round(d['roofline']['kernel_ms'],2))"; done
