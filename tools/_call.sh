set -x
export PRE=150 K=30
python tools/steady_diag.py acorn 2048,4096,8192,16384 2>&1 | tee gpurun_out/r2_scene_sweep.log
for sc in sugar_cube sand_ball bread_crumb; do python tools/steady_diag.py $sc 4096 2>&1 | tee -a gpurun_out/r2_scene_sweep.log; done
python tools/steady_diag.py sugar_cube 16384 2>&1 | tee -a gpurun_out/r2_scene_sweep.log
