set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2f_tests.log
export PRE=150 K=20
run() { echo "== $*" | tee -a gpurun_out/r2f_regs.log; env "$@" python tools/steady_diag.py acorn 4096 2>&1 | tee -a gpurun_out/r2f_regs.log; }
run GRS_LIB=$PWD/tools/_libD.so
run GRS_LIB=$PWD/tools/_libE.so
run GRS_LIB=$PWD/tools/_libF.so
run GRS_LIB=$PWD/tools/_libC.so
run A=1
python tools/ppo_rollout.py --envs 1024 --steps 8 --iters 2 2>&1 | tail -2 | cut -c1-400
