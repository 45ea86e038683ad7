set -x
export PRE=150 K=20
run() { echo "== $*" | tee -a gpurun_out/r2k_flags.log; env "$@" python tools/steady_diag.py acorn 4096 2>&1 | tee -a gpurun_out/r2k_flags.log; }
run GRS_LIB=$PWD/tools/_libB.so
GRS_LIB=$PWD/tools/_libB.so python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r2k_tests_libB.log
