export PRE=150 K=40
run() { echo "== $*" | tee -a gpurun_out/r2zl_helper.log; env "$@" python tools/steady_diag.py acorn 4096 2>&1 | tee -a gpurun_out/r2zl_helper.log; }
run GRS_HELPER_WARPS=0
run GRS_HELPER_WARPS=1
run GRS_HELPER_WARPS=0
run GRS_HELPER_WARPS=1
