set -x
export PRE=150 K=40
run() { echo "== $*" | tee -a gpurun_out/r2q_long_per_block.log; env "$@" python tools/steady_diag.py acorn 4096 2>&1 | tee -a gpurun_out/r2q_long_per_block.log; }
run GRS_SLOT_ORDER=0
run GRS_SLOT_ORDER=3 GRS_LONG_PER_BLOCK=7
run GRS_SLOT_ORDER=3 GRS_LONG_PER_BLOCK=6
run GRS_SLOT_ORDER=3 GRS_LONG_PER_BLOCK=5
run GRS_SLOT_ORDER=3 GRS_LONG_PER_BLOCK=4
