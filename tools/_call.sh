set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2c_tests.log
export PRE=150 K=30
for cfg in "GRS_GAP_SKIP=0" "GRS_GAP_SKIP=1" "GRS_GAP_SKIP=1 GRS_SLOT_ORDER=1" "GRS_GAP_SKIP=1 GRS_SLOT_ORDER=1 GRS_ORDER_NCON=-1"; do
  echo "== $cfg" | tee -a gpurun_out/r2c_steady.log
  env $cfg python tools/steady_diag.py acorn 4096,16384 2>&1 | tee -a gpurun_out/r2c_steady.log
done
