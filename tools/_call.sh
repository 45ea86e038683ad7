for v in 0 1; do echo "== GRS_HESS_FP64=$v"; GRS_HESS_FP64=$v python -m pytest tests/test_physics_parity.py -m gpu -q -s -k "matched_states" 2>&1 | grep -E "outlier|\[.*\]|passed|failed" | cut -c1-220; done
export PRE=150 K=30
GRS_HESS_FP64=1 python tools/steady_diag.py acorn 4096 2>&1 | cut -c1-200
python tools/steady_diag.py acorn 4096 2>&1 | cut -c1-200
