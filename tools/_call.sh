for v in 1 2 3 4 5; do echo "== ablation $v"; GRS_LIB=$PWD/tools/_libA$v.so GRP_EVENTS=1 timeout 120 python tools/policy_bench.py 4096 2>&1 | grep grp_forward | tail -1; done
