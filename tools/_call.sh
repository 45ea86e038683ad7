ncu --set full --clock-control none -k regex:"k_conv1_ws|k_conv_direct" --launch-skip 9 -c 3 -o gpurun_out/r2_policy_convs python tools/policy_bench.py 4096 > gpurun_out/r2_ncu_convs.log 2>&1
