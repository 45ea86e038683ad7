timeout 120 tools/micro/umma_rate 2>&1 | tee gpurun_out/r2_umma_rate.log
