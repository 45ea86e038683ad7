timeout 200 tools/micro/bulk_stream 2>&1 | tee gpurun_out/r2_bulk_stream.log
