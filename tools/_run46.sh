timeout 400 python tools/bench_update.py 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/update46_pack.log
timeout 400 python tools/bench_update.py --tf32 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/update46_pack_tf32.log
