timeout 400 python tools/profile_update.py 2>&1 | tail -36 > gpurun_out/update_prof48.log
timeout 400 python tools/bench_update.py 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/update48.log
timeout 400 python tools/bench_update.py --tf32 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/update48_tf32.log
