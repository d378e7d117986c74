timeout 600 python -m pytest tests/test_gpu_gst.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest41.log
timeout 600 python tools/bench_configs.py --configs c3 --warmup 30 --steps 30 2>&1 | tail -1 | tee gpurun_out/configs41.log
