CN_GST_MODE=tc timeout 600 python -m pytest tests/test_gpu_gst.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest52.log
CN_GST_MODE=tc timeout 600 python tools/bench_configs.py --configs c3 --warmup 30 --steps 30 2>&1 | tail -1 | tee gpurun_out/configs52.log
