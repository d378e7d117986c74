timeout 600 python -m pytest tests/test_gpu_gst.py tests/test_gpu_eval.py -x -q 2>&1 | tail -6 | tee gpurun_out/pytest53.log
timeout 600 python tools/bench_configs.py --configs c3 --warmup 30 --steps 30 2>&1 | tail -1 | tee gpurun_out/configs53.log
