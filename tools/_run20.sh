timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_gemm_tc.py tests/test_gpu_train_loop.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest20.log
python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench20.log
python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench20b.log
