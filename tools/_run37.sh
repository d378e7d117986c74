timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest37a.log
timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py tests/test_gpu_train_loop.py tests/test_gpu_eval.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest37.log
timeout 300 python tools/timeline.py 2>&1 | tail -1 > gpurun_out/timeline37_pdl.log
CN_PDL=0 timeout 300 python tools/timeline.py 2>&1 | tail -1 > gpurun_out/timeline37_nopdl.log
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench37.log
