timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest19.log
for qb in 1 2 3 4; do CN_ATTN_QB=$qb python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench19_qb$qb.log; done
timeout 300 python tools/bench_configs.py --configs c4 2>&1 | tail -1 | tee gpurun_out/configs19.log
