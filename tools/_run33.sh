timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest33.log
timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench33.log
timeout 300 python tools/profile_e2e.py > gpurun_out/e2e_prof33.log 2>&1
