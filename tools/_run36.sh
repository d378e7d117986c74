timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest36.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9400 -c 230 --csv --log-file gpurun_out/launches36.csv python tools/bench_configs.py --configs c2 --warmup 420 --steps 10 > gpurun_out/ncu36.log 2>&1
