python tools/timeline.py 2>&1 | tail -1 > gpurun_out/timeline23.log
CN_NO_SIDE_STREAM=1 python tools/timeline.py 2>&1 | tail -1 > gpurun_out/timeline23_noside.log
