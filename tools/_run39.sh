timeout 600 python -m pytest tests/test_gpu_gst.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest39.log
