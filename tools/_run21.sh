timeout 600 python -m pytest tests/test_gpu_policy.py tests/test_gpu_rollout.py -x -q 2>&1 | tail -3 | tee gpurun_out/pytest21.log
python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench21.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cn_hh_attention|cn_embed1|cn_row_offsets|cn_pack_inputs|cn_hr_attention|cn_gru_gate|cn_heads" --launch-skip 420 --launch-count 7 -f -o gpurun_out/r1_policy_small_v7 python tools/bench_configs.py --configs c2 --warmup 60 --steps 3 > gpurun_out/ncu21.log 2>&1
