(time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") 2>&1 | tail -4 | tee gpurun_out/smoke43.log
(time timeout 900 python bench.py) 2>&1 | tail -4 > gpurun_out/bench43.log
(time timeout 900 python bench.py --impl reference --steps 20 --warmup 3) 2>&1 | tail -4 > gpurun_out/bench43_ref.log
