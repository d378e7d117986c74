timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke44.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench44.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench44_ref.log
