timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 90 --warmup 10 --no-cpu-baseline > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','n_gpus']}, d['e2e']['value'], d['update']['ms'], d['update']['ms_samples'], d['update']['allreduce_ms'], {k:(v['ms_per_step'], v.get('update',{}).get('ms')) for k,v in d['configs'].items()})
PY
tail -2 gpurun_out/r2_bench_2gpu_final.err
