timeout 500 python bench.py > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e'], d['update']['ms'], {k:(v['ms_per_step'], v.get('update',{}).get('ms')) for k,v in d['configs'].items()}, d['cpu_baseline']['value'])
print(d['roofline']); print(d['breakdown_ms'])
PY
tail -3 gpurun_out/r2_bench_1gpu_final.err
