mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
for m in 1 3; do
DGB_CHAIN_MASK=$m timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err; tail -c 300 gpurun_out/bench_m$m.err
done
python - <<'PY'
import json
for f in ('gpurun_out/bench_m1.json','gpurun_out/bench_m3.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items()}, d['vcycle']['normalised_residual_after_timed_cycles'], d['gpu_launches'])
PY
