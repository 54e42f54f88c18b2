mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 300 gpurun_out/bench_a.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_a.json',):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items() if isinstance(v,dict)}, d['vcycle']['normalised_residual_after_timed_cycles'], d['gpu_launches'])
PY
