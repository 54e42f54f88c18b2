# round 2: full GPU test tier, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
        print('roofline', d['roofline']['kernel'][:40], round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_vcycle'],3))
        print({k:(round(v['ms'],3), round(v['frac'],3), v.get('launches_per_vcycle')) for k,v in d['kernels'].items() if isinstance(v,dict)})
        print('vcycle', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['vcycle'].items()})
        print('cpu', d['cpu_baseline'] and d['cpu_baseline']['measured'], 'parity', d['parity'])
        print('p5', d.get('apply_p5'), 'mem', d['memory'])
PY
