# round 2, final verification: full GPU test tier, smoke, the bench line as the driver runs it (both arms), set-up profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_full.err
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_ref.err
timeout 600 python tools/profile_setup.py 2048 2 > gpurun_out/profile_setup.txt 2>&1; grep "===" gpurun_out/profile_setup.txt
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])
        print('roofline', d['roofline']['kernel'][:40], round(d['roofline']['frac'],3), 'share', round(d['roofline']['share_of_vcycle'],3), 'traffic', d['roofline']['traffic'])
        print({k:(round(v['ms'],3), round(v['frac'],3), v.get('launches_per_vcycle')) for k,v in d['kernels'].items() if isinstance(v,dict)})
        print('vcycle', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['vcycle'].items()})
        print('solve', d.get('solve'))
        print('cpu', d['cpu_baseline'] and d['cpu_baseline']['measured'], 'parity', d['parity'])
        print('p5', d.get('apply_p5'), 'setup', d['setup_s'], d['assemble_s'], d['smoother_setup_s'])
for l in open('gpurun_out/bench_ref.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ref', d['value'], d['ms_per_step'], d['cpu_baseline']['measured'])
PY
