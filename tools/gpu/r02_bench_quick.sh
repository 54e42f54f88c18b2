mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_quick.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_quick.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2))
        print({k:(round(v['ms'],3), round(v['frac'],3), v.get('launches_per_vcycle')) for k,v in d['kernels'].items() if isinstance(v,dict)})
        print('vcycle', round(d['vcycle']['frac_bytes_min'],3), round(d['vcycle']['frac_bytes_moved'],3))
PY
