mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_full.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'], d['cpu_baseline'], d['gpu_launches'])
        print({k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['kernels'].items() if isinstance(v,dict)})
PY
