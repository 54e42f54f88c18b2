mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
rm -f gpurun_out/probe18.jsonl
for cfg in "2048 2048 2" "2048 2048 1" "512 512 5" "1024 1024 3"; do
  timeout 300 python tools/probe_kernels.py $cfg 5 stream >> gpurun_out/probe18.jsonl 2>gpurun_out/probe18.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe18.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],{k:v for k,v in d.items() if k.startswith('stream.')}, d['device_error'])
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 300 gpurun_out/bench_a.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_a.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['vcycle']['normalised_residual_after_timed_cycles'])
        print({k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['kernels'].items() if isinstance(v,dict)})
PY
