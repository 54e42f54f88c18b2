mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain_b4.json 2> gpurun_out/bench_chain_b4.err; tail -c 600 gpurun_out/bench_chain_b4.err
DGB_CHAIN_MASK=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain_b4b9.json 2> gpurun_out/bench_chain_b4b9.err
rm -f gpurun_out/probe14.jsonl
for cfg in "1024 1024 3" "512 512 4"; do
  for m in 0 15; do
    DGB_CHAIN_MASK=$m timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe14.jsonl 2>gpurun_out/probe14.err || echo "fail $cfg"
  done
done
python - <<'PY'
import json
for f in ('gpurun_out/bench_chain_b4.json','gpurun_out/bench_chain_b4b9.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['kernels'].items()}, d['vcycle']['normalised_residual_after_timed_cycles'])
for l in open('gpurun_out/probe14.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
