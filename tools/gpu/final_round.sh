# final verification of a round: full GPU test tier, smoke(), bench (default + variants), then the ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --check-residual 0 > gpurun_out/bench_nocheck.json 2>/dev/null; echo "nocheck rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --gs-mode redblack > gpurun_out/bench_redblack.json 2>/dev/null; echo "redblack rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ('bench_full','bench_nocheck','bench_redblack'):
    for l in open(f'gpurun_out/{f}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(f, round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['roofline']['kernel'][:30], round(d['roofline']['frac'],3), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('value'))
PY
