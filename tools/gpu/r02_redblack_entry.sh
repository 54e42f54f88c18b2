# round 2: the fused 2-colour entry (residual + first colour): its tests, then the 1-GPU red-black bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "colour_entry or redblack or streaming_kernels or synthetic or midsize" > gpurun_out/pytest_rb.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_rb.log | cut -c1-300
for v in 0 43; do
DGB_GS_VARIANT=$v timeout 600 python bench.py --gs-mode redblack --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rb_v$v.json 2> gpurun_out/bench_rb_v$v.err; echo "bench v$v rc=$?"; tail -c 300 gpurun_out/bench_rb_v$v.err
done
python - <<'PY'
import json
for v in (0, 43):
    for l in open(f'gpurun_out/bench_rb_v{v}.json'):
        if l.startswith('{'):
            d=json.loads(l); print('variant', v, 'ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], d['vcycle'].get('normalised_residual_after_timed_cycles'))
PY
