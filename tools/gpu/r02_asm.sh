mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "assembled or synthetic or smoother_only or chained_gauss_seidel_kernel" > gpurun_out/pytest_asm.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_asm.log | cut -c1-300
timeout 600 python bench.py --config c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; tail -c 300 gpurun_out/bench_c4.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_c4.json'):
    if l.startswith('{'):
        d=json.loads(l)['details']; print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items()})
PY
