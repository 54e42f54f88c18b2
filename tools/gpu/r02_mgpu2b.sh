mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "redblack or slab or streaming" > gpurun_out/pytest_rb.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_rb.log | cut -c1-300
bash tools/gpu/r02_mgpu2.sh
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --p5-apply 0 --gs-mode redblack > gpurun_out/bench_n1_redblack.json 2>/dev/null; python -c "
import json
for l in open('gpurun_out/bench_n1_redblack.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=1 redblack', round(d['ms_per_step'],2), d['gpu_launches'])"
