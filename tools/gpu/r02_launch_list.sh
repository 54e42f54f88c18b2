# ncu launch list of the bench command (after the same command exited 0 without ncu)
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/plain_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ms', d['ms_per_step'])
PY
