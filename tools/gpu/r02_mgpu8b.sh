mkdir -p gpurun_out
for mr in 8 32 64; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((mr%10)) bench.py --gpus 8 --steps 10 --warmup 3 --min-rows $mr > gpurun_out/bench_n8_mr$mr.json 2> gpurun_out/bench_n8_mr$mr.err; echo "bench mr=$mr rc=$?"
done
python - <<'PY'
import json
for mr in (8,32,64):
    for l in open(f'gpurun_out/bench_n8_mr{mr}.json'):
        if l.startswith('{'):
            d=json.loads(l); print('min_rows', mr, d['n_gpus'], d['config']['gs_mode'], round(d['ms_per_step'],3), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['vcycle']['normalised_residual_after_timed_cycles'])
PY
