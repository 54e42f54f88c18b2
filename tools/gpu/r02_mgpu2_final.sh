# round 2, final 2-GPU check: parity of the slab paths, then the bench as the driver launches it (default mode, both arms)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 64 > gpurun_out/mgpu_check.log 2>&1; echo "check rc=$?"; grep -E "mgpu_check|Error|error" gpurun_out/mgpu_check.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n2_final.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['config'].get('gs_mode'), d['config'].get('transport'), d.get('scaling'))
PY
