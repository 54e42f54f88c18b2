# 2 GPUs: parity check of the slab paths, then the bench in both N>1 modes
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 64 > gpurun_out/mgpu_check.log 2>&1; echo "check rc=$?"; grep -E "mgpu_check|Error|error" gpurun_out/mgpu_check.log | tail -12
for mode in redblack slab_lexicographic; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --gs-mode $mode > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"; tail -c 400 gpurun_out/bench_n2_$mode.err
done
python - <<'PY'
import json
for m in ('redblack','slab_lexicographic'):
    for l in open(f'gpurun_out/bench_n2_{m}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(m, d['n_gpus'], round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['vcycle'], d['config'].get('transport'))
PY
