# round 2: the bench as the driver launches it at N GPUs (default mode)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_final.json 2> gpurun_out/bench_n${N}_final.err; echo "bench rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n${N}_final.err | tail -c 600
python - <<PY
import json
for l in open('gpurun_out/bench_n${N}_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['config'].get('gs_mode'), d.get('scaling'), d.get('clocks'))
PY
