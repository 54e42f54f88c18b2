mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_slab.json 2> gpurun_out/bench_n${N}_slab.err; echo "slab rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --gs-mode redblack > gpurun_out/bench_n${N}_rb.json 2> gpurun_out/bench_n${N}_rb.err; echo "rb rc=$?"
python - <<PY
import json
for f in ('gpurun_out/bench_n${N}_slab.json','gpurun_out/bench_n${N}_rb.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['vcycle'])
PY
