# 8 GPUs: red-black parity at world 4 + the bench at N = 2, 4, 8 (native transport)
mkdir -p gpurun_out
echo skip check4
for n in 2 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --gs-mode redblack > gpurun_out/bench_n${n}_redblack.json 2> gpurun_out/bench_n${n}_redblack.err; echo "bench $n rc=$?"
done
true
python - <<'PY'
import json
for n in (1,2,8):
    for l in open(f'gpurun_out/bench_n{n}_redblack.json'):
        if l.startswith('{'):
            d=json.loads(l); print('redblack', d['n_gpus'], round(d['ms_per_step'],2), round(d['value'],2), round(d['e2e']['value'],2), d['gpu_launches'], d['vcycle']['normalised_residual_after_timed_cycles'])
PY
