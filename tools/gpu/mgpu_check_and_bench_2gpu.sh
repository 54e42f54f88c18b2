mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 64 > gpurun_out/mgpu_check.log 2>&1; echo "check rc=$?"; grep mgpu_check gpurun_out/mgpu_check.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_slab.json 2> gpurun_out/bench_n2_slab.err; echo "bench slab rc=$?"; tail -c 300 gpurun_out/bench_n2_slab.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --gs-mode redblack > gpurun_out/bench_n2_rb.json 2> gpurun_out/bench_n2_rb.err; echo "bench rb rc=$?"
grep -h '^{' gpurun_out/bench_n2_slab.json gpurun_out/bench_n2_rb.json | cut -c1-700
