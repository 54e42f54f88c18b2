mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for cfg in "1024 1024 3" "512 512 5"; do timeout 300 python tools/probe_kernels.py $cfg 5 stream 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['Ni'],d['b'],{k.split('.')[1]:(v['ms'],v['GB/s']) for k,v in d.items() if k.startswith('stream.') and k.split('.')[1] in ('apply','residual','jacobi','redblack')})"; done
timeout 600 python tools/bench_configs.py c5 1024 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('C5', d['apply_ms'], d['apply_GBs'], d['apply_frac'])"
