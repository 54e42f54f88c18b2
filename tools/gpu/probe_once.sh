mkdir -p gpurun_out
for cfg in "2048 2048 1" "1024 1024 3" "512 512 5" "2048 2048 2"; do timeout 300 python tools/probe_kernels.py $cfg 5 stream 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['Ni'],d['b'],{k.split('.')[1]:(v['ms'],v['GB/s']) for k,v in d.items() if k.startswith('stream.')})"; done
