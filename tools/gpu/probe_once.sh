mkdir -p gpurun_out
for cfg in "2048 2048 2" "2048 2048 1" "512 512 5" "1024 1024 3"; do DGB_PROBE_ASSEMBLY=1 timeout 300 python tools/probe_kernels.py $cfg 2 stream:apply 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['Ni'],d['b'],d.get('assemble_poisson_ms'), d.get('assemble_elements_per_s'))"; done
