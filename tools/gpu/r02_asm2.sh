mkdir -p gpurun_out; rm -f gpurun_out/probe_asm.jsonl
for v in 0 51; do for cfg in "512 512 5" "1024 1024 3"; do
  DGB_PROBE_ASSEMBLY=1 DGB_GS_VARIANT=$v timeout 300 python tools/probe_kernels.py $cfg 2 stream:apply >> gpurun_out/probe_asm.jsonl 2>gpurun_out/probe_asm.err || echo "fail $cfg"
done; done
python - <<'PY'
import json
for l in open('gpurun_out/probe_asm.jsonl'):
    d=json.loads(l); print(d['Ni'],d['p'],d['b'],round(d['assemble_poisson_ms'],2),'ms', round(d['assemble_elements_per_s']/1e6,3),'M el/s')
PY
