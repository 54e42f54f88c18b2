mkdir -p gpurun_out; rm -f gpurun_out/probe_var.jsonl
for v in 22 31 32 33 34; do
  DGB_GS_VARIANT=$v timeout 300 python tools/probe_kernels.py 2048 8 1 5 stream:gs_fwd >> gpurun_out/probe_var.jsonl 2>gpurun_out/probe_var.err || echo "fail $v"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe_var.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
