mkdir -p gpurun_out; rm -f gpurun_out/probe_helper.jsonl
for cfg in "2048 2048 2" "2048 2048 1"; do
  timeout 300 python tools/probe_kernels.py $cfg 5 stream:rec_residual >> gpurun_out/probe_helper.jsonl 2>gpurun_out/probe_helper.err || echo "fail $cfg"
  timeout 300 python tools/probe_kernels.py $cfg 5 stream:entry_residual >> gpurun_out/probe_helper.jsonl 2>gpurun_out/probe_helper.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe_helper.jsonl'):
    d=json.loads(l); print(d['Ni'],d['b'],{k:v for k,v in d.items() if k.startswith('stream.')}, d['device_error'])
PY
