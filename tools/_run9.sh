mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-400
rm -f gpurun_out/probe17.jsonl
for cfg in "1024 1024 3" "512 512 4" "2048 2048 2" "2048 2048 1"; do
  for m in 0 15; do
    DGB_CHAIN_MASK=$m timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe17.jsonl 2>gpurun_out/probe17.err || echo "fail $cfg"
  done
  DGB_CHAIN_MASK=15 DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe17.jsonl 2>gpurun_out/probe17.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe17.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
