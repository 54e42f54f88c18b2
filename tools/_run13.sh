mkdir -p gpurun_out

rm -f gpurun_out/probe19.jsonl
for cfg in "2048 4 2" "2048 8 2" "2048 64 2" "2048 2048 2" "2048 2048 1" "2048 24 1" "2048 8 1"; do
  DGB_CHAIN_MASK=15 DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe19.jsonl 2>gpurun_out/probe19.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe19.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
