mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained" > gpurun_out/pytest_chain.log 2>&1; tail -8 gpurun_out/pytest_chain.log | cut -c1-300
rm -f gpurun_out/probe16.jsonl
for cl in 1 8; do
for cfg in "2048 2048 1" "1024 1024 1" "256 256 1" "64 64 1" "2048 2048 2" "512 512 4"; do
  DGB_CHAIN_CLUSTER=$cl DGB_CHAIN_MASK=15 DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe16.jsonl 2>gpurun_out/probe16.err || echo "fail $cfg"
done
done
python - <<'PY'
import json
for l in open('gpurun_out/probe16.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
