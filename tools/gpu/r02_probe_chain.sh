# round 2: chained GS kernel -- tests first, then chain-only (DGB_GS_VARIANT=22) and full-pass timings per level size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained or streaming or smoother" > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_chain.log | cut -c1-300
rm -f gpurun_out/probe_chain.jsonl
for cfg in "2048 2048 2" "2048 2048 1" "1024 1024 1" "512 512 1" "128 128 1" "2048 8 1" "2048 3 2"; do
  DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe_chain.jsonl 2>gpurun_out/probe_chain.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe_chain.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
