# chain-kernel-only timings (DGB_GS_VARIANT=22) of single bands and full grids, after the chained-kernel tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained" > gpurun_out/pytest_chain.log 2>&1; tail -3 gpurun_out/pytest_chain.log | cut -c1-300
rm -f gpurun_out/probe_chain.jsonl
for cfg in "2048 8 2" "2048 2048 2" "1024 1024 3" "2048 2048 1"; do
  DGB_CHAIN_MASK=15 DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe_chain.jsonl 2>gpurun_out/probe_chain.err || echo "fail $cfg"
done
python - <<'PY'
import json
for l in open('gpurun_out/probe_chain.jsonl'):
    d=json.loads(l); print(d['Ni'],d['Nj'],d['b'],d.get('stream.gs_fwd'), d['device_error'])
PY
