mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained" > gpurun_out/pytest_chain.log 2>&1; tail -15 gpurun_out/pytest_chain.log | cut -c1-300
for cfg in "2048 2048 1" "2048 2048 2" "1024 1024 1" "256 256 1" "64 64 1"; do
  for v in 0 21 22 9; do
    DGB_GS_VARIANT=$v timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe11.jsonl 2>gpurun_out/probe11.err || echo "fail $cfg $v"
  done
done
cut -c1-400 gpurun_out/probe11.jsonl
