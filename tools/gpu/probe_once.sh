mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained" > gpurun_out/pytest_chain.log 2>&1; tail -12 gpurun_out/pytest_chain.log | cut -c1-300
