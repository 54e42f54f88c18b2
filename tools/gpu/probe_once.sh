mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/bench_configs.py c4 1024 > gpurun_out/c4.json 2>gpurun_out/c4.err; tail -c 300 gpurun_out/c4.err; cut -c1-1300 gpurun_out/c4.json
