mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py c5 1024 > gpurun_out/c5.json 2>/dev/null; cut -c1-700 gpurun_out/c5.json
timeout 900 python tools/bench_configs.py c4 1024 > gpurun_out/c4.json 2>/dev/null; cut -c1-1100 gpurun_out/c4.json
