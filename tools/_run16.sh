mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py c5 1024 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; tail -c 400 gpurun_out/c5.err; cat gpurun_out/c5.json | cut -c1-900
timeout 900 python tools/bench_configs.py c4 1024 > gpurun_out/c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; tail -c 600 gpurun_out/c4.err; cat gpurun_out/c4.json | cut -c1-1200
