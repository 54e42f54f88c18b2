# round 2: after dropping the helper's 8-byte d store into the opposite stream: chain / smoother tests, the entry kernel alone, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained or streaming or smoother or midsize or synthetic or history or solve_loop or slab" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-300
timeout 300 python tools/probe_kernels.py 2048 2048 2 5 stream:entry_residual 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('entry b9', d['stream.entry_residual'])"
timeout 300 python tools/probe_kernels.py 2048 2048 1 5 stream:entry_residual 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('entry b4', d['stream.entry_residual'])"
bash tools/gpu/r02_bench_quick.sh
