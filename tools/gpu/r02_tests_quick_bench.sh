# round 2: the chain / residual tests, then the quick bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained or streaming or smoother or midsize or synthetic or golden" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-300
bash tools/gpu/r02_bench_quick.sh
