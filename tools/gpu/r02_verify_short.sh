# round 2: full GPU test tier, smoke, short bench line (after a rebuild)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu/r02_bench_quick.sh
