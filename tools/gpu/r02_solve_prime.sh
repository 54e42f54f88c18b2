# round 2: the solve loop's residual handed to the next cycle (dgb_vcycle_ex): parity tests, then the timing of a full solve
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "history or solve_loop or midsize or cli" 2>&1 | tail -3
timeout 600 python tools/solve_time.py 2048 2 > gpurun_out/solve_time.json 2> gpurun_out/solve_time.err || tail -5 gpurun_out/solve_time.err
cat gpurun_out/solve_time.json
