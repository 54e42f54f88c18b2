mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python tools/probe_kernels.py 2048 2048 2 2 stream:residual > gpurun_out/probe_res.json 2>&1; tail -c 300 gpurun_out/probe_res.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_rows -c 3 -o gpurun_out/k_rows_b9_residual -f python tools/probe_kernels.py 2048 2048 2 1 stream:residual > gpurun_out/ncu_res.log 2>&1; tail -2 gpurun_out/ncu_res.log
