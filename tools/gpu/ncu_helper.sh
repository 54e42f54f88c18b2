mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gs_helper -c 2 -o gpurun_out/helper_res_b9 -f python tools/probe_kernels.py 2048 2048 2 1 stream:entry_residual > gpurun_out/ncu_helper.log 2>&1; tail -2 gpurun_out/ncu_helper.log
