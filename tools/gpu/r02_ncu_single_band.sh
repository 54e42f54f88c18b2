# ncu --set full (source-level stall sampling) of one band of the chained kernel: b=4 (2048 x 8) and b=9 (2048 x 3)
mkdir -p gpurun_out
DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py 2048 8 1 3 stream:gs_fwd > gpurun_out/r02_plain_sb4.json 2>&1 &&
DGB_GS_VARIANT=22 timeout 600 ncu --set full --warp-sampling-interval 0 --import-source on --clock-control none -k regex:k_gs_chain -s 1 -c 1 -o gpurun_out/r02_chain_b4_single -f python tools/probe_kernels.py 2048 8 1 3 stream:gs_fwd > gpurun_out/r02_ncu_sb4.log 2>&1
tail -2 gpurun_out/r02_ncu_sb4.log
DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py 2048 3 2 3 stream:gs_fwd > gpurun_out/r02_plain_sb9.json 2>&1 &&
DGB_GS_VARIANT=22 timeout 600 ncu --set full --warp-sampling-interval 0 --import-source on --clock-control none -k regex:k_gs_chain -s 1 -c 1 -o gpurun_out/r02_chain_b9_single -f python tools/probe_kernels.py 2048 3 2 3 stream:gs_fwd > gpurun_out/r02_ncu_sb9.log 2>&1
tail -2 gpurun_out/r02_ncu_sb9.log
