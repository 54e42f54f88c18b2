mkdir -p gpurun_out
DGB_CHAIN_MASK=15 DGB_GS_VARIANT=22 timeout 600 ncu --set full --warp-sampling-interval 0 --import-source on --clock-control none -k regex:k_gs_chain -c 2 -o gpurun_out/chain_b9_single -f python tools/probe_kernels.py 2048 3 2 1 stream:gs_fwd > gpurun_out/ncu4.log 2>&1
tail -5 gpurun_out/ncu4.log
ls -la gpurun_out/*.ncu-rep
