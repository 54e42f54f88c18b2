# round 2 baseline: ncu --set full of the chained GS kernels on the full 2048^2 grid (b=9 and b=4) and of the
# fused entry-residual helper; each after the same command exited 0 without ncu
mkdir -p gpurun_out
for p in 2 1; do
  timeout 300 python tools/probe_kernels.py 2048 2048 $p 3 stream:gs_fwd > gpurun_out/r02_plain_chain_p$p.json 2>gpurun_out/r02_plain_chain_p$p.err &&
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gs_chain -s 1 -c 2 -o gpurun_out/r02_chain_full_p$p -f python tools/probe_kernels.py 2048 2048 $p 3 stream:gs_fwd > gpurun_out/r02_ncu_chain_p$p.log 2>&1
  tail -2 gpurun_out/r02_ncu_chain_p$p.log
done
timeout 300 python tools/probe_kernels.py 2048 2048 2 3 stream:entry_residual > gpurun_out/r02_plain_helper.json 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gs_helper -s 1 -c 1 -o gpurun_out/r02_helper_res_b9 -f python tools/probe_kernels.py 2048 2048 2 3 stream:entry_residual > gpurun_out/r02_ncu_helper.log 2>&1
tail -2 gpurun_out/r02_ncu_helper.log
cat gpurun_out/r02_plain_chain_p2.json gpurun_out/r02_plain_chain_p1.json gpurun_out/r02_plain_helper.json
ls -la gpurun_out/*.ncu-rep
