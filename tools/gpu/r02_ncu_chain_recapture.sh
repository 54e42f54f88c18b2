# round 2: ncu --set full of the two chain kernels after the (c, d)-pair stores (DRAM traffic per launch)
mkdir -p gpurun_out
run() {  # name, p, probe target, kernel regex
  timeout 300 python tools/probe_kernels.py 2048 2048 $2 3 stream:$3 > gpurun_out/r02_plain_$1.json 2>gpurun_out/r02_plain_$1.err &&
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$4 -s 1 -c 1 -o gpurun_out/r02_full_$1 -f python tools/probe_kernels.py 2048 2048 $2 3 stream:$3 > gpurun_out/r02_ncu_$1.log 2>&1
  tail -1 gpurun_out/r02_ncu_$1.log
}
run chain_b9 2 gs_fwd k_gs_chain
run chain_b4 1 gs_fwd k_gs_chain
