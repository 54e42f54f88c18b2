# ncu --set full of the operator-assembly kernel: p=2 (2048^2 fine level) and p=5 (512^2 O-grid), after plain runs
mkdir -p gpurun_out
timeout 300 python tools/probe_kernels.py 2048 2048 2 1 stream:apply > /dev/null 2>&1 && echo plain_p2_ok
timeout 300 python tools/bench_configs.py c4 512 > gpurun_out/c4_512.json 2>/dev/null && echo plain_p5_ok
timeout 600 ncu --set full --clock-control none -k regex:k_assemble_poisson -c 1 -o gpurun_out/assemble_p2 -f python tools/probe_kernels.py 2048 2048 2 1 stream:apply > gpurun_out/ncu_asm_p2.log 2>&1; tail -1 gpurun_out/ncu_asm_p2.log
timeout 600 ncu --set full --clock-control none -k regex:k_assemble_poisson -c 1 -o gpurun_out/assemble_p5 -f python tools/bench_configs.py c4 512 > gpurun_out/ncu_asm_p5.log 2>&1; tail -1 gpurun_out/ncu_asm_p5.log
