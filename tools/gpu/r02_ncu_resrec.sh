mkdir -p gpurun_out
timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_residual_rec<.int.9" -s 1 -c 1 -o gpurun_out/r02_resrec_b9 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/ncu_resrec.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_resrec.log
