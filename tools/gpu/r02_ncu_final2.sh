# round 2, final state: ncu --set full of the kernels that changed after r02_ncu_final.sh (chain b=9 with clusters of 9, chain b=4
# with the producer warp, record residual with TMA tiles), each after the same command exited 0 without ncu; then the launch list
mkdir -p gpurun_out
run() {  # name, p, probe target, kernel regex
  timeout 300 python tools/probe_kernels.py 2048 2048 $2 3 stream:$3 > gpurun_out/r02_plain_$1.json 2>gpurun_out/r02_plain_$1.err &&
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$4 -s 1 -c 1 -o gpurun_out/r02_full_$1 -f python tools/probe_kernels.py 2048 2048 $2 3 stream:$3 > gpurun_out/r02_ncu_$1.log 2>&1
  tail -1 gpurun_out/r02_ncu_$1.log
}
run chain_b9 2 gs_fwd k_gs_chain
run chain_b4 1 gs_fwd k_gs_chain
run resrec_b9 2 rec_residual k_residual_rec
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/r02_full_*.ncu-rep
