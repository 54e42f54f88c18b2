# round 2, final verification + launch list (after the same command exited 0 without ncu)
bash tools/gpu/r02_final.sh
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 --solve 0 > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --p5-apply 0 --solve 0 > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
