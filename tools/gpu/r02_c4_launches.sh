mkdir -p gpurun_out
timeout 600 python bench.py --config c4 --size 512 > gpurun_out/plain_c4.json 2>gpurun_out/plain_c4.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv python bench.py --config c4 --size 512 > gpurun_out/ncu_c4.log 2>&1; echo "rc=$?"
