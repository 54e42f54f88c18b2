mkdir -p gpurun_out
rm -f gpurun_out/probe12.jsonl
for cfg in "2048 8 1" "2048 16 1" "2048 32 1" "2048 64 1" "2048 256 1" "256 2048 1" "2048 3 2" "2048 12 2" "2048 24 2" "2048 96 2"; do
  DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe12.jsonl 2>gpurun_out/probe12.err || echo "fail $cfg"
done
cut -c1-60,170-400 gpurun_out/probe12.jsonl
