# round 2: timeline of the chain kernel (diagnostic build with -DDGB_CHAIN_TRACE)
mkdir -p gpurun_out
rm -f gpurun_out/chain_trace.jsonl
for cfg in "2048 2048 1" "2048 2048 2"; do
  DGB_LIB=$PWD/dg_multigrid_solver_b200/libdgb200_trace.so timeout 300 python tools/chain_trace.py $cfg >> gpurun_out/chain_trace.jsonl 2>gpurun_out/chain_trace.err || { echo "fail $cfg"; tail -5 gpurun_out/chain_trace.err; }
done
cat gpurun_out/chain_trace.jsonl
