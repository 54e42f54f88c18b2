# round 2: timeline of the chain kernel (diagnostic build with -DDGB_CHAIN_TRACE)
# usage: bash tools/gpu/r02_chain_trace.sh ["NI NJ P" ...]
mkdir -p gpurun_out
rm -f gpurun_out/chain_trace.jsonl
if [ $# -eq 0 ]; then set -- "2048 2048 1" "2048 2048 2"; fi
for cfg in "$@"; do
  DGB_LIB=$PWD/dg_multigrid_solver_b200/libdgb200_trace.so timeout 300 python tools/chain_trace.py $cfg >> gpurun_out/chain_trace.jsonl 2>gpurun_out/chain_trace.err || { echo "fail $cfg"; tail -5 gpurun_out/chain_trace.err; }
done
cat gpurun_out/chain_trace.jsonl
