# usage: probe_once.sh "<NI NJ P>" <call> [env...]   -- one probe_kernels.py timing
mkdir -p gpurun_out
for c in entry_residual residual gs_fwd; do timeout 300 python tools/probe_kernels.py 2048 2048 2 5 stream:$c 2>&1 | tail -1 | cut -c150-400; done
DGB_GS_VARIANT=21 timeout 300 python tools/probe_kernels.py 2048 2048 2 5 stream:gs_fwd 2>&1 | tail -1 | cut -c150-400
