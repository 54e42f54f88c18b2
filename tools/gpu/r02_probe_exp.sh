# round 2: chain kernel alone (DGB_GS_VARIANT=22), experiment builds libdgb200_exp*.so (-DDGB_CHAIN_* tuning macros)
# usage: bash tools/gpu/r02_probe_exp.sh "<libs whose chain tests run first>"
mkdir -p gpurun_out
rm -f gpurun_out/probe_exp.jsonl gpurun_out/probe_exp.err
for e in $(ls dg_multigrid_solver_b200/libdgb200_exp*.so); do
case " $1 " in *" $(basename $e) "*)
DGB_LIB=$PWD/$e timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "chained or streaming" 2>&1 | tail -1;;
esac
for cfg in "2048 2048 2" "2048 2048 1" "1024 1024 1" "512 512 1"; do
  DGB_LIB=$PWD/$e DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe_exp.jsonl 2>>gpurun_out/probe_exp.err || echo "fail $cfg"
done
done
python - <<'PY'
import json, glob
rows=[json.loads(l) for l in open('gpurun_out/probe_exp.jsonl')]
libs=sorted(glob.glob('dg_multigrid_solver_b200/libdgb200_exp*.so'))
k=0
for e in libs:
    print(e.split('_')[-1], [(d['Ni'], d['Nj'], d['b'], d['stream.gs_fwd']['ms'], d['device_error']) for d in rows[k:k+4]]); k+=4
PY
