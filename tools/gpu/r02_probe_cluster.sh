# round 2: chain kernel alone (DGB_GS_VARIANT=22) by cluster size
mkdir -p gpurun_out
rm -f gpurun_out/probe_cluster.jsonl
for cs in 16 12 10 9; do
for cfg in "2048 2048 2" "2048 2048 1" "1024 1024 1" "512 512 1"; do
  DGB_CHAIN_VERBOSE=1 DGB_CHAIN_CLUSTER=$cs DGB_GS_VARIANT=22 timeout 300 python tools/probe_kernels.py $cfg 5 stream:gs_fwd >> gpurun_out/probe_cluster.jsonl 2>>gpurun_out/probe_cluster.err || echo "fail $cfg"
done
done
grep "resident" gpurun_out/probe_cluster.err | sort | uniq
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/probe_cluster.jsonl')]
k=0
for cs in (16,12,10,9):
    print(cs, [(d['Ni'], d['b'], d['stream.gs_fwd']['ms'], d['device_error']) for d in rows[k:k+4]]); k+=4
PY
