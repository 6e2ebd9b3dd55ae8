#!/bin/bash
# 1/2/4/8-GPU runs on ONE 8-GPU box: the copy-only ceiling and bench.py (C3 strong scaling + soup legs).
# usage (under gpurun --gpus 8): bash profiles/run_scale.sh [steps]
STEPS=${1:-5}
OUT=gpurun_out/scale
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topology.txt 2>&1
lscpu | grep -i "numa\|^CPU(s)\|model name\|socket" > $OUT/host.txt 2>&1
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N))"; fi
  $L profiles/copy_only.py 2>$OUT/copy_n$N.err | grep '^{' > $OUT/copy_n$N.json
  EXTRA="--skip-soup10m"
  if [ $N -eq 2 ] || [ $N -eq 4 ]; then EXTRA="--skip-soup10m --skip-c5"; fi
  $L bench.py --gpus $N --steps $STEPS --warmup 3 $EXTRA 2>$OUT/bench_n$N.err | grep '^{' > $OUT/bench_n$N.json
  echo "N=$N: $(python - <<PY
import json
try:
    c=json.load(open("$OUT/copy_n$N.json")); d=json.load(open("$OUT/bench_n$N.json"))
    print("copy both %.1f GB/s total (%.1f per GPU) | C3 %.1f spp/s %.1f ms allreduce %s | exact %.0f fp32 %.0f e2e %.0f Mrays/s" % (
        c["both"]["GBps_total"], c["both"]["GBps_per_gpu"], d["value"], d["ms_per_step"], d["allreduce"] and round(d["allreduce"]["ms_per_frame"],3),
        d["closest_hit"]["exact"]["value"], d["closest_hit"]["fp32"]["value"], d["closest_hit"]["e2e"]["value"]))
except Exception as e:
    print("parse error", e)
PY
)"
done
