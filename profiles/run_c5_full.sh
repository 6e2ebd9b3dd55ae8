#!/bin/bash
# BASELINE configs[4] as written: subdivided Cornell (4.7M triangles, dielectric + conductor), 3840x2160, 4096 spp, on 8 B200
# -- one frame of 4096 spp (512 per GPU) through bench.py's c5 leg (1 warm-up frame + 3 timed).  Also the multi-GPU test.
# usage (under gpurun --gpus 8): bash profiles/run_c5_full.sh
mkdir -p gpurun_out/scale
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | grep -o "\[multi-gpu\].*\|passed.*\|failed.*"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 3 --warmup 3 \
  --skip-soup --skip-soup10m --skip-cpu --c5-spp 4096 2>gpurun_out/scale/c5_4096_n8.err | grep '^{' > gpurun_out/scale/c5_4096_n8.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/scale/c5_4096_n8.json"))
c=d["c5"]
print("C3 N=8:", round(d["value"],1), "spp/s; C5:", c["spp_per_frame"], "spp per frame in", round(c["ms_per_frame"]/1e3,3), "s =", round(c["spp_per_s"],1), "spp/s,", round(c["mrays_per_s"]), "Mrays/s")
PY
