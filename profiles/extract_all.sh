#!/bin/bash
# gpurun_out/prof_*.ncu-rep (profiles/capture_all.sh) -> profiles/ncu_*.json + the launch-list summary (CPU box)
set -e
R=gpurun_out
RAYS=$(python -c "import json;d=json.loads([l for l in open('$R/plain_render.log') if l.startswith('{')][0]);print(d['rays_per_bounce_closest'][1], d['rays_per_bounce_shadow'][1])")
set -- $RAYS
N=16777216
python profiles/ncu_extract.py $R/prof_soup1m_exact.ncu-rep profiles/ncu_soup1m_exact.json trace_persistent "soup-1M, one 2^24-ray PRT_TRACE_EXACT launch (profiles/prof_trace.py 3 exact)" rays_in_launch=$N > /dev/null
python profiles/ncu_extract.py $R/prof_soup1m_fp32.ncu-rep profiles/ncu_soup1m_fp32.json trace_persistent "soup-1M, one 2^24-ray plain launch (profiles/prof_trace.py 3 fp32)" rays_in_launch=$N > /dev/null
python profiles/ncu_extract.py $R/prof_soup10m_exact.ncu-rep profiles/ncu_soup10m_exact.json trace_persistent "soup-10M, one 2^24-ray PRT_TRACE_EXACT launch" rays_in_launch=$N > /dev/null
python profiles/ncu_extract.py $R/prof_soup10m_fp32.ncu-rep profiles/ncu_soup10m_fp32.json trace_persistent "soup-10M, one 2^24-ray plain launch" rays_in_launch=$N > /dev/null
python profiles/ncu_extract.py $R/prof_cornell_closest.ncu-rep profiles/ncu_cornell_closest.json closest_kernel "Cornell 1024^2, 16-spp wave, bounce 1 (profiles/prof_render.py)" rays_in_launch=$1 > /dev/null
python profiles/ncu_extract.py $R/prof_cornell_shade.ncu-rep profiles/ncu_cornell_shade.json shade_kernel "Cornell 1024^2, 16-spp wave, bounce 1: paths shaded = closest rays of bounce 1" rays_in_launch=$1 > /dev/null
python profiles/ncu_extract.py $R/prof_cornell_shadow.ncu-rep profiles/ncu_cornell_shadow.json shadow_kernel "Cornell 1024^2, 16-spp wave, bounce 1 shadow rays" rays_in_launch=$2 > /dev/null
for f in profiles/ncu_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(f"{sys.argv[1]:38s} {d['duration_ns']/1e6:7.3f} ms  dram {d['dram_bytes']/1e9:6.3f} GB ({d.get('dram_throughput_pct',0):4.1f} %)  L1 pipe {d.get('l1_data_pipe_pct',0):4.1f} %  L2 {d.get('l2_throughput_pct',0):4.1f} %  issue {d.get('issue_active_pct',0):4.1f} %  lanes {d.get('lanes_per_instruction',0):4.1f}  occ {d.get('occupancy_pct',0):4.1f} %  regs {d.get('registers',0):.0f}  L2 hit {d.get('l2_hit_pct',0):4.1f} %")
PY
done
# the --page details text of the same captures
for p in soup1m_exact soup1m_fp32 soup10m_fp32 cornell_closest cornell_shade; do
  ncu -i $R/prof_$p.ncu-rep --page details 2>/dev/null | grep -v -e '^ *-\{5,\}' -e '^ *$' > profiles/r2_${p}_ncu.txt
done
cp $R/bench_launches.csv profiles/r2_bench_launches.csv
python profiles/summarize_launches.py $R/bench_launches.csv > profiles/r2_bench_launches.txt
