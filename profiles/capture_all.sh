#!/bin/bash
# All ncu --set full captures the bench's roofline lines quote (run under gpurun, ONE GPU).  Each capture is
# preceded by a plain run of the same command.  Reports land in gpurun_out/; turn them into profiles/ncu_*.json
# on the CPU box with profiles/extract_all.sh.
set -u
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python profiles/prof_trace.py 3 exact 1000000 > $O/plain_soup1m_exact.log 2>&1 && $NCU -k regex:trace_persistent -s 2 -c 1 -o $O/prof_soup1m_exact python profiles/prof_trace.py 3 exact 1000000 > $O/ncu_soup1m_exact.log 2>&1
python profiles/prof_trace.py 3 fp32 1000000 > $O/plain_soup1m_fp32.log 2>&1 && $NCU -k regex:trace_persistent -s 2 -c 1 -o $O/prof_soup1m_fp32 python profiles/prof_trace.py 3 fp32 1000000 > $O/ncu_soup1m_fp32.log 2>&1
python profiles/prof_trace.py 3 exact 10000000 > $O/plain_soup10m_exact.log 2>&1 && $NCU -k regex:trace_persistent -s 2 -c 1 -o $O/prof_soup10m_exact python profiles/prof_trace.py 3 exact 10000000 > $O/ncu_soup10m_exact.log 2>&1
python profiles/prof_trace.py 3 fp32 10000000 > $O/plain_soup10m_fp32.log 2>&1 && $NCU -k regex:trace_persistent -s 2 -c 1 -o $O/prof_soup10m_fp32 python profiles/prof_trace.py 3 fp32 10000000 > $O/ncu_soup10m_fp32.log 2>&1
# Cornell: 8 depth-probing renders (1+2+..+8 = 36 bounces: 28 closest_kernel launches, 36 shade / shadow launches), then the
# timed waves; capture bounce 1 of the first timed wave (closest_kernel #29, shade_kernel #38 = its bounce 1, shadow_kernel #38)
python profiles/prof_render.py 2 > $O/plain_render.log 2>&1 && $NCU -k regex:closest_kernel -s 28 -c 1 -o $O/prof_cornell_closest python profiles/prof_render.py 2 > $O/ncu_cornell_closest.log 2>&1
$NCU -k regex:shade_kernel -s 37 -c 1 -o $O/prof_cornell_shade python profiles/prof_render.py 2 > $O/ncu_cornell_shade.log 2>&1
$NCU -k regex:shadow_kernel -s 37 -c 1 -o $O/prof_cornell_shadow python profiles/prof_render.py 2 > $O/ncu_cornell_shadow.log 2>&1
# launch list of a short bench run (shares of the step per kernel)
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-soup10m --skip-c5 --total-spp 64 > $O/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-soup10m --skip-c5 --total-spp 64 > $O/ncu_bench.log 2>&1
ls -la $O/*.ncu-rep; tail -n 3 $O/plain_render.log
