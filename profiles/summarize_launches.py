"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/rNN_launches.csv > profiles/r1_bench_launches.txt"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
rows = [row for row in r if len(row) == len(hdr)]
def ms(row):
    v, u = float(row[ix["Metric Value"]]), row[ix["Metric Unit"]]
    return v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
agg = collections.OrderedDict()
for row in rows:
    name = re.sub(r"\(.*", "", row[ix["Kernel Name"]]).replace("void ", "").replace("prt::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms(row)
tot = sum(a[1] for a in agg.values())
print("# launch list of `python bench.py --steps 2 --warmup 3 --skip-cpu`: ALL launches of the run")
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv  (cold-cache, serialised: compare SHARES, not absolutes)")
print("# soup leg: 5 BVH builds, then trace_persistent_kernel<CLOSEST>: 1 counted twin <0,1>, 3 warm-up + 2 timed launches of 2^24 rays,")
print("#   e2e = 8 chunks x (1 warm + 2) of 2^21 rays; render leg (Cornell 1024^2, 16 spp/step): raygen, then per bounce")
print("#   closest_kernel / shade_kernel / shadow_kernel / advance_kernel, accumulate_kernel per step")
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name[:78]:78s} n={n:4d} {t:10.3f} ms {100 * t / tot:5.1f}%")
print(f"total {tot:.3f} ms over {len(rows)} launches")
big = sorted(ms(row) for row in rows if "trace_persistent_kernel<0, 0>" in row[ix["Kernel Name"]] or "trace_persistent_kernel<(int)0, (bool)0>" in row[ix["Kernel Name"]])
print(f"soup step = ONE launch of trace_persistent_kernel<CLOSEST> over 2^24 rays (kernel share of the timed soup step: 100 %); "
      f"the five 2^24-ray launches under ncu: " + ", ".join(f"{b:.2f}" for b in big[-5:]) + " ms")
per = {k: agg[k][1] for k in agg if k.split("<")[0] in ("closest_kernel", "shade_kernel", "shadow_kernel", "raygen_kernel", "accumulate_kernel", "advance_kernel")}
st = sum(per.values())
print("render step kernel shares: " + ", ".join(f"{k} {100 * v / st:.0f} %" for k, v in sorted(per.items(), key=lambda x: -x[1])))
