"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/bench_launches.csv > profiles/r2_bench_launches.txt"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
rows = [row for row in r if len(row) == len(hdr)]
def ms(row):
    v, u = float(row[ix["Metric Value"]]), row[ix["Metric Unit"]]
    return v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
def short(n):
    n = n.replace("(int)", "").replace("(bool)", "").replace("void ", "").replace("prt::", "").replace("<unnamed>::", "")
    return re.sub(r"\(.*", "", n)
agg = collections.OrderedDict()
for row in rows:
    a = agg.setdefault(short(row[ix["Kernel Name"]]), [0, 0.0]); a[0] += 1; a[1] += ms(row)
tot = sum(a[1] for a in agg.values())
print("# launch list of `python bench.py --steps 2 --warmup 3 --skip-cpu --skip-soup10m --skip-c5 --total-spp 64`: ALL launches of the run")
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv  (cold-cache, serialised: compare SHARES, not absolutes)")
print("# headline leg: Cornell 1024^2 frames of 64 spp (= 4 waves of 16 spp): raygen, bounce 0 = trace_persistent_kernel<0,0,1> (EXACT) +")
print("#   finalize + resolve, bounces 1..7 = closest_kernel<0>, per bounce shade_kernel / shadow_kernel / advance_kernel, accumulate per wave;")
print("# soup leg: 5 BVH builds, trace_persistent_kernel<0,0,1> (exact) and <0,0,0> (plain): counted twin + 3 warm-up + 2 timed launches of 2^24 rays each,")
print("#   e2e = 8 chunks of 2^21 rays x (1 warm + 2) per mode")
for name, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{name[:78]:78s} n={n:4d} {t:10.3f} ms {100 * t / tot:5.1f}%")
print(f"total {tot:.3f} ms over {len(rows)} launches")
for tag, label in (("trace_persistent_kernel<0, 0, 1>", "exact"), ("trace_persistent_kernel<0, 0, 0>", "plain")):
    big = sorted(ms(row) for row in rows if short(row[ix["Kernel Name"]]) == tag)
    print(f"soup step ({label}): the 2^24-ray launches under ncu: " + ", ".join(f"{b:.2f}" for b in big[-5:]) + " ms")
fix = sum(agg[k][1] for k in agg if k.startswith(("finalize_kernel", "resolve_kernel")))
ex = agg.get("trace_persistent_kernel<0, 0, 1>", [0, 0.0])[1]
print(f"exact mode: finalize + resolve = {100 * fix / max(ex + fix, 1e-9):.1f} % of (persistent EXACT + fix-up) over the whole run")
render = ("closest_kernel", "shade_kernel", "shadow_kernel", "raygen_kernel", "accumulate_kernel", "advance_kernel")
per = {k: agg[k][1] for k in agg if k.split("<")[0] in render}
st = sum(per.values())
print("render kernels (bounces 1..7 + shade/shadow of every bounce), shares: " + ", ".join(f"{k} {100 * v / st:.0f} %" for k, v in sorted(per.items(), key=lambda x: -x[1])))
