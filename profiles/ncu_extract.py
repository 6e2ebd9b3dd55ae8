"""ncu report -> small JSON of the counters bench.py and DESIGN.md quote (run on the CPU box).

  python profiles/ncu_extract.py gpurun_out/prof.ncu-rep profiles/ncu_<name>.json [kernel-substring] [note] [key=number ...]

Writes, per captured launch of the first kernel whose name contains the substring: duration, DRAM bytes
(read + write = roofline.traffic), L1 data-pipe / L2 / issue utilisation, SIMT lanes per instruction,
occupancy, registers -- plus the fingerprint of the CUDA sources the capture belongs to."""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyrenderer_b200.kernel_fingerprint import fingerprint  # noqa: E402

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sectors_srcunit_tex.sum": "l2_sectors_from_sm",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "registers",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_active.min": "sm_cycles_active_min",
    "sm__cycles_active.max": "sm_cycles_active_max",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warp_latency_per_inst_issued.ratio": "warp_cycles_per_instruction",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}  # ncu scales units per value
    launches = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub not in name:
            continue
        d = {"kernel": name, "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for h, u, v in zip(hdr, units, r):
            if h in WANT and v not in ("", "n/a"):
                d[WANT[h]] = float(v.replace(",", "")) * scale.get(u, 1.0)
        launches.append(d)
    if not launches:
        raise SystemExit(f"no kernel matching {sub!r} in {rep}")
    first = launches[0]
    rec = dict(first)
    rec["dram_bytes"] = first.get("dram_bytes_read", 0.0) + first.get("dram_bytes_write", 0.0)
    rec["captured_launches"] = len(launches)
    rec["source_fingerprint"] = fingerprint()
    rec["report"] = os.path.basename(rep)
    rec["note"] = note
    for kv in sys.argv[5:]:
        k, v = kv.split("=", 1)
        rec[k] = float(v)
    json.dump(rec, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
