#!/usr/bin/env python
"""Turns the raw ncu outputs of a GPU session (gpurun_out/<tag>/) into the small, tracked
summaries under profiles/:  <tag>_launches.csv (per-kernel totals of the launch list),
<tag>_<name>_kernel.json (key metrics of the --set full capture) and, for the headline kernel,
profiles/headline_kernel_traffic.json which bench.py reads for roofline.traffic.

    python scripts/summarize_ncu.py r01b prof_c2c c2c [--headline]
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg",
]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", tag, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r["Metric Value"])
        except Exception:
            continue
        name = r["Kernel Name"]
        short = name.split("(")[0][-110:] if len(name) > 140 else name
        agg[short][0] += 1
        agg[short][1] += v
    total = sum(v[1] for v in agg.values())
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 3 --warmup 3`\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_ms,avg_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t / 1e6:.3f},{t / n / 1e3:.1f},{100 * t / total:.2f}\n")
    print("wrote", out)


def kernel(tag, rep, name, headline):
    path = os.path.join(ROOT, "gpurun_out", tag, rep + ".ncu-rep")
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = {"value": float(r[i]), "unit": units[i]}
                except Exception:
                    pass
        stalls = {}
        for i, k in enumerate(hdr):
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    stalls[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(r[i])
                except Exception:
                    pass
        tot = sum(stalls.values()) or 1.0
        d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        res.append(d)
    out = os.path.join(ROOT, "profiles", f"{tag}_{name}_kernel.json")
    json.dump({"command": f"ncu --set full --clock-control none --import-source on ({rep}.ncu-rep)", "launches": res},
              open(out, "w"), indent=1)
    print("wrote", out)
    if headline and res:
        d = res[-1]
        gb = lambda k: d[k]["value"] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d[k]["unit"]]
        traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        sys.path.insert(0, ROOT)
        import bench  # the same hash bench.py recomputes at run time: a capture of other SASS is refused as stale

        json.dump({"kernel": d["kernel"], "dram_bytes_per_launch": traffic, "source": os.path.basename(out),
                   "sass_sha256": bench.headline_sass_hash()},
                  open(os.path.join(ROOT, "profiles", "headline_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    if len(sys.argv) > 3:
        kernel(tag, sys.argv[2], sys.argv[3], "--headline" in sys.argv)
