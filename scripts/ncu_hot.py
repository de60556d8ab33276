"""Reads an .ncu-rep (--set full --import-source on) and prints where the warp-stall samples are: the hottest
SASS instructions and the samples per block of 200 instructions with their main stall reasons.
   python scripts/ncu_hot.py gpurun_out/<tag>/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
for r in rr[2:3]:
    for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "launch__registers_per_thread"):
        if k in h:
            print(k, "=", r[h.index(k)][:110], rr[1][h.index(k)])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
print("total samples", tot, "instructions", len(data))
for i, r in enumerate(data):
    if "USETMAXREG" in r[ix["Source"]] or "EXIT" in r[ix["Source"]]:
        print("  marker", i, r[ix["Address"]][-5:], r[ix["Source"]][:50])
stalls = [k for k in hdr if k.startswith("stall_") and "Not" not in k]
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][ix["# Samples"]] or 0))[:topn]:
    s = int(r[ix["# Samples"]])
    st = sorted(((k, int(r[ix[k]] or 0)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {r[ix['Address']][-5:]} {100 * s / tot:5.2f}% exec={r[ix['Instructions Executed']]:>9} {r[ix['Source']][:60]:60s} {st}")
b = collections.OrderedDict()
for i, r in enumerate(data):
    e = b.setdefault(i // 200, [0, 0, collections.Counter()])
    e[0] += int(r[ix["# Samples"]] or 0)
    e[1] += int(r[ix["Instructions Executed"]] or 0)
    for k in stalls:
        e[2][k] += int(r[ix[k]] or 0)
for k, (s, n, c) in b.items():
    if s * 100 / tot > 0.3:
        print(f"{k * 200:5d} {100 * s / tot:5.1f}% exec={n / 1e6:7.1f}M", c.most_common(4))
