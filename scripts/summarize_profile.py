#!/usr/bin/env python
"""Turn the raw ncu output a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:
   profiles/rNN_launches.csv        per-kernel launch list of one bench step sequence (gpu__time_duration.sum)
   profiles/rNN_frontend_ncu.txt    key metrics of the dominant kernel from the --set full capture
   profiles/traffic.json            DRAM bytes per k_frontend launch (read by bench.py -> roofline.traffic)
usage: scripts/summarize_profile.py <round tag, e.g. r01>"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

# ---- launch list
lines = [l for l in open(os.path.join(G, "launches.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg.setdefault(row["Kernel Name"], []).append(v)
ours = {k: v for k, v in agg.items() if "csdr::" in k}
tot = sum(sum(v) for v in ours.values())
with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu\n")
    f.write("# (cold-cache, serialised launches: compare SHARES, not absolutes); chunk = 2^26 samples\n")
    f.write("kernel,launches,avg_us,min_us,max_us,share_of_csdr_time\n")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"\"{k}\",{len(v)},{sum(v)/len(v):.1f},{min(v):.1f},{max(v):.1f},{sum(v)/tot:.4f}\n")

# ---- full capture of the dominant kernel
rep = os.path.join(G, "prof_frontend.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, rows = r[0], r[1], r[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
with open(os.path.join(OUT, f"{tag}_frontend_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 3 -c 2 python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"{w:92s} {units[i]:14s} {[row[i] for row in rows]}\n")

def col(name):
    i = hdr.index(name)
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
    return [float(row[i].replace(",", "")) * scale for row in rows]
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
json.dump({"kernel": "k_frontend", "chunk_samples": 1 << 26,
           "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
           "dram_bytes_read": sum(rd) / len(rd), "dram_bytes_write": sum(wr) / len(wr),
           "source": f"profiles/{tag}_frontend_ncu.txt (ncu --set full, 2^26-sample launch)"},
          open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
# back-end kernels (one launch each), same metric list; kept as is when this gpurun call captured none of them
_be = ("k_agc_chain", "k_be_emit", "k_be_prep", "k_dc_local", "k_pfb_tile", "k_pfb_ring")
with open(os.path.join(OUT, f"{tag}_backend_ncu.txt"),
          "w" if any(os.path.exists(os.path.join(G, f"prof_{k}.ncu-rep")) for k in _be) else "a") as f:
  if f.mode == "w":
    f.write("# ncu --set full --clock-control none -k regex:<kernel> -s 4 -c 1 python bench.py --steps 2 --warmup 3 --log2n 26 --no-cpu\n"
            "# (k_pfb_*: python scripts/prof_c3.py 16|1024 agc, 2^24 input samples)\n")
    for k in ("k_agc_chain", "k_be_emit", "k_be_prep", "k_dc_local", "k_pfb_tile", "k_pfb_ring"):
        r = os.path.join(G, f"prof_{k}.ncu-rep")
        if not os.path.exists(r):
            continue
        raw = subprocess.run(["ncu", "-i", r, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        if len(rr) < 3:
            continue
        h2, u2, v2 = rr[0], rr[1], rr[2]
        f.write(f"## {k}\n")
        for w in want:
            if w in h2:
                i = h2.index(w)
                f.write(f"{w:92s} {u2[i]:14s} {v2[i]}\n")
print(open(os.path.join(OUT, f"{tag}_launches.csv")).read())
