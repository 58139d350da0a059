"""Summarise an .ncu-rep of one kernel: headline metrics + instructions / shared wavefronts / stall samples per
barrier-delimited code region (usage: ncu_regions.py file.ncu-rep)."""
import csv, sys, subprocess, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
for h, v in zip(hdr, vals):
    if h in want: print(f"{h:90s} {v}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
ends = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
end = ends[1] if len(ends) > 1 else len(lines)
rd = csv.DictReader(io.StringIO("\n".join(lines[start:end])))
seg = 0; segs = collections.defaultdict(collections.Counter); stalls = collections.defaultdict(collections.Counter)
for r in rd:
    toks = r["Source"].strip().split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    if not op.startswith(('LDS', 'STS', 'LDG', 'STG', 'BAR', 'MUFU', 'LDGSTS')): op = op.split('.')[0]
    n = int(r["Instructions Executed"])
    segs[seg]['inst'] += n; segs[seg]['samples'] += int(r["# Samples"]); segs[seg]['wf'] += int(r["L1 Wavefronts Shared"] or 0)
    segs[seg][op] += n
    for k, v in r.items():
        if k.startswith('stall_') and 'Not Issued' not in k and v and int(v): stalls[seg][k[6:]] += int(v)
    if op.startswith('BAR'): seg += 1
tot = sum(s['inst'] for s in segs.values()); ts = sum(s['samples'] for s in segs.values())
for k, s in segs.items():
    top = [(o, c) for o, c in s.most_common(14) if o not in ('inst', 'samples', 'wf')]
    print(f"region {k}: inst {100 * s['inst'] / tot:.1f}% ({s['inst'] / 1e6:.1f}M) samples {100 * s['samples'] / max(ts, 1):.1f}% smem-wf {s['wf'] / 1e6:.2f}M")
    print("    ops   ", " ".join(f"{o}:{c / 1e6:.2f}M" for o, c in top[:11]))
    print("    stalls", " ".join(f"{o}:{c}" for o, c in stalls[k].most_common(7)))
