"""opcode mix of one kernel in an .ncu-rep (usage: ncu_opmix.py file.ncu-rep kernel-substring [top])"""
import csv, sys, subprocess, collections, io
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
names = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
for n, i in enumerate(names):
    if sub in lines[i]:
        start = i + 1; end = names[n + 1] if n + 1 < len(names) else len(lines); break
else:
    sys.exit("kernel not found")
rd = csv.DictReader(io.StringIO("\n".join(lines[start:end])))
ops = collections.Counter(); stall = collections.Counter(); tot = 0; smp = 0; wf = 0
for r in rd:
    toks = r["Source"].strip().split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.rstrip(';')
    n = int(r["Instructions Executed"]); tot += n; ops[op] += n; s = int(r["# Samples"]); stall[op] += s; smp += s
    wf += int(r["L1 Wavefronts Shared"] or 0)
print(f"{lines[start-1][:120]}\ntotal warp instructions {tot/1e6:.2f}M, samples {smp}, shared wavefronts {wf/1e6:.2f}M")
for op, c in ops.most_common(top):
    print(f"  {op:34s} {c/1e6:8.3f}M {100*c/tot:5.1f}%   samples {100*stall[op]/max(smp,1):5.1f}%")
