"""warp instructions and stall samples of one kernel by CUDA source line: ncu's SASS page joined (in instruction order) with
nvdisasm -g of the same build.  usage: ncu_lines.py file.ncu-rep kernel-substring mangled-substring [lib.so] [top]"""
import csv, sys, subprocess, collections, io, os, re, tempfile
rep, sub, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(__file__), "..", "composable-sdr_b200", "libcsdr_b200.so")
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l][0]
lines_of = []; cur = None
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines_of.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
names = [i for i, l in enumerate(out) if l.startswith('"Kernel Name"')]
k = [n for n, i in enumerate(names) if sub in out[i]][0]
rows = list(csv.DictReader(io.StringIO("\n".join(out[names[k] + 1: names[k + 1] if k + 1 < len(names) else len(out)]))))
assert len(rows) == len(lines_of), (len(rows), len(lines_of))
inst = collections.Counter(); smp = collections.Counter()
for r, ln in zip(rows, lines_of):
    inst[ln] += int(r["Instructions Executed"]); smp[ln] += int(r["# Samples"])
ti, ts = sum(inst.values()), sum(smp.values())
src = {}
print(f"{out[names[k]][:100]}  {ti/1e6:.2f}M warp instructions")
order = smp if os.environ.get("NCU_LINES_BY") == "samples" else inst      # NCU_LINES_BY=samples: rank by stall samples
for ln, _ in order.most_common(top):
    c = inst[ln]
    if ln and ln[0] not in src:
        p = [os.path.join(d, ln[0]) for d in ("composable-sdr_b200/csrc",) if os.path.exists(os.path.join(d, ln[0]))]
        src[ln[0]] = open(p[0]).read().splitlines() if p else []
    text = src[ln[0]][ln[1] - 1].strip()[:110] if ln and src[ln[0]] else ""
    print(f"{100*c/ti:5.1f}% inst {100*smp[ln]/max(ts,1):5.1f}% smp  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5d} {text}")
