"""Dev tool: like tools_linemap.py but aggregates executed instructions / stall samples by source-line ranges."""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1:4]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ii, si, sass_i = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
inst = [(r[sass_i], int(r[ii] or 0), int(r[si] or 0)) for r in rows[2:] if len(r) > si]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; infunc = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kname in ln
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln): lines.append(cur)
groups = eval(open(sys.argv[4]).read())  # list of (name, file, lo, hi)
agg = collections.OrderedDict((g[0], [0, 0, 0]) for g in groups); agg["other"] = [0, 0, 0]
ops = collections.Counter()
for (s, n, smp), loc in zip(inst, lines):
    f, l = loc if loc else ("?", 0)
    name = "other"
    for g in groups:
        if g[1] == f and g[2] <= l <= g[3]: name = g[0]; break
    a = agg[name]; a[0] += n; a[1] += smp; a[2] += 1
    ops[s.split()[0] if not s.strip().startswith("@") else s.split()[1]] += n
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
for k, (n, smp, cnt) in agg.items():
    print(f"{k:28s} {n/tot*100:5.1f}% inst  {smp/max(tots,1)*100:5.1f}% samples  sass={cnt}")
print("top opcodes:", [(k, round(v/tot*100,1)) for k, v in ops.most_common(25)])
