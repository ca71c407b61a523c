"""Dev tool: join an `ncu --page source --csv` SASS dump with nvdisasm line info and print the
hottest source lines.  usage: python tools_linemap.py <src.csv> <cubin> <kernel-substring>"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1:4]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ai, ii, si = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass_i = hdr.index("Source")
inst = [(r[sass_i], int(r[ii] or 0), int(r[si] or 0)) for r in rows[2:] if len(r) > si]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
lines = []; cur = None; infunc = False; inl = ""
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kname in ln
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln):
        lines.append(cur)
print("ncu instrs", len(inst), "nvdisasm instrs", len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (s, n, smp), loc in zip(inst, lines):
    a = agg[loc]; a[0] += n; a[1] += smp; a[2] += 1
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
srcs = {}
for loc, (n, smp, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    f, l = loc if loc else ("?", 0)
    if f not in srcs:
        try: srcs[f] = open(f"/root/repo/vectorizedbayesiannetwork_b200/csrc/{f}").read().splitlines()
        except Exception: srcs[f] = []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print(f"{n/tot*100:5.1f}% inst {smp/max(tots,1)*100:5.1f}% stall-smp  sass={cnt:4d}  {f}:{l}  {text}")
