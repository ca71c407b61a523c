"""Dev tool: top source lines for one stall reason column of an `ncu --page source --csv` dump."""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname, col = sys.argv[1:5]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci, sass_i = hdr.index(col), hdr.index("Source")
inst = [(r[sass_i], int(r[ci] or 0)) for r in rows[2:] if len(r) > ci]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; infunc = False
for ln in dis:
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln): infunc = kname in ln
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln): lines.append(cur)
agg = collections.Counter(); ex = {}
for (s, n), loc in zip(inst, lines):
    agg[loc] += n
    if n > ex.get(loc, ("", 0))[1]: ex[loc] = (s.strip()[:60], n)
tot = sum(agg.values())
for loc, n in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 15):
    print(f"{n/max(tot,1)*100:5.1f}%  {loc}  e.g. {ex[loc][0]}")
