"""Turns the raw ncu outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under
profiles/: per-workload launch shares, key `--set full` metrics of the dominant kernel, SASS evidence,
and profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


traffic = {}
for w in ("cfg5", "cfg2", "cfg3", "cfg4"):
    lines = []
    # ---- launch list -------------------------------------------------------------------------
    lp = os.path.join(GP, f"{R}_launches_{w}.csv")
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
        hdr = next((r for r in rows if "Kernel Name" in r), None)
        if hdr:
            ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
            agg = collections.OrderedDict()
            for r in rows[rows.index(hdr) + 1:]:
                try:
                    t = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(r[ui], 1e-6)
                except ValueError:
                    continue
                a = agg.setdefault(r[ki].split("(")[0][:90], [0, 0.0])
                a[0] += 1
                a[1] += t
            tot = sum(a[1] for a in agg.values())
            lines.append(f"# launch list ({os.path.basename(lp)}): whole bench process incl. warm-up, probes and e2e loop; "
                         "cold-cache serialised times -> compare SHARES")
            for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                lines.append(f"{t / tot * 100:6.2f}%  {t:10.3f} ms  x{n:<4d} {k}")
    # ---- full capture ------------------------------------------------------------------------
    rp = os.path.join(GP, f"{R}_prof_{w}.raw.csv")  # `ncu -i <rep> --page raw --csv`, exported on the GPU box
    if os.path.exists(rp):
        raw = open(rp).read()
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) > 2:
            hdr, units, vals = rows[0], rows[1], rows[2]
            lines.append("")
            lines.append(f"# ncu --set full, one launch of the dominant kernel ({os.path.basename(rp)})")
            lines.append("Kernel Name = " + vals[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    lines.append(f"{k} = {vals[hdr.index(k)]} {units[hdr.index(k)]}")
            for i, k in enumerate(hdr):
                if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                    lines.append(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', 'stall_')} = {vals[i]}")
            try:
                plain = json.loads([l for l in open(os.path.join(GP, f"{R}_plain_{w}.json")) if l.startswith("{")][-1])
                rows_launch = plain["roofline"].get("rows_per_launch") or plain["config"].get("rows_per_gpu")
                db = to_bytes(vals[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) + \
                    to_bytes(vals[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
                traffic[w] = {"rows": int(rows_launch), "dram_bytes": db, "file": f"profiles/{R}_ncu_{w}.txt"}
                ia = "smsp__issue_active.avg.pct_of_peak_sustained_active"
                if ia in hdr:
                    traffic[w]["issue_active_pct"] = round(float(vals[hdr.index(ia)]), 2)
                    traffic[w]["inst_per_row"] = round(
                        float(vals[hdr.index("smsp__inst_executed.sum")].replace(",", "")) * 32 / int(rows_launch), 1)
                lines.append(f"rows in the captured launch = {rows_launch}; dram bytes read+write = {db:.0f}")
            except Exception as exc:  # noqa: BLE001
                lines.append(f"(traffic not derived: {exc})")
    if lines:
        open(os.path.join(OUT, f"{R}_ncu_{w}.txt"), "w").write("\n".join(lines) + "\n")
# ---- SASS evidence ---------------------------------------------------------------------------
so = os.path.join(ROOT, "vectorizedbayesiannetwork_b200", "libvbn_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cnt = collections.Counter()
fn = None
for ln in sass.splitlines():
    if "Function :" in ln:
        fn = ln.split("Function :")[1].strip()
    for tag in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "HMMA"):
        if f" {tag}" in ln or f"\t{tag}" in ln:
            cnt[(fn, tag)] += 1
ev = ["# Blackwell-native SASS mnemonics per kernel (cuobjdump -sass libvbn_cuda.so); see B200_PROFILING.md table"]
for (f, tag), n in sorted(cnt.items()):
    ev.append(f"{n:6d}  {tag:10s} {f[:100]}")
open(os.path.join(OUT, f"{R}_sass_evidence.txt"), "w").write("\n".join(ev) + "\n")
if traffic:
    json.dump(traffic, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(OUT)))
