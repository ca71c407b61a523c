"""Dev tool: function-level groups (name, file, first line, last line) for tools/linemap_groups.py, derived from the
source files themselves so that they follow edits.  usage: python tools/auto_groups.py > groups.py"""
import re, os
CS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vectorizedbayesiannetwork_b200", "csrc")
out = []
for f in ("vbn_schedule_tc.cuh", "vbn_schedule.cuh", "vbn_device.cuh"):
    lines = open(os.path.join(CS, f)).read().splitlines()
    starts = []
    for i, ln in enumerate(lines, 1):
        m = re.match(r"\s*(?:static )?(?:__device__|__global__|__host__)[^;]*?\b([A-Za-z_0-9]+)\s*\(", ln)
        if m and not ln.strip().startswith("//"):
            starts.append((i, m.group(1)))
    for (a, name), (b, _) in zip(starts, starts[1:] + [(len(lines) + 1, "")]):
        out.append((f"{f.split('_')[-1][:3]}:{name}", f, a, b - 1))
out.append(("ldg intrinsic", "sm_32_intrinsics.hpp", 1, 100000))
print(repr(out))
