"""Dev / evidence tool (GPU): KDE log-density at high dimension, tensor-core kernel vs the FP32-pipe generic kernel.
usage: python tools/bench_kde_tc.py [n_points] [rows] [dp] [dx]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectorizedbayesiannetwork_b200 as V

n, rows, dp, dx = [int(a) for a in (sys.argv[1:5] + ["50000", "262144", "7", "1"][len(sys.argv) - 1:])]
g = torch.Generator().manual_seed(0)
p = torch.randn(n, dp, generator=g)
y = torch.sin(p.sum(1, keepdim=True)) + 0.4 * torch.randn(n, dx, generator=g)
c = {"kind": "kde", "input_dim": dp, "output_dim": dx, "bandwidth": 0.5, "parent_bandwidth": 0.6, "min_scale": 1e-4,
     "parents": p, "targets": y}
dev = torch.device("cuda", 0)
cpd = V.cpd_from_spec(c, device=dev)
x, q = torch.randn(rows, dx, generator=g).to(dev), torch.randn(rows, dp, generator=g).to(dev)
for mode in ("1", "0"):
    os.environ["VBN_KDE_TC"] = mode
    r = rows if mode == "1" else min(rows, 16384)
    for _ in range(2):
        out = cpd.log_prob(x[:r], q[:r])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = cpd.log_prob(x[:r], q[:r])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    pairs = r * n / (ms * 1e-3)
    frac_exact = float(torch.isnan(out).float().mean())
    print(f"{'tcgen05' if mode == '1' else 'fp32-pipe generic'}: rows {r} x points {n} (dp {dp}, dx {dx}): {ms:.2f} ms, "
          f"{pairs:.3e} pairs/s, {2 * pairs / (16 * 148 * 1.965e9):.2%} of the MUFU ceiling (2 ex2/pair)", flush=True)
