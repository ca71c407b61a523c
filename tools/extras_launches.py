"""Dev tool: one call each of the SURVEY 8f additions (Gibbs chain, adapter summaries, rff / embedded CPDs), so that an
`ncu --metrics gpu__time_duration.sum` launch list shows which native kernels serve them."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vectorizedbayesiannetwork_b200 as V  # noqa: E402
from vectorizedbayesiannetwork_b200 import summaries as SM  # noqa: E402
from vectorizedbayesiannetwork_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
# Gibbs: 4096 chains x (10 + 200) sweeps on a 12-node LG/MDN DAG -- one launch
model = V.VBN.from_spec(S.random_dag_lg_mdn(12, seed=3), device=dev)
model.set_sampling_method("gibbs", n_samples=200, burn_in=10)
ev = {"n11": torch.randn(4096, 1, generator=g)}
for _ in range(3):
    s = model.sample({"target": "n5", "evidence": ev}, n_samples=200)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s = model.sample({"target": "n5", "evidence": ev}, n_samples=200)
e1.record()
torch.cuda.synchronize()
print(f"gibbs: 4096 chains x 210 sweeps x 11 latent nodes x 8 candidates in {e0.elapsed_time(e1):.2f} ms (one launch)")
# ALARM posterior + class histogram on the device
alarm = V.VBN.from_spec(S.alarm_softmax(seed=0), device=dev)
alarm.set_inference_method("likelihood_weighting", n_samples=16384)
evd = {n: torch.randint(0, S.ALARM[n][0], (4096, 1), generator=g).float() for n in ("HRBP", "BP", "EXPCO2", "PRESS")}
w, x = alarm.infer_posterior({"target": "LVFAILURE", "evidence": evd})
for _ in range(3):
    p = SM.estimate_discrete_posterior_tensor(x, w, 2)
torch.cuda.synchronize()
e0.record()
p = SM.estimate_discrete_posterior_tensor(x, w, 2)
e1.record()
torch.cuda.synchronize()
print(f"weighted histogram: 4096 queries x 16384 samples -> [4096, 2] in {e0.elapsed_time(e1):.3f} ms "
      f"({(2 * 4 * 4096 * 16384) / e0.elapsed_time(e1) / 1e6:.0f} GB/s of the two [B,S] inputs); rows sum to {p.sum(1).mean().item():.6f}")
for name in ("rff", "embedded"):
    spec = torch.load(os.path.join(ROOT, "tests", "golden", f"{name}.pt"), weights_only=False)["spec"]
    m = V.VBN.from_spec(spec, device=dev)
    m.set_inference_method("likelihood_weighting", n_samples=4096)
    t = spec["nodes"][-1]
    first = spec["topo"][0]
    val = torch.zeros(256, spec["cpds"][first]["output_dim"])
    w, x = m.infer_posterior({"target": t, "evidence": {first: val}})
    torch.cuda.synchronize()
    print(f"{name}: LW 256 x 4096 ok, weights sum {w.sum(1).mean().item():.5f}")
