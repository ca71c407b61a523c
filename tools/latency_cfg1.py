"""cfg1 (README minimal example: gaussian_nn x2 -> mdn K=3, monte_carlo_marginalization, 1 query x 200 samples):
per-call latency of the public API on the GPU, next to the oracle port on the host CPU."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vectorizedbayesiannetwork_b200 as V  # noqa: E402
from oracle import vbn_oracle as O  # noqa: E402

blob = torch.load(os.path.join(ROOT, "tests", "golden", "readme.pt"), weights_only=False)
spec = blob["spec"]
q = {"target": "feature_2", "evidence": {"feature_0": torch.tensor([[0.3]]), "feature_1": torch.tensor([[-0.2]])}}
model = V.VBN.from_spec(spec, device="cuda")
model.set_inference_method("monte_carlo_marginalization", n_samples=200)
qd = {"target": "feature_2", "evidence": {k: v.cuda() for k, v in q["evidence"].items()}}
for _ in range(20):
    model.infer_posterior(qd)
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    pdf, s = model.infer_posterior(qd)
torch.cuda.synchronize()
gpu = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for _ in range(n):
    pdf, s = model.infer_posterior(qd)
    pdf.cpu()
sync = (time.perf_counter() - t0) / n
qo = {"target": "feature_2", "evidence": q["evidence"], "do": {}}
for _ in range(5):
    O.monte_carlo_marginalization(spec, qo, 200)
t0 = time.perf_counter()
for _ in range(100):
    O.monte_carlo_marginalization(spec, qo, 200)
cpu = (time.perf_counter() - t0) / 100
print(f"cfg1 MCM 1x200: GPU {gpu * 1e6:.0f} us/call pipelined, {sync * 1e6:.0f} us/call with result on host; "
      f"oracle port on CPU {cpu * 1e6:.0f} us/call")
