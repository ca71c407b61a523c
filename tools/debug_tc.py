"""Dev tool: determinism / batch-independence probes of the tensor-core schedule kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vectorizedbayesiannetwork_b200 as V
from vectorizedbayesiannetwork_b200 import synthetic as S

dev = torch.device("cuda", 0)
spec = S.random_dag_lg_mdn(12, seed=3)
ev = torch.tensor([[0.2], [-0.4]])
for target in ("n0", "n1", "n2", "n3", "n5", "n9"):
    model = V.VBN.from_spec(spec, device=dev)
    model.set_inference_method("importance_sampling", n_samples=64)
    model._inference.ess_threshold = 0.0
    outs = []
    for e in (ev, ev, ev[:1], ev[:1]):
        w, s = model.infer_posterior({"target": target, "evidence": {"n11": e}}, seed=11)
        outs.append((w.cpu(), s.cpu()))
    d_rep2 = (outs[0][1] - outs[1][1]).abs().max().item()
    d_rep1 = (outs[2][1] - outs[3][1]).abs().max().item()
    d_b = (outs[0][1][:1] - outs[2][1]).abs().max().item()
    print(f"target {target}: repeat(B=2) {d_rep2:.3e} repeat(B=1) {d_rep1:.3e} B2[:1] vs B1 {d_b:.3e}",
          "s2[0,:3]", outs[0][1][0, :3, 0].tolist(), "s1[0,:3]", outs[2][1][0, :3, 0].tolist())
