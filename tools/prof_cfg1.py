import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vectorizedbayesiannetwork_b200 as V
blob = torch.load("tests/golden/readme.pt", weights_only=False)
model = V.VBN.from_spec(blob["spec"], device="cuda")
model.set_inference_method("monte_carlo_marginalization", n_samples=200)
q = {"target": "feature_2", "evidence": {"feature_0": torch.tensor([[0.3]]).cuda(), "feature_1": torch.tensor([[-0.2]]).cuda()}}
for _ in range(50): model.infer_posterior(q)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): model.infer_posterior(q)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
