"""Parity of the product path against (a) outputs recorded from the UNMODIFIED reference
(tests/golden/*.pt) and (b) the oracle, with the reference's own draws injected.
Tolerance: 1e-5 relative in fp32 (BASELINE.json north_star), plus a 1e-6 absolute floor for
values near zero.  Runs on the B200 (``gpu``) and, for the schedule logic, in host emulation."""
import glob
import os

import pytest
import torch

from backends import backend  # noqa: F401
from noise_util import log_from_strkeys, to_injection
from oracle import vbn_oracle as O

import vectorizedbayesiannetwork_b200 as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt")))
RTOL, ATOL = 1e-5, 1e-6
METHOD_NAMES = {"lw": "likelihood_weighting", "is": "importance_sampling", "mcm": "monte_carlo_marginalization",
                "gexact": "gaussian_exact", "cexact": "categorical_exact", "ris": "resampled_importance_sampling",
                "rb": "rao_blackwellized_marginalization"}


def _load(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def _close(got, want, what, rtol=RTOL, atol=ATOL):
    got = got.detach().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    both_inf = (torch.isinf(got) & torch.isinf(want) & (torch.sign(got) == torch.sign(want))) | (
        torch.isnan(got) & torch.isnan(want))  # same non-finite value on both sides counts as equal
    g = torch.where(both_inf, torch.zeros_like(got), got)
    w = torch.where(both_inf, torch.zeros_like(want), want)
    err = (g - w).abs()
    bound = atol + rtol * w.abs()
    assert bool((err <= bound).all()), f"{what}: max abs err {err.max().item():.3e}, worst ratio {(err / bound).max().item():.2f}"


def _dev_noise(noise, device):
    return {n: ([t.to(device) for t in d] if isinstance(d, list) else {k: t.to(device) for k, t in d.items()})
            for n, d in noise.items()}


@pytest.mark.parametrize("name", FILES)
def test_inference_methods_match_reference(backend, name):
    blob = _load(name)
    spec = blob["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    assert len(blob["cases"]) > 0
    for case in blob["cases"]:
        q, S, method = case["query"], case["S"], case["method"]
        inj = to_injection(log_from_strkeys(case["noise"]), spec, S)
        tag = f"{name}:{method}:{q['target']}|{sorted(q['evidence'])}|{sorted(q['do'])}"
        if method == "anc":
            model.set_sampling_method("ancestral")
            got = model.sample(q, n_samples=S, noise=_dev_noise(inj.get("anc", {}), backend.device))
            _close(got, case["expect"]["samples"], tag + " samples")
            continue
        model.set_inference_method(METHOD_NAMES[method], n_samples=S)
        if method in ("is", "rb"):  # two passes with their own draws: {"is"|"rb": ..., "lw": fallback}
            noise = {k: _dev_noise(v, backend.device) for k, v in inj.items()}
        else:  # the exact methods draw only through their likelihood-weighting fallback (scope "lw")
            scope = "lw" if method in ("gexact", "cexact") else method
            noise = _dev_noise(inj.get(scope, {}), backend.device)
        pdf, samples = model.infer_posterior(q, noise=noise)
        _close(samples, case["expect"]["samples"], tag + " samples")
        _close(pdf, case["expect"]["pdf"], tag + " pdf", rtol=5e-5)
        if method == "ris":
            assert model._inference._last_resampled == case["info"]["resampled"], tag
        if method == "is":
            assert model._inference._last_fallback == case["info"]["fallback"], tag
            _close(model._inference._last_ess, case["info"]["ess"], tag + " ess", rtol=1e-4)


@pytest.mark.parametrize("name", [f for f in FILES])
def test_cpd_sample_log_prob_match_reference(backend, name):
    blob = _load(name)
    spec = blob["spec"]
    for case in blob["cpd_cases"]:
        node, S = case["node"], case["S"]
        cpd = V.cpd_from_spec(spec["cpds"][node], device=backend.device)
        inj = to_injection(log_from_strkeys(case["noise"]), spec, S)["cpd"][node]
        parents = case["parents"]
        p_dev = None if parents is None else parents.to(backend.device)
        got = cpd.sample(p_dev, S, noise={k: v.to(backend.device) for k, v in inj.items()})
        tag = f"{name}:{node}:{None if parents is None else tuple(parents.shape)}"
        _close(got, case["samples"], tag + " sample")
        x = case["samples"].to(backend.device)
        _close(cpd.log_prob(x, p_dev), case["log_prob"], tag + " log_prob", rtol=2e-5, atol=2e-6)
        p2 = p_dev if p_dev is None or p_dev.dim() == 2 else p_dev[:, :1]
        _close(cpd.log_prob(x[:, 0], p2), case["log_prob_2d"], tag + " log_prob 2d", rtol=2e-5, atol=2e-6)
