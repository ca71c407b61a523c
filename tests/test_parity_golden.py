"""Parity of the product path against (a) outputs recorded from the UNMODIFIED reference
(tests/golden/*.pt) and (b) the oracle, with the reference's own draws injected.
Tolerance: 1e-5 relative in fp32 (BASELINE.json north_star), plus a 1e-6 absolute floor for
values near zero.  Runs on the B200 (``gpu``) and, for the schedule logic, in host emulation."""
import glob
import os

import pytest
import torch

from backends import backend  # noqa: F401
from noise_util import gibbs_injection, log_from_strkeys, to_injection
from oracle import vbn_oracle as O

import vectorizedbayesiannetwork_b200 as V

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOT_MODELS = {"summaries", "gibbs"}  # fixtures that are not (model, cases) files
FILES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt"))
               if os.path.basename(p)[:-3] not in NOT_MODELS)
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


def _dot_floor(spec) -> float:
    """Absolute floor for models with rff_gaussian nodes.  Their loc is an fp32 dot product of F terms
    phi*cos(.)*coef_f whose magnitudes sum to T = sqrt(2/F) * sum|coef| * std_y while loc itself is O(1); the
    order of summation alone (BLAS gemv in the reference, sequential FMA here) moves loc by ~eps * T * few.
    The floor is 8 * 2^-24 * T; everything else keeps the 1e-6 floor."""
    floor = 0.0
    for c in spec["cpds"].values():
        if c["kind"] == "rff_gaussian" and c["input_dim"] > 0:
            t = (2.0 / c["n_features"]) ** 0.5 * (c["coef"].abs().sum(0) * c["std_y"].abs()).max().item()
            floor = max(floor, 8 * 2.0**-24 * t)
    return floor


def _dev_noise(noise, device):
    return {n: ([t.to(device) for t in d] if isinstance(d, list) else {k: t.to(device) for k, t in d.items()})
            for n, d in noise.items()}


@pytest.mark.parametrize("name", FILES)
def test_inference_methods_match_reference(backend, name):
    blob = _load(name)
    spec = blob["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    assert len(blob["cases"]) > 0
    atol = max(ATOL, _dot_floor(spec))
    for case in blob["cases"]:
        q, S, method = case["query"], case["S"], case["method"]
        inj = to_injection(log_from_strkeys(case["noise"]), spec, S)
        tag = f"{name}:{method}:{q['target']}|{sorted(q['evidence'])}|{sorted(q['do'])}"
        if method == "anc":
            model.set_sampling_method("ancestral")
            got = model.sample(q, n_samples=S, noise=_dev_noise(inj.get("anc", {}), backend.device))
            _close(got, case["expect"]["samples"], tag + " samples", atol=atol)
            continue
        model.set_inference_method(METHOD_NAMES[method], n_samples=S)
        if method in ("is", "rb"):  # two passes with their own draws: {"is"|"rb": ..., "lw": fallback}
            noise = {k: _dev_noise(v, backend.device) for k, v in inj.items()}
        else:  # the exact methods draw only through their likelihood-weighting fallback (scope "lw")
            scope = "lw" if method in ("gexact", "cexact") else method
            noise = _dev_noise(inj.get(scope, {}), backend.device)
        pdf, samples = model.infer_posterior(q, noise=noise)
        _close(samples, case["expect"]["samples"], tag + " samples", atol=atol)
        _close(pdf, case["expect"]["pdf"], tag + " pdf", rtol=5e-5, atol=atol)
        if method == "ris":
            assert model._inference._last_resampled == case["info"]["resampled"], tag
        if method == "is":
            assert model._inference._last_fallback == case["info"]["fallback"], tag
            _close(model._inference._last_ess, case["info"]["ess"], tag + " ess", rtol=1e-4)


@pytest.mark.parametrize("name", [f for f in FILES])
def test_cpd_sample_log_prob_match_reference(backend, name):
    blob = _load(name)
    spec = blob["spec"]
    floor = _dot_floor(spec)
    for case in blob["cpd_cases"]:
        node, S = case["node"], case["S"]
        cpd = V.cpd_from_spec(spec["cpds"][node], device=backend.device)
        inj = to_injection(log_from_strkeys(case["noise"]), spec, S)["cpd"][node]
        parents = case["parents"]
        p_dev = None if parents is None else parents.to(backend.device)
        got = cpd.sample(p_dev, S, noise={k: v.to(backend.device) for k, v in inj.items()})
        tag = f"{name}:{node}:{None if parents is None else tuple(parents.shape)}"
        _close(got, case["samples"], tag + " sample", atol=max(ATOL, floor))
        x = case["samples"].to(backend.device)
        # d logp = (x - loc) / var * d loc: a loc floor of `floor` is worth ~floor * 4 sigma / var in logp
        lp_atol = max(2e-6, 40.0 * floor)
        _close(cpd.log_prob(x, p_dev), case["log_prob"], tag + " log_prob", rtol=2e-5, atol=lp_atol)
        p2 = p_dev if p_dev is None or p_dev.dim() == 2 else p_dev[:, :1]
        _close(cpd.log_prob(x[:, 0], p2), case["log_prob_2d"], tag + " log_prob 2d", rtol=2e-5, atol=lp_atol)


# ---- posterior summaries of the benchmark adapter ----------------------------------
def test_posterior_summaries_match_reference(backend):
    """benchmarking/models/vbn.py:202-242 (_estimate_discrete_posterior(_batch)) and :381-423
    (_continuous_from_samples): outputs of the reference's own functions stored in summaries.pt."""
    from vectorizedbayesiannetwork_b200 import summaries as SM

    for i, c in enumerate(_load("summaries")["cases"]):
        x = c["samples"].to(backend.device)
        w = None if c["weights"] is None else c["weights"].to(backend.device)
        if c["kind"] == "discrete":
            got = torch.tensor(SM._estimate_discrete_posterior_batch(x, w, c["k"]), dtype=torch.float64)
            _close(got, c["probs"], f"summaries[{i}] probs", rtol=1e-5, atol=1e-7)
            assert got.shape == (x.shape[0], c["k"])
            torch.testing.assert_close(got.sum(1), torch.ones(x.shape[0], dtype=torch.float64), rtol=0, atol=1e-5)
            one = SM._estimate_discrete_posterior(x, w, c["k"])  # first query only (:207-212)
            _close(torch.tensor(one, dtype=torch.float64), c["probs"][0], f"summaries[{i}] single", rtol=1e-5, atol=1e-7)
            _close(torch.tensor(O.estimate_discrete_posterior_batch(c["samples"], c["weights"], c["k"]),
                                dtype=torch.float64), c["probs"], f"summaries[{i}] oracle", rtol=0, atol=0)
        else:
            got = SM._continuous_from_samples(x, weights=w)
            want = c["out"]
            assert got["format"] == want["format"] == "normal_params" and got["n_samples"] == want["n_samples"]
            assert abs(got["mean"] - want["mean"]) <= 1e-6 + 1e-5 * abs(want["mean"]), (i, got["mean"], want["mean"])
            assert abs(got["std"] - want["std"]) <= 1e-6 + 1e-5 * abs(want["std"]), (i, got["std"], want["std"])
            assert got["samples"] == want["samples"]
            assert O.continuous_from_samples(c["samples"], weights=c["weights"]) == want


def test_posterior_summaries_argument_errors(backend):
    from vectorizedbayesiannetwork_b200 import summaries as SM

    x, w = torch.zeros(2, 5, device=backend.device), torch.ones(2, 5, device=backend.device)
    with pytest.raises(ValueError):  # :233-234
        SM._estimate_discrete_posterior_batch(x[0], w, 3)
    with pytest.raises(ValueError):  # :235-236
        SM._estimate_discrete_posterior_batch(x, w[0], 3)
    with pytest.raises(ValueError):  # :237-238
        SM._estimate_discrete_posterior_batch(x, w[:1], 3)
    with pytest.raises(ValueError):  # :376
        SM._continuous_from_samples(torch.zeros(4, 2, device=backend.device))
    probs = SM._estimate_discrete_posterior_batch(x, w, 3)  # every sample is class 0
    assert probs == [[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]


# ---- Gibbs sampler ---------------------------------------------------------------
def test_gibbs_chains_match_reference(backend):
    """vbn/sampling/gibbs.py:23-92: chains recorded from the reference (gibbs.pt), replayed with the reference's own
    draws (initial ancestral state, 8 candidates per latent node and sweep, the selected candidate)."""
    for m in _load("gibbs")["models"]:
        spec = m["spec"]
        model = V.VBN.from_spec(spec, device=backend.device)
        atol = max(ATOL, _dot_floor(spec))
        for i, case in enumerate(m["cases"]):
            inj = gibbs_injection(log_from_strkeys(case["noise"]), spec)
            noise = {scope: {n: {k: t.to(backend.device) for k, t in d.items()} for n, d in nodes.items()}
                     for scope, nodes in inj.items()}
            model.set_sampling_method("gibbs", n_samples=case["n"], **case["kw"])
            got = model.sample(case["query"], n_samples=case["n"], noise=noise)
            _close(got, case["expect"], f"gibbs:{m['name']}[{i}]", atol=atol)
            # the reference returns the chain's final state n times (views of the live state, gibbs.py:83-91)
            assert bool((got == got[:, :1]).all())


def test_gibbs_multidim_latent_matches_oracle(backend):
    """A 2-D latent node with a child (rff model: a, b -> c[2] -> d): candidates, scores and the selected value are
    D-wide.  The oracle (pinned to the reference) records its draws; the CUDA chain replays them."""
    spec = _load("rff")["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    # one query: with B > 1 and latent roots the reference (and the oracle) cannot run at all -- root candidates
    # are [1, 8, D] there, see inference.GibbsSampler
    q = {"target": "c", "evidence": {"d": torch.tensor([[0.4]])}, "do": {}}
    rec = O.RecordingNoise()
    torch.manual_seed(17)
    want = O.gibbs_sample(spec, q, 5, noise=rec, burn_in=2, n_steps=2)
    inj = gibbs_injection(rec.log, spec)
    noise = {scope: {n: {k: t.to(backend.device) for k, t in d.items()} for n, d in nodes.items()}
             for scope, nodes in inj.items()}
    model.set_sampling_method("gibbs", n_samples=5, burn_in=2, n_steps=2)
    got = model.sample(q, n_samples=5, noise=noise)
    assert got.shape == (1, 5, 2)
    _close(got, want, "gibbs 2-D latent", atol=max(ATOL, _dot_floor(spec)))
