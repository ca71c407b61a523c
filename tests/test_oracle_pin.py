"""Pins the oracle (oracle/vbn_oracle.py) to the UNMODIFIED reference, bit for bit, under the
same torch seed.  Runs only where /root/reference exists (the build container)."""
import pytest
import torch

import refmodels
from oracle import vbn_oracle as O

pytestmark = [
    pytest.mark.reference,
    pytest.mark.skipif(not refmodels.have_reference(), reason="/root/reference not present"),
]


def _eq(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.equal(torch.nan_to_num(a, nan=123.0), torch.nan_to_num(b, nan=123.0)), (
        (a - b).abs().max()
    )


def _run_methods(model, spec, query, S, methods=("lw", "is", "mcm", "anc")):
    tq = {"target": query["target"], "evidence": query.get("evidence", {}), "do": query.get("do", {})}
    if "lw" in methods:
        model.set_inference_method("likelihood_weighting", n_samples=S)
        torch.manual_seed(11)
        rw, rs = model.infer_posterior(tq)
        torch.manual_seed(11)
        ow, os_ = O.likelihood_weighting(spec, tq, S)
        _eq(rw, ow), _eq(rs, os_)
    if "is" in methods:
        model.set_inference_method("importance_sampling", n_samples=S)
        torch.manual_seed(12)
        rw, rs = model.infer_posterior(tq)
        torch.manual_seed(12)
        ow, os_, info = O.importance_sampling(spec, tq, S, return_info=True)
        _eq(rw, ow), _eq(rs, os_)
        assert info["fallback"] == model._inference._last_fallback
        _eq(info["ess"], model._inference._last_ess)
    if "mcm" in methods:
        model.set_inference_method("monte_carlo_marginalization", n_samples=S)
        torch.manual_seed(13)
        rw, rs = model.infer_posterior(tq)
        torch.manual_seed(13)
        ow, os_ = O.monte_carlo_marginalization(spec, tq, S)
        _eq(rw, ow), _eq(rs, os_)
    if "anc" in methods:
        model.set_sampling_method("ancestral")
        torch.manual_seed(14)
        rs = model.sample(tq, n_samples=S)
        torch.manual_seed(14)
        os_ = O.ancestral_sample(spec, tq, S)
        _eq(rs, os_)


def test_readme_minimal_example():
    model = refmodels.readme_model(epochs=5)
    spec = O.spec_from_reference(model)
    q = {"target": "feature_2",
         "evidence": {"feature_0": torch.tensor([[0.3]]), "feature_1": torch.tensor([[-0.2]])}}
    _run_methods(model, spec, q, 200)
    q2 = {"target": "feature_2", "evidence": {"feature_0": torch.randn(5, 1)}}
    _run_methods(model, spec, q2, 64)
    q3 = {"target": "feature_0", "evidence": {"feature_2": torch.randn(3, 1)}}
    _run_methods(model, spec, q3, 64)


def test_lg_chain_all_methods_and_exact_posterior():
    model = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(model)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    q = {"target": "x2", "evidence": {"x5": ev}}
    _run_methods(model, spec, q, 128)
    q = {"target": "x4", "evidence": {"x5": ev}, "do": {"x1": torch.tensor([[0.5], [0.5], [-1.0]])}}
    _run_methods(model, spec, q, 128)
    # closed form vs a large LW run of the reference itself
    model.set_inference_method("likelihood_weighting", n_samples=200_000)
    torch.manual_seed(0)
    w, s = model.infer_posterior({"target": "x2", "evidence": {"x5": ev}})
    mean = (w * s[..., 0]).sum(1)
    var = (w * (s[..., 0] - mean[:, None]) ** 2).sum(1)
    em, evar = O.lg_exact_posterior(spec, "x2", {"x5": ev})
    assert torch.allclose(mean.double(), em, atol=0.03)
    assert torch.allclose(var.double(), evar, rtol=0.05)


@pytest.mark.parametrize("activation", ["relu", "tanh", "gelu", "elu"])
def test_mixed_dag_all_cpd_kinds(activation):
    model = refmodels.mixed_model(activation=activation, epochs=2)
    spec = O.spec_from_reference(model)
    B = 3
    q = {"target": "e", "evidence": {"g": torch.randn(B, 2), "h": torch.randn(B, 1)}}
    _run_methods(model, spec, q, 32)
    q = {"target": "a", "evidence": {"f": torch.randn(B, 1) ** 2, "c": torch.randn(B, 1)},
         "do": {"d": torch.randn(B, 1)}}
    _run_methods(model, spec, q, 32)
    q = {"target": "h", "evidence": {}}
    _run_methods(model, spec, q, 32)


def test_discrete_softmax_model():
    model = refmodels.discrete_model()
    spec = O.spec_from_reference(model)
    q = {"target": "rain", "evidence": {"slip": torch.tensor([[1.0], [0.0]]),
                                        "wet": torch.tensor([[2.0], [1.0]])}}
    _run_methods(model, spec, q, 64)
    q = {"target": "slip", "evidence": {"rain": torch.tensor([[1.0], [0.0], [1.0]])}}
    _run_methods(model, spec, q, 64)
    with pytest.raises(ValueError):
        O.likelihood_weighting(spec, {"target": "rain", "evidence": {"wet": torch.tensor([[0.5]])}}, 8)


def test_categorical_table_model():
    """categorical_table (vbn/cpds/categorical_table.py) mixed with discrete softmax_nn nodes."""
    model = refmodels.table_model()
    spec = O.spec_from_reference(model)
    q = {"target": "rain", "evidence": {"slip": torch.tensor([[1.0], [0.0]]), "wet": torch.tensor([[2.0], [1.0]])}}
    _run_methods(model, spec, q, 64)
    q = {"target": "slip", "evidence": {"season": torch.tensor([[3.0], [0.0], [1.0]])}}
    _run_methods(model, spec, q, 64)
    with pytest.raises(ValueError):  # parent value outside the table's support (categorical_table.py:12-21)
        O.likelihood_weighting(spec, {"target": "slip", "evidence": {"rain": torch.tensor([[0.5]])}}, 8)


def test_categorical_embedded_softmax_model():
    """categorical_embedded_softmax (vbn/cpds/categorical_embedded_softmax.py:280-329, 469-511)."""
    model = refmodels.embedded_model()
    spec = O.spec_from_reference(model)
    _run_methods(model, spec, {"target": "v", "evidence": {"z": torch.tensor([[1.0], [0.0]])}}, 64)
    _run_methods(model, spec, {"target": "u", "evidence": {"y": torch.tensor([[2.0], [0.0], [1.0]])}}, 64)
    _run_methods(model, spec, {"target": "y", "evidence": {"t": torch.tensor([[1.0]])}, "do": {"u": torch.tensor([[2.0]])}}, 64)
    for name in ("categorical_exact", "rao_blackwellized_marginalization"):
        fn = getattr(O, name)
        _exact_eq(model, spec, name, {"target": "v", "evidence": {"u": torch.tensor([[0.0], [2.0]]), "t": torch.tensor([[1.0], [0.0]])}}, 16, fn)
        _exact_eq(model, spec, name, {"target": "u", "evidence": {}}, 16, fn)  # root: unmasked _logits
        _exact_eq(model, spec, name, {"target": "y", "evidence": {"t": torch.tensor([[1.0], [0.0]])}}, 16, fn)
    with pytest.raises(ValueError):  # parent value outside the support (:36-46)
        O.likelihood_weighting(spec, {"target": "y", "evidence": {"v": torch.tensor([[0.5]])}}, 8)
    for node, cpd in model.nodes.items():
        c = spec["cpds"][node]
        if c["kind"] != "categorical_embedded_softmax":
            continue
        parents = None
        if cpd.input_dim:
            parents = torch.stack([v[torch.randint(0, v.numel(), (5,))] for v in cpd._parent_values], dim=1)
        torch.manual_seed(3)
        rs = cpd.sample(parents, 6)
        torch.manual_seed(3)
        _eq(rs, O.cpd_sample(c, parents, 6))
        _eq(cpd.log_prob(rs, parents), O.cpd_log_prob(c, rs, parents))


def test_rff_gaussian_model():
    """rff_gaussian (vbn/cpds/rff_gaussian.py:131-146, 185-206, 254-291): root, 1-D and 2-D nodes."""
    model = refmodels.rff_model()
    spec = O.spec_from_reference(model)
    ev = torch.tensor([[0.4], [-1.1], [2.0]])
    _run_methods(model, spec, {"target": "b", "evidence": {"d": ev}}, 48)
    _run_methods(model, spec, {"target": "c", "evidence": {"a": ev, "e": -ev}}, 48)
    _run_methods(model, spec, {"target": "a", "evidence": {"c": torch.randn(2, 2)}}, 48)
    for name in ("gaussian_exact", "rao_blackwellized_marginalization"):
        fn = getattr(O, name)
        _exact_eq(model, spec, name, {"target": "b", "evidence": {"a": ev, "e": ev * 0.5}}, 21, fn)  # parents fixed
        _exact_eq(model, spec, name, {"target": "a", "evidence": {}}, 21, fn)                         # root
        _exact_eq(model, spec, name, {"target": "d", "evidence": {"a": ev}}, 21, fn)
    for node, cpd in model.nodes.items():
        c = spec["cpds"][node]
        parents = None if cpd.input_dim == 0 else torch.randn(4, cpd.input_dim)
        torch.manual_seed(3)
        rs = cpd.sample(parents, 6)
        torch.manual_seed(3)
        _eq(rs, O.cpd_sample(c, parents, 6))
        _eq(cpd.log_prob(rs, parents), O.cpd_log_prob(c, rs, parents))


@pytest.mark.parametrize("within_bin,clip", [("uniform", False), ("triangular", False),
                                             ("gaussian", False), ("uniform", True),
                                             ("triangular", True)])
def test_binned_softmax_model(within_bin, clip):
    model = refmodels.binned_model(within_bin=within_bin, clip=clip)
    spec = O.spec_from_reference(model)
    q = {"target": "p", "evidence": {"q": torch.randn(4, 2) * 0.5}}
    _run_methods(model, spec, q, 32)
    q = {"target": "q", "evidence": {"p": torch.randn(4, 2) * 3.0}}  # some outside the bins
    _run_methods(model, spec, q, 32)


def test_kde_model():
    model = refmodels.kde_model()
    spec = O.spec_from_reference(model)
    q = {"target": "p", "evidence": {"y": torch.randn(3, 1)}}
    _run_methods(model, spec, q, 40)
    q = {"target": "y", "evidence": {"p2": torch.randn(3, 1)}}
    _run_methods(model, spec, q, 40)


def test_cpd_level_sample_log_prob_forward():
    model = refmodels.mixed_model(epochs=2)
    spec = O.spec_from_reference(model)
    for node, cpd in model.nodes.items():
        c = spec["cpds"][node]
        dp = c["input_dim"]
        for parents in ([None] if dp == 0 else [torch.randn(4, dp), torch.randn(4, 7, dp)]):
            torch.manual_seed(3)
            rs = cpd.sample(parents, 7)
            torch.manual_seed(3)
            os_ = O.cpd_sample(c, parents, 7)
            _eq(rs.detach(), os_)
            _eq(cpd.log_prob(rs, parents).detach(), O.cpd_log_prob(c, os_, parents))
            x2 = rs[:, 0]
            _eq(cpd.log_prob(x2, parents if parents is None or parents.dim() == 2 else parents[:, :1]).detach(),
                O.cpd_log_prob(c, x2, parents if parents is None or parents.dim() == 2 else parents[:, :1]))


def test_record_replay_roundtrip():
    model = refmodels.mixed_model(epochs=2)
    spec = O.spec_from_reference(model)
    q = {"target": "e", "evidence": {"g": torch.randn(3, 2), "h": torch.randn(3, 1)}}
    rec = O.RecordingNoise()
    torch.manual_seed(5)
    w1, s1 = O.importance_sampling(spec, q, 16, noise=rec)
    w2, s2 = O.importance_sampling(spec, q, 16, noise=O.ReplayNoise(rec.log))
    _eq(w1, w2), _eq(s1, s2)


def _exact_eq(model, spec, method, q, S, oracle_fn, seed=3):
    tq = {"target": q["target"], "evidence": q.get("evidence", {}), "do": q.get("do", {})}
    model.set_inference_method(method, n_samples=S)
    torch.manual_seed(seed)
    rw, rs = model.infer_posterior(tq, n_samples=S)
    torch.manual_seed(seed)
    ow, os_ = oracle_fn(spec, tq, S)
    _eq(rw, ow), _eq(rs, os_)


def test_gaussian_exact_and_fallback():
    """vbn/inference/gaussian_exact.py: closed form when the target's parents are fixed, LW otherwise."""
    model = refmodels.lg_chain_model(n_nodes=5)
    spec = O.spec_from_reference(model)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    _exact_eq(model, spec, "gaussian_exact", {"target": "x3", "evidence": {"x2": ev}}, 31, O.gaussian_exact)
    _exact_eq(model, spec, "gaussian_exact", {"target": "x0", "evidence": {"x4": ev}}, 31, O.gaussian_exact)  # root
    _exact_eq(model, spec, "gaussian_exact", {"target": "x1", "evidence": {"x4": ev}}, 32, O.gaussian_exact)  # fallback
    _exact_eq(model, spec, "gaussian_exact", {"target": "x2", "evidence": {"x2": ev}}, 9, O.gaussian_exact)   # fixed target
    m2 = refmodels.readme_model(n=300, epochs=2)
    s2 = O.spec_from_reference(m2)
    _exact_eq(m2, s2, "gaussian_exact", {"target": "feature_0", "evidence": {"feature_2": ev}}, 17, O.gaussian_exact)  # gnn root
    q = {"target": "feature_2", "evidence": {"feature_0": ev, "feature_1": ev}}
    _exact_eq(m2, s2, "gaussian_exact", q, 16, O.gaussian_exact)  # mdn target -> fallback


def test_categorical_exact_and_fallback():
    """vbn/inference/categorical_exact.py."""
    model = refmodels.table_model()
    spec = O.spec_from_reference(model)
    q = {"target": "wet", "evidence": {"rain": torch.tensor([[1.0], [0.0]]), "sprinkler": torch.tensor([[0.0], [1.0]])}}
    _exact_eq(model, spec, "categorical_exact", q, 64, O.categorical_exact)   # categorical_table, parents fixed
    q = {"target": "slip", "evidence": {"wet": torch.tensor([[2.0], [0.0], [1.0]])}}
    _exact_eq(model, spec, "categorical_exact", q, 64, O.categorical_exact)   # softmax_nn, parents fixed
    q = {"target": "season", "evidence": {"slip": torch.tensor([[1.0]])}}
    _exact_eq(model, spec, "categorical_exact", q, 64, O.categorical_exact)   # root categorical_table
    q = {"target": "sprinkler", "evidence": {"slip": torch.tensor([[1.0], [0.0]])}}
    _exact_eq(model, spec, "categorical_exact", q, 64, O.categorical_exact)   # softmax_nn root (root_ready)
    q = {"target": "rain", "evidence": {"slip": torch.tensor([[1.0]])}}
    _exact_eq(model, spec, "categorical_exact", q, 48, O.categorical_exact)   # parent not fixed -> LW fallback


def test_resampled_importance_sampling():
    """vbn/inference/resampled_importance_sampling.py: ESS-triggered multinomial resampling after evidence nodes."""
    model = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(model)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    for q in ({"target": "x2", "evidence": {"x3": ev, "x5": ev * 2}},       # resamples (sharp evidence)
              {"target": "x5", "evidence": {"x1": ev}},                    # target after the cut
              {"target": "x0", "evidence": {"x4": ev}, "do": {"x2": ev}}):
        tq = {"target": q["target"], "evidence": q.get("evidence", {}), "do": q.get("do", {})}
        model.set_inference_method("resampled_importance_sampling", n_samples=64)
        torch.manual_seed(5)
        rw, rs = model.infer_posterior(tq)
        torch.manual_seed(5)
        ow, os_, info = O.resampled_importance_sampling(spec, tq, 64, return_info=True)
        _eq(rw, ow), _eq(rs, os_)
        assert info["resampled"] == model._inference._last_resampled
    mixed = refmodels.mixed_model(epochs=1)
    mspec = O.spec_from_reference(mixed)
    tq = {"target": "e", "evidence": {"g": torch.randn(3, 2), "h": torch.randn(3, 1)}, "do": {}}
    mixed.set_inference_method("resampled_importance_sampling", n_samples=32)
    torch.manual_seed(6)
    rw, rs = mixed.infer_posterior(tq)
    torch.manual_seed(6)
    ow, os_ = O.resampled_importance_sampling(mspec, tq, 32)
    _eq(rw, ow), _eq(rs, os_)


def test_rao_blackwellized_marginalization():
    """vbn/inference/rao_blackwellized_marginalization.py: gaussian mixture grid, categorical marginal, fallbacks."""
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    model = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(model)
    for q, S in (({"target": "x5", "evidence": {"x1": ev}}, 33),                  # gaussian mixture over sampled parents
                 ({"target": "x0", "evidence": {}, "do": {}}, 16),                # root target
                 ({"target": "x2", "evidence": {"x4": ev}}, 24),                  # observed descendant -> fallback
                 ({"target": "x3", "evidence": {"x3": ev}}, 8),                   # fixed target
                 ({"target": "x4", "evidence": {"x1": ev}, "do": {"x2": ev}}, 20)):
        _exact_eq(model, spec, "rao_blackwellized_marginalization", q, S, O.rao_blackwellized_marginalization)
    m2 = refmodels.table_model()
    s2 = O.spec_from_reference(m2)
    for q in ({"target": "slip", "evidence": {"season": torch.tensor([[3.0], [0.0], [1.0]])}},   # softmax_nn target
              {"target": "wet", "evidence": {"season": torch.tensor([[1.0], [2.0]])}},            # categorical_table target
              {"target": "season", "evidence": {}},                                               # root categorical
              {"target": "sprinkler", "evidence": {"season": torch.tensor([[1.0]])}}):            # softmax_nn root
        _exact_eq(m2, s2, "rao_blackwellized_marginalization", q, 40, O.rao_blackwellized_marginalization)
    m3 = refmodels.readme_model(n=300, epochs=2)
    s3 = O.spec_from_reference(m3)
    _exact_eq(m3, s3, "rao_blackwellized_marginalization", {"target": "feature_2", "evidence": {"feature_0": ev}}, 16,
              O.rao_blackwellized_marginalization)  # mdn target -> unsupported -> fallback
    _exact_eq(m3, s3, "rao_blackwellized_marginalization", {"target": "feature_0", "evidence": {}}, 16,
              O.rao_blackwellized_marginalization)  # gaussian_nn root


def test_posterior_summaries_of_the_benchmark_adapter():
    """benchmarking/models/vbn.py:116-121, 202-242, 365-423: float64-exact."""
    refmodels.import_reference()
    from benchmarking.models import vbn as RB

    g = torch.Generator().manual_seed(31)
    for (b, s, k) in ((1, 5, 2), (4, 333, 5), (2, 1200, 9)):
        x = torch.randint(-2, k + 2, (b, s), generator=g).float() + 0.5 * torch.randint(0, 2, (b, s), generator=g)
        w = torch.rand(b, s, generator=g)
        w[0, 0] = float("nan")
        assert RB._estimate_discrete_posterior_batch(x, w, k) == O.estimate_discrete_posterior_batch(x, w, k)
        assert RB._estimate_discrete_posterior(x, w, k) == O.estimate_discrete_posterior_batch(x[:1], w[:1], k)[0]
    model = RB.VBNBenchmarkModel.__new__(RB.VBNBenchmarkModel)
    for s in (3, 500, 5000):
        x = torch.randn(1, s, 1, generator=g)
        for w in (None, torch.rand(1, s, generator=g), torch.zeros(1, s), torch.rand(1, s, generator=g) - 0.5):
            assert model._continuous_from_samples(x, weights=w) == O.continuous_from_samples(x, weights=w)


def test_gibbs_sampler():
    """vbn/sampling/gibbs.py:23-92: same torch seed -> the same chain, bit for bit."""
    def run(model, spec, q, n, **kw):
        model.set_sampling_method("gibbs", n_samples=n, **kw)
        torch.manual_seed(21)
        ref = model.sample(q, n_samples=n)
        torch.manual_seed(21)
        got = O.gibbs_sample(spec, {"target": q["target"], "evidence": q.get("evidence", {}), "do": q.get("do", {})}, n, **kw)
        _eq(ref, got)

    m = refmodels.readme_model(n=300, epochs=2)
    spec = O.spec_from_reference(m)
    run(m, spec, {"target": "feature_2", "evidence": {"feature_0": torch.tensor([[0.3]]), "feature_1": torch.tensor([[-0.2]])}}, 12)
    run(m, spec, {"target": "feature_0", "evidence": {"feature_2": torch.tensor([[0.1]])}}, 9, burn_in=3, n_steps=2)
    m = refmodels.lg_chain_model(n_nodes=5)
    spec = O.spec_from_reference(m)
    run(m, spec, {"target": "x2", "evidence": {"x4": torch.tensor([[0.7]])}}, 10, burn_in=4)
    run(m, spec, {"target": "x3", "evidence": {"x0": torch.tensor([[0.2], [1.0], [-0.7]])}}, 6, burn_in=2)  # B = 3, no latent root
    m = refmodels.mixed_model(rows=256, epochs=1)
    spec = O.spec_from_reference(m)
    run(m, spec, {"target": "e", "evidence": {"g": torch.tensor([[0.3, -0.2]])}}, 5, burn_in=2)
    m = refmodels.table_model(rows=300, epochs=1)
    spec = O.spec_from_reference(m)
    run(m, spec, {"target": "rain", "evidence": {"slip": torch.tensor([[1.0]])}}, 7, burn_in=2)


def test_lbp_wrapper():
    """vbn/inference/lbp.py:11-70 with both fallbacks, converged and not (tol = 0 forces the retry pass)."""
    model = refmodels.lg_chain_model(n_nodes=5)
    spec = O.spec_from_reference(model)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    q = {"target": "x1", "evidence": {"x4": ev}, "do": {}}
    for kw in ({}, {"fallback": "monte_carlo_marginalization"}, {"n_iters": 2, "damping": 0.9}):
        model.set_inference_method("lbp", n_samples=40, **kw)
        for call_kw in ({}, {"tol": 0.0}):
            torch.manual_seed(8)
            rw, rs = model.infer_posterior(q, **call_kw)
            torch.manual_seed(8)
            ow, os_ = O.lbp(spec, q, 40, **kw, **call_kw)
            _eq(rw, ow), _eq(rs, os_)
