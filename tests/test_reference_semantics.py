"""Behavioural contract of the drop-in classes, written after the reference's own tests
(reference file:line cited per test).  Every test runs twice: on the B200 through libvbn_cuda.so
(``gpu``) and in the build container through the host emulation of the same device code."""
import math

import pytest
import torch

from backends import backend  # noqa: F401
from oracle import vbn_oracle as O

import vectorizedbayesiannetwork_b200 as V
from vectorizedbayesiannetwork_b200 import synthetic as S


def _chain(device, n=3):
    return V.VBN.from_spec(S.lg_chain(n), device=device)


def _mixed(device, n=12, seed=3):
    return V.VBN.from_spec(S.random_dag_lg_mdn(n, seed=seed), device=device)


# ---- tests/test_inference.py:27-59, tests/test_sampling.py:46-56 -----------------------------
@pytest.mark.parametrize("method", ["likelihood_weighting", "importance_sampling", "monte_carlo_marginalization"])
def test_inference_output_shapes(backend, method):
    model = _chain(backend.device)
    model.set_inference_method(method, n_samples=5)
    pdf, samples = model.infer_posterior({"target": "x2", "evidence": {"x0": torch.randn(4, 1)}})
    assert pdf.shape == (4, 5) and samples.shape == (4, 5, 1)
    assert torch.isfinite(pdf).all() and torch.isfinite(samples).all()
    assert not pdf.requires_grad and pdf.grad_fn is None  # tests/test_public_outputs.py:24-60
    assert not samples.requires_grad and samples.grad_fn is None


def test_sampling_output_shapes_and_joint(backend):
    model = _chain(backend.device)
    model.set_sampling_method("ancestral")
    s = model.sample({"target": "x2", "evidence": {"x0": torch.randn(4, 1)}}, n_samples=5)
    assert s.shape == (4, 5, 1) and not s.requires_grad
    with pytest.raises(ValueError):  # the facade insists on a target (vbn/vbn.py:596-597)
        model.sample({"target": None, "evidence": {"x0": torch.randn(2, 1)}}, n_samples=7)
    # the sampler itself returns the joint dict for a target-less Query (sampling/ancestral.py:62-65)
    q = V.Query(target="", evidence={"x0": torch.randn(2, 1).to(backend.device)}, do={})
    joint = model._sampling.sample(model, q, n_samples=7)
    assert sorted(joint) == ["x0", "x1", "x2"] and all(v.shape == (2, 7, 1) for v in joint.values())


def test_do_only_query_and_query_validation(backend):
    model = _chain(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=6)
    pdf, samples = model.infer_posterior({"target": "x2", "do": {"x1": torch.zeros(3, 1)}})
    assert pdf.shape == (3, 6) and samples.shape == (3, 6, 1)
    with pytest.raises(ValueError):  # evidence and do on the same node (vbn/vbn.py:579-618)
        model.infer_posterior({"target": "x2", "evidence": {"x1": torch.zeros(2, 1)}, "do": {"x1": torch.zeros(2, 1)}})
    with pytest.raises(ValueError):  # unknown node
        model.infer_posterior({"target": "nope", "evidence": {"x0": torch.zeros(2, 1)}})
    with pytest.raises(ValueError):  # batch mismatch (vbn/utils/__init__.py:46-61)
        model.infer_posterior({"target": "x2", "evidence": {"x0": torch.zeros(2, 1), "x1": torch.zeros(3, 1)}})
    fresh = _chain(backend.device)
    with pytest.raises(RuntimeError):  # method not set (vbn/vbn.py:474-481)
        fresh.infer_posterior({"target": "x2", "evidence": {"x0": torch.zeros(2, 1)}})


def test_n_samples_override_per_call(backend):
    model = _chain(backend.device)
    model.set_inference_method("importance_sampling", n_samples=8)
    pdf, _ = model.infer_posterior({"target": "x2", "evidence": {"x0": torch.zeros(2, 1)}}, n_samples=11)
    assert pdf.shape == (2, 11)  # importance_sampling.py:25


# ---- tests/test_performance_upgrades.py:23-46, 82-97 ------------------------------------------
def test_importance_sampling_batched_row0_equals_single_query(backend):
    model = _mixed(backend.device)
    ev = torch.tensor([[0.2], [-0.4]])
    model.set_inference_method("importance_sampling", n_samples=64)
    model._inference.ess_threshold = 0.0  # keep both calls on the IS pass
    torch.manual_seed(7)
    w2, s2 = model.infer_posterior({"target": "n5", "evidence": {"n11": ev}})
    torch.manual_seed(7)
    w1, s1 = model.infer_posterior({"target": "n5", "evidence": {"n11": ev[:1]}})
    torch.testing.assert_close(s2[:1], s1, rtol=0, atol=0)
    torch.testing.assert_close(w2[:1], w1, rtol=1e-6, atol=1e-9)
    # rows of one batch are drawn independently in IS (importance_sampling.py:37-54)
    assert not torch.equal(s2[0], s2[1])


def test_importance_sampling_ess_fallback(backend):
    model = _chain(backend.device, n=4)
    model.set_inference_method("importance_sampling", n_samples=256)
    far = {"target": "x0", "evidence": {"x3": torch.full((2, 1), 40.0)}}  # evidence deep in the tail
    w, s = model.infer_posterior(far)
    assert model._inference._last_fallback is True  # ESS < max(1, 0.1 S) -> likelihood weighting
    assert model._inference._last_ess.shape == (2,) and w.shape == (2, 256)
    near = {"target": "x0", "evidence": {"x3": torch.full((2, 1), 0.3)}}
    model._inference.ess_threshold = 0.0
    model.infer_posterior(near)
    assert model._inference._last_fallback is False
    assert bool((model._inference._last_ess >= 1.0).all())


def test_likelihood_weighting_shares_root_draws_across_queries(backend):
    """LW / MCM / ancestral broadcast one [1,S,D] root draw to every query."""
    model = _chain(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=32)
    _, s = model.infer_posterior({"target": "x0", "evidence": {"x2": torch.randn(3, 1)}})
    assert torch.equal(s[0], s[1]) and torch.equal(s[0], s[2])


def test_root_cpd_sample_shape_and_handle(backend):
    model = _mixed(backend.device)
    root = next(n for n in model.dag.nodes() if not model.dag.parents(n))
    h = model.get_cpd(root)
    assert h.sample(None, 5).shape == (1, 5, 1)  # linear_gaussian.py:187, mdn.py:211
    child = next(n for n in model.dag.nodes() if len(model.dag.parents(n)) == 1)
    hc = model.get_cpd(child)
    parent = model.dag.parents(child)[0]
    x = hc.conditional_samples({parent: torch.zeros(4, 1)}, 9)
    assert x.shape == (4, 9, 1) and not x.requires_grad
    lp = hc.log_prob(x, {parent: torch.zeros(4, 1)})
    assert lp.shape == (4, 9) and torch.isfinite(lp).all()
    assert hc.log_prob(x[:, 0], torch.zeros(4, 1)).shape == (4, 1)  # 2-D x -> [B,1] (tests/test_cpds.py:9-47)
    with pytest.raises(ValueError):  # tests/test_cpd_handle.py:97-101
        hc.sample({}, 3)
    out = hc.forward(torch.zeros(2, 1), 6)
    torch.testing.assert_close(out.pdf, torch.exp(out.log_prob))


# ---- vbn/inference/monte_carlo_marginalization.py:33-92 ----------------------------------------
def test_mcm_three_paths_and_shape_quirk(backend):
    model = _chain(backend.device)
    model.set_inference_method("monte_carlo_marginalization", n_samples=10)
    pdf, s = model.infer_posterior({"target": "x1", "do": {"x1": torch.full((3, 1), 2.5)}})
    assert torch.equal(pdf, torch.ones(3, 10, device=pdf.device)) and bool((s == 2.5).all())  # :33-37
    pdf, s = model.infer_posterior({"target": "x1", "evidence": {"x0": torch.randn(3, 1)}})  # parents fixed :39-58
    assert pdf.shape == (3, 10) and bool((pdf > 0).all())
    pdf, s = model.infer_posterior({"target": "x0", "evidence": {"x2": torch.randn(3, 1)}})  # parentless target
    assert pdf.shape == (1, 10) and s.shape == (1, 10, 1)
    pdf, s = model.infer_posterior({"target": "x2", "evidence": {"x0": torch.randn(3, 1)}})  # full pass :60-92
    assert pdf.shape == (3, 10) and s.shape == (3, 10, 1)
    # pdf is the target CPD density at the drawn value given the drawn parents
    model2 = _chain(backend.device)
    model2.set_sampling_method("ancestral")


# ---- tests/test_sampling.py:59-75 --------------------------------------------------------------
def test_ancestral_do_semantics(backend):
    model = _chain(backend.device)
    model.set_sampling_method("ancestral")
    s = model.sample({"target": "x1", "do": {"x1": torch.tensor([[1.0], [-1.0]])}}, n_samples=16)
    assert bool((s[0] == 1.0).all()) and bool((s[1] == -1.0).all())
    y = model.sample({"target": "x2", "do": {"x1": torch.tensor([[1.0], [-1.0]])}}, n_samples=4000)
    assert float(y[0].mean() - y[1].mean()) > 0.5  # slope 1 chain: do(x1 = +-1) moves x2 by 2


# ---- closed-form Gaussian posterior (BASELINE north_star) ------------------------
@pytest.mark.parametrize("method", ["likelihood_weighting", "importance_sampling"])
def test_weighted_posterior_matches_closed_form_gaussian(backend, method):
    spec = S.lg_chain(8)
    model = V.VBN.from_spec(spec, device=backend.device)
    n = 200_000 if backend.name == "cuda" else 20_000
    ev = torch.tensor([[0.5], [1.5], [3.0]])
    model.set_inference_method(method, n_samples=n)
    w, s = model.infer_posterior({"target": "x3", "evidence": {"x7": ev}}, seed=5)
    w, s = w.double().cpu(), s[..., 0].double().cpu()
    mean = (w * s).sum(1)
    var = (w * (s - mean[:, None]) ** 2).sum(1)
    em, evar = O.lg_exact_posterior(spec, "x3", {"x7": ev})
    ess = 1.0 / (w**2).sum(1)
    tol = 5.0 * evar.sqrt() / ess.sqrt()  # 5 sigma of the self-normalised estimator
    assert bool(((mean - em).abs() < tol).all()), (mean, em, tol)
    assert bool(((var - evar).abs() < 0.1 * evar).all()), (var, evar)
    torch.testing.assert_close(w.sum(1), torch.ones(3, dtype=torch.float64), rtol=1e-4, atol=1e-4)


# ---- tests/test_performance_upgrades.py:49-79 (KDE known-answer formula) -------------------------
def test_kde_bulk_log_prob_matches_closed_expression(backend):
    n_points, rows = (3000, 4200) if backend.name == "cuda" else (300, 4100)
    spec = S.kde_pair(n_points)
    cpd = V.cpd_from_spec(spec["cpds"]["y"], device=backend.device)
    g = torch.Generator().manual_seed(2)
    x, p = torch.randn(rows, 1, generator=g), torch.randn(rows, 1, generator=g)
    # far-tail rows: every kernel term underflows with a fixed shift -> the exact online-max pass
    x[0], p[1], x[2], p[2] = 40.0, -30.0, -25.0, 35.0
    x[700], p[701] = 60.0, 45.0
    got = cpd.log_prob(x, p).cpu()  # >= 4096 rows -> the tiled stand-alone kernel
    want = O.kde_log_prob(spec["cpds"]["y"], x, p)
    # rows with far PARENTS subtract two logsumexps of magnitude ~4e3: fp32 cancellation leaves ~1e-4
    # absolute noise in the reference itself, so those rows get an absolute tolerance
    far_p = torch.zeros(rows, dtype=torch.bool)
    far_p[[1, 2, 701]] = True
    torch.testing.assert_close(got[~far_p], want[~far_p], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(got[far_p], want[far_p], rtol=1e-5, atol=1e-3)
    small = cpd.log_prob(x[:17], p[:17]).cpu()  # schedule-kernel path (op_kde)
    torch.testing.assert_close(small[~far_p[:17]], want[:17][~far_p[:17]], rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(small[far_p[:17]], want[:17][far_p[:17]], rtol=1e-5, atol=1e-3)


def test_kde_root_log_prob(backend):
    spec = S.kde_pair(500)
    c = dict(spec["cpds"]["y"], input_dim=0, parents=None)
    cpd = V.cpd_from_spec(c, device=backend.device)
    x = torch.linspace(-2, 2, 4100)[:, None]
    torch.testing.assert_close(cpd.log_prob(x, None).cpu(), O.kde_log_prob(c, x, None), rtol=1e-5, atol=2e-6)


# ---- softmax_nn.py:620-625 ----------------------------------------------------------------------
def test_discrete_value_outside_class_set_raises(backend):
    spec = S.alarm_softmax(seed=0)
    model = V.VBN.from_spec(spec, device=backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=8)
    ok = {"target": "LVFAILURE", "evidence": {"HRBP": torch.tensor([[1.0], [2.0]])}}
    w, s = model.infer_posterior(ok)
    assert set(s.unique().tolist()) <= {0.0, 1.0}  # discrete mode returns exact class values
    with pytest.raises(ValueError):
        model.infer_posterior({"target": "LVFAILURE", "evidence": {"HRBP": torch.tensor([[0.5], [2.0]])}})


def test_same_seed_is_deterministic_and_seeds_differ(backend):
    model = _mixed(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=128)
    q = {"target": "n5", "evidence": {"n11": torch.tensor([[0.1], [0.2]])}}
    w1, s1 = model.infer_posterior(q, seed=99)
    w2, s2 = model.infer_posterior(q, seed=99)
    w3, s3 = model.infer_posterior(q, seed=100)
    assert torch.equal(s1, s2) and torch.equal(w1, w2) and not torch.equal(s1, s3)


def test_no_evidence_gives_uniform_weights(backend):
    model = _chain(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=16)
    w, _ = model.infer_posterior({"target": "x2", "do": {"x0": torch.zeros(2, 1)}})
    torch.testing.assert_close(w, torch.full_like(w, 1.0 / 16))


def test_ragged_row_counts_cover_tile_tails(backend):
    """Row counts that are not multiples of any tile size (1, 127, 129, 513 rows)."""
    model = _mixed(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=1)
    for b, s in ((1, 1), (1, 127), (3, 43), (1, 513)):
        w, x = model.infer_posterior({"target": "n5", "evidence": {"n11": torch.zeros(b, 1)}}, n_samples=s, seed=3)
        assert w.shape == (b, s) and torch.isfinite(w).all() and torch.isfinite(x).all()
        torch.testing.assert_close(w.sum(1), torch.ones(b, device=w.device), rtol=1e-5, atol=1e-5)
    # the first 43 samples of query 0 do not depend on how many rows the launch had
    _, a = model.infer_posterior({"target": "n5", "evidence": {"n11": torch.zeros(1, 1)}}, n_samples=43, seed=3)
    _, bb = model.infer_posterior({"target": "n5", "evidence": {"n11": torch.zeros(1, 1)}}, n_samples=513, seed=3)
    torch.testing.assert_close(a[0], bb[0, :43], rtol=1e-6, atol=1e-7)


def test_registry_install_is_a_drop_in_for_the_reference(backend):
    """The live drop-in: the UNMODIFIED reference package (baseline/_ref on the GPU box, /root/reference in the build
    container), its own VBN facade and CPD modules on ``backend.device``, with this library's classes swapped into
    its registries."""
    from oracle import reference_arm as R

    vbn = R.load_reference()
    if vbn is None:
        pytest.skip("reference package not present (baseline/_ref, /root/reference)")
    ref = R.reference_from_spec(S.lg_chain(4), device=backend.device)
    V.install(vbn)
    try:
        assert vbn.core.registry.INFERENCE_REGISTRY["importance_sampling"] is V.ImportanceSampling
        ref.set_inference_method("likelihood_weighting", n_samples=64)  # resolved by the reference's own facade
        assert isinstance(ref._inference, V.LikelihoodWeighting)
        q = {"target": "x0", "evidence": {"x3": torch.tensor([[0.3], [0.6]], device=backend.device)}}
        pdf, samples = ref.infer_posterior(q)
        assert pdf.shape == (2, 64) and samples.shape == (2, 64, 1)
        torch.testing.assert_close(pdf.sum(1).cpu(), torch.ones(2), rtol=1e-4, atol=1e-4)
        ref.set_sampling_method("ancestral", n_samples=16)
        assert ref.sample({"target": "x3", "evidence": {}}, n_samples=16).shape == (1, 16, 1)
    finally:
        V.uninstall(vbn)
    assert vbn.core.registry.INFERENCE_REGISTRY["importance_sampling"] is not V.ImportanceSampling


# ---- full BASELINE sizes: size-independent properties (B200 only) --------------------------------
@pytest.mark.gpu
def test_cfg2_full_size_properties():
    """64 queries x 1,000,000 samples on the 50-node chain: normalisation, ESS range, fallback
    bookkeeping and the closed-form posterior mean per query."""
    dev = torch.device("cuda", 0)
    spec = S.lg_chain(50)
    model = V.VBN.from_spec(spec, device=dev)
    g = torch.Generator().manual_seed(1)
    sd = (1 + 49 * 0.25) ** 0.5
    ev = 4.9 + sd * (torch.rand(64, 1, generator=g) * 3 - 1.5)
    model.set_inference_method("importance_sampling", n_samples=1_000_000)
    w, s = model.infer_posterior({"target": "x25", "evidence": {"x49": ev}}, seed=11)
    assert w.shape == (64, 1_000_000) and s.shape == (64, 1_000_000, 1)
    torch.testing.assert_close(w.sum(1).cpu(), torch.ones(64), rtol=2e-4, atol=2e-4)
    wd, sd_ = w.double(), s[..., 0].double()
    ess = 1.0 / (wd**2).sum(1)
    assert bool((ess >= 1).all()) and bool((ess <= 1_000_000).all())
    assert model._inference._last_fallback == bool((model._inference._last_ess < 100_000).any())
    mean = (wd * sd_).sum(1).cpu()
    em, evar = O.lg_exact_posterior(spec, "x25", {"x49": ev})
    assert bool(((mean - em).abs() < 6.0 * evar.sqrt() / ess.sqrt().cpu() + 1e-3).all())


@pytest.mark.gpu
def test_cfg5_block_properties():
    dev = torch.device("cuda", 0)
    spec = S.random_dag_lg_mdn(1000, seed=0)
    model = V.VBN.from_spec(spec, device=dev)
    g = torch.Generator().manual_seed(1)
    ev = {n: 0.3 * torch.randn(96, 1, generator=g) for n in spec["nodes"][-5:]}
    model.set_inference_method("importance_sampling", n_samples=4096)
    w, s = model.infer_posterior({"target": "n500", "evidence": ev}, seed=4)
    assert torch.isfinite(w).all() and torch.isfinite(s).all()
    torch.testing.assert_close(w.sum(1).cpu(), torch.ones(96), rtol=1e-4, atol=1e-4)
    w2, s2 = model.infer_posterior({"target": "n500", "evidence": ev}, seed=4)
    assert torch.equal(s, s2) and torch.equal(w, w2)  # bit-reproducible for a fixed seed
    # a block of queries gives the same rows as the full batch (what query sharding relies on)
    sub = {k: v[32:64] for k, v in ev.items()}
    model._inference.ess_threshold = 0.0
    wa, sa = model.infer_posterior({"target": "n500", "evidence": ev}, seed=4)
    wb, sb = model.infer_posterior({"target": "n500", "evidence": sub}, seed=4,
                                   shard=None)
    assert wa.shape == (96, 4096) and wb.shape == (32, 4096)
    # ... provided the block is run with its global query offset (dist.Shard does exactly this): per-row Philox
    # streams are keyed by the GLOBAL (query, sample) index
    runner = model._inference._runner
    subq = V.core.Query(target="n500", evidence={k: v.to(dev) for k, v in sub.items()}, do={})
    plan = runner.plan_for(model, subq, "is")
    logw = torch.empty(32, 4096, device=dev)
    sc = torch.empty(32, 4096, 1, device=dev)
    fixed = runner.fixed_table(plan, subq, 32, clamp_obs=False, shard=None)
    plan.run(32, 4096, fixed=fixed, stores=[sc], logw=logw, seed=4, query_offset=32)
    wc = torch.softmax(logw, dim=1)
    assert torch.equal(sc, sa[32:64])
    torch.testing.assert_close(wc, wa[32:64], rtol=2e-5, atol=1e-9)  # same log-weights; softmax by torch here


@pytest.mark.gpu
def test_cfg5_posterior_mean_matches_oracle_estimate():
    """BASELINE cfg5 (1000-node LG + MDN DAG, importance sampling, tcgen05 kernel with every fast path on): the
    self-normalised posterior mean of the target for a few queries against the oracle's own importance-sampling
    estimate (independent draws) within the combined Monte-Carlo standard error."""
    dev = torch.device("cuda", 0)
    spec = S.random_dag_lg_mdn(1000, seed=0)
    g = torch.Generator().manual_seed(1)
    ev = {n: 0.3 * torch.randn(3, 1, generator=g) for n in spec["nodes"][-5:]}
    model = V.VBN.from_spec(spec, device=dev)
    model.set_inference_method("importance_sampling", n_samples=65536, ess_threshold=0.0)
    w, x = model.infer_posterior({"target": "n500", "evidence": ev}, seed=9)
    w, x = w.double().cpu(), x[..., 0].double().cpu()
    mean = (w * x).sum(1)
    var = (w * (x - mean[:, None]) ** 2).sum(1)
    ess = 1.0 / (w**2).sum(1)
    for b in range(3):
        q = {"target": "n500", "evidence": {n: v[b:b + 1] for n, v in ev.items()}, "do": {}}
        torch.manual_seed(50 + b)
        ow, ox = O.importance_sampling(spec, q, 8192, ess_threshold=0.0)
        ow, ox = ow.double()[0], ox[0, :, 0].double()
        om = float((ow * ox).sum())
        ovar = float((ow * (ox - om) ** 2).sum())
        oess = float(1.0 / (ow**2).sum())
        se = (float(var[b]) / float(ess[b]) + ovar / oess) ** 0.5
        assert abs(float(mean[b]) - om) < 5.0 * se + 1e-3, (b, float(mean[b]), om, se, float(ess[b]), oess)


@pytest.mark.gpu
def test_cfg3_full_size_posterior_matches_oracle_estimate():
    """BASELINE cfg3 at full size (ALARM, 4096 queries x 16 384 samples, likelihood weighting), summarised on the
    device by the class histogram (benchmarking/models/vbn.py:202-242).  Size-independent properties, plus the
    posterior of the first queries against the oracle's own likelihood-weighting estimate (independent draws):
    the two Monte-Carlo estimates must agree within their combined standard error."""
    from vectorizedbayesiannetwork_b200 import summaries as SM

    dev = torch.device("cuda", 0)
    spec = S.alarm_softmax(seed=0)
    g = torch.Generator().manual_seed(11)
    ev = {n: torch.randint(0, S.ALARM[n][0], (4096, 1), generator=g).float() for n in ("HRBP", "BP", "EXPCO2", "PRESS")}
    model = V.VBN.from_spec(spec, device=dev)
    model.set_inference_method("likelihood_weighting", n_samples=16384)
    w, x = model.infer_posterior({"target": "LVFAILURE", "evidence": ev}, seed=3)
    assert w.shape == (4096, 16384) and x.shape == (4096, 16384, 1)
    assert bool(((x == 0) | (x == 1)).all()) and torch.isfinite(w).all() and bool((w >= 0).all())
    k = S.ALARM["LVFAILURE"][0]
    probs = SM.estimate_discrete_posterior_tensor(x, w, k)
    torch.testing.assert_close(probs.sum(1).cpu(), torch.ones(4096), rtol=0, atol=1e-5)
    # the histogram kernel against a torch reduction over the same tensors
    wn = w / w.sum(1, keepdim=True)
    want1 = (wn * (x[..., 0] == 1)).sum(1)
    torch.testing.assert_close(probs[:, 1], want1, rtol=1e-4, atol=1e-6)
    ess = (1.0 / (wn**2).sum(1)).cpu()
    for b in range(3):
        q = {"target": "LVFAILURE", "evidence": {n: v[b:b + 1] for n, v in ev.items()}, "do": {}}
        torch.manual_seed(100 + b)
        ow, ox = O.likelihood_weighting(spec, q, 16384)
        op = O.estimate_discrete_posterior_batch(ox, ow, k)[0][1]
        on = ow / ow.sum(1, keepdim=True)
        oess = float(1.0 / (on**2).sum())
        pg = float(probs[b, 1])
        pm = 0.5 * (pg + op)
        se = (pm * (1 - pm) * (1.0 / float(ess[b]) + 1.0 / oess)) ** 0.5
        assert abs(pg - op) < 5.0 * se + 1e-3, (b, pg, op, se)


@pytest.mark.gpu
def test_cfg4_size_spot_check():
    dev = torch.device("cuda", 0)
    spec = S.kde_pair(200_000)
    cpd = V.cpd_from_spec(spec["cpds"]["y"], device=dev)
    g = torch.Generator().manual_seed(1)
    x, p = torch.randn(20_000, 1, generator=g), torch.randn(20_000, 1, generator=g)
    got = cpd.log_prob(x, p).cpu()
    assert got.shape == (20_000, 1) and torch.isfinite(got).all()
    want = O.kde_log_prob(spec["cpds"]["y"], x[:256], p[:256])
    torch.testing.assert_close(got[:256], want, rtol=1e-5, atol=2e-6)


# ---- discrete-parent lookup tables (VBN_OP_TAB) vs the per-row MLP path ---------------------------
def test_discrete_tables_reproduce_the_mlp_path(backend, monkeypatch):
    """ALARM (all softmax_nn discrete): with the same Philox seed the table-compiled schedule must
    draw the same states and give the same weights as the schedule that evaluates every MLP per row."""
    from vectorizedbayesiannetwork_b200 import _lib as L

    spec = S.alarm_softmax(seed=0)
    g = torch.Generator().manual_seed(4)
    ev = {n: torch.randint(0, S.ALARM[n][0], (5, 1), generator=g).float() for n in ("HRBP", "BP", "EXPCO2", "PRESS")}
    q = {"target": "LVFAILURE", "evidence": ev}
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VBN_TABLE", flag)
        model = V.VBN.from_spec(spec, device=backend.device)
        model.set_inference_method("likelihood_weighting", n_samples=300)
        w, s = model.infer_posterior(q, seed=21)
        kinds = [int(op["kind"]) for op in next(iter(model._inference._runner._cache.values())).program.ops]
        out[flag] = (w.cpu(), s.cpu(), kinds)
    assert L.OP_TAB not in out["0"][2] and L.OP_SNN in out["0"][2]
    assert all(k == L.OP_TAB for k in out["1"][2])  # every ALARM node became a lookup
    same = (out["0"][1] == out["1"][1]).float().mean().item()
    assert same > 0.999, same  # a pick can flip only when u sits within rounding of a CDF edge
    ok = ((out["0"][0] - out["1"][0]).abs() <= 1e-7 + 1e-4 * out["0"][0].abs()).float().mean().item()
    assert ok > 0.995, ok


# ---- parent-less MDN nodes precomputed by the plan compiler (VBN_F_MDNROOT) vs the per-row path ----
def test_root_mdn_fast_path_reproduces_the_generic_path(backend, monkeypatch):
    """mdn.py:190-196, 227-235: the root mixture is row-independent; with the same Philox seed the
    op-embedded mixture must draw what the per-row evaluation draws."""
    from vectorizedbayesiannetwork_b200 import _lib as L

    spec = S.random_dag_lg_mdn(40, seed=5)
    assert any(spec["cpds"][n]["kind"] == "mdn" and not spec["parents"][n] for n in spec["nodes"])
    g = torch.Generator().manual_seed(2)
    ev = {n: 0.3 * torch.randn(3, 1, generator=g) for n in spec["nodes"][-2:]}
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VBN_MDNROOT", flag)
        model = V.VBN.from_spec(spec, device=backend.device)
        model.set_inference_method("importance_sampling", n_samples=700, ess_threshold=0.0)
        w, s = model.infer_posterior({"target": "n20", "evidence": ev}, seed=9)
        ops = next(iter(model._inference._runner._cache.values())).program.ops
        out[flag] = (w.cpu(), s.cpu(), sum(1 for op in ops if int(op["flags"]) & L.F_MDNROOT))
    assert out["0"][2] == 0 and out["1"][2] > 0
    close = ((out["0"][1] - out["1"][1]).abs() <= 1e-6 + 1e-5 * out["0"][1].abs()).float().mean().item()
    assert close > 0.995, close  # a pick can flip only when u sits within rounding of a CDF edge
    ok = ((out["0"][0] - out["1"][0]).abs() <= 1e-7 + 1e-3 * out["0"][0].abs()).float().mean().item()
    assert ok > 0.99, ok


def test_plain_table_ops_reproduce_the_generic_lookup(backend, monkeypatch):
    """VBN_F_TABPLAIN (class index = value, descriptor-embedded strides) vs the generic VBN_OP_TAB body on the same
    Philox streams: identical draws and weights, bit for bit."""
    from vectorizedbayesiannetwork_b200 import _lib as L

    spec = S.alarm_softmax(seed=0)
    g = torch.Generator().manual_seed(4)
    ev = {n: torch.randint(0, S.ALARM[n][0], (6, 1), generator=g).float() for n in ("HRBP", "BP", "EXPCO2", "PRESS")}
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("VBN_TABPLAIN", flag)
        model = V.VBN.from_spec(spec, device=backend.device)
        model.set_inference_method("likelihood_weighting", n_samples=257)
        w, s = model.infer_posterior({"target": "LVFAILURE", "evidence": ev}, seed=21)
        ops = next(iter(model._inference._runner._cache.values())).program.ops
        out[flag] = (w.cpu(), s.cpu(), sum(1 for op in ops if int(op["flags"]) & L.F_TABPLAIN))
    assert out["0"][2] == 0 and out["1"][2] >= 30
    assert torch.equal(out["0"][1], out["1"][1]) and torch.equal(out["0"][0], out["1"][0])


# ---- CPDHandle.conditional formats (vbn/core/cpd_handle.py:40-118, 348-402; tests/test_cpd_handle.py:61-88)
def test_conditional_formats_match_the_parameter_heads(backend):
    import os

    blob = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mixed_relu.pt"),
                      weights_only=False)
    spec = blob["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    g = torch.Generator().manual_seed(9)
    seen = set()
    for node in spec["nodes"]:
        c = spec["cpds"][node]
        h = model.get_cpd(node)
        dp = c["input_dim"]
        parents = torch.randn(5, dp, generator=g) if dp else None
        out = h.conditional(parents)
        seen.add(out["format"])
        t = lambda key: torch.tensor(out[key])
        if c["kind"] == "linear_gaussian":
            loc, scale = O.lg_params(c, parents)
            assert out["format"] == "normal_params"
            torch.testing.assert_close(t("mean"), loc.reshape(t("mean").shape), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(t("std"), scale.reshape(t("std").shape), rtol=1e-5, atol=1e-6)
        elif c["kind"] == "gaussian_nn":
            loc, scale = O.gnn_params(c, parents)
            assert out["format"] == "normal_params"
            torch.testing.assert_close(t("mean"), loc.reshape(t("mean").shape), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(t("std"), scale.reshape(t("std").shape), rtol=1e-5, atol=1e-6)
        elif c["kind"] == "mdn":
            logits, loc, scale = O.mdn_params(c, parents)
            assert out["format"] == "mixture_params"
            w = torch.softmax(logits, dim=-1)
            torch.testing.assert_close(t("weights"), w.reshape(t("weights").shape), rtol=2e-5, atol=1e-6)
            torch.testing.assert_close(t("loc"), loc.reshape(t("loc").shape), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(t("scale"), scale.reshape(t("scale").shape), rtol=1e-5, atol=1e-6)
        elif c["kind"] == "softmax_nn":
            b = 5 if dp else 1
            probs = torch.softmax(O.snn_logits(c, parents, b, 1), dim=-1)
            assert out["format"] == "categorical_probs" and out["k"] == c["n_classes"]
            torch.testing.assert_close(t("probs"), probs.reshape(t("probs").shape), rtol=2e-5, atol=1e-6)
        else:
            assert out["format"] == "empirical_samples"
        ms = h.conditional_mean_std(parents, n_samples=64)
        assert torch.isfinite(ms["mean"]).all() and torch.isfinite(ms["std"]).all()
    assert seen == {"normal_params", "mixture_params", "categorical_probs", "empirical_samples"}


# ---- rff_gaussian (vbn/cpds/rff_gaussian.py) ---------------------------------------
def test_rff_gaussian_conditional_and_unfitted(backend):
    import os

    spec = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rff.pt"),
                      weights_only=False)["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    g = torch.Generator().manual_seed(3)
    for node in ("a", "b", "c"):
        c = spec["cpds"][node]
        parents = torch.randn(6, c["input_dim"], generator=g) if c["input_dim"] else None
        out = model.get_cpd(node).conditional(parents)  # a callable _params -> "normal_params" (cpd_handle.py:59-66)
        loc, scale = O.rff_params(c, parents)
        assert out["format"] == "normal_params"
        mean, std = torch.tensor(out["mean"]), torch.tensor(out["std"])
        torch.testing.assert_close(mean, loc.reshape(mean.shape), rtol=1e-5, atol=4e-6)
        torch.testing.assert_close(std, scale.expand(loc.shape).reshape(std.shape), rtol=1e-5, atol=1e-6)
    # Philox path: the drawn column has the conditional's moments
    cpd = V.cpd_from_spec(spec["cpds"]["b"], device=backend.device)
    pa = torch.tensor([[0.3, -0.5]])
    s = cpd.sample(pa, 20000, seed=5).cpu()
    loc, scale = O.rff_params(spec["cpds"]["b"], pa)
    assert abs(s.mean().item() - loc.item()) < 5 * scale.item() / 20000**0.5
    assert abs(s.std().item() / scale.item() - 1.0) < 0.03
    unfitted = dict(spec["cpds"]["b"], stats_ready=False)
    with pytest.raises(RuntimeError):  # rff_gaussian.py:76-78
        V.cpd_from_spec(unfitted, device=backend.device).sample(pa, 4)


# ---- CPDHandle bookkeeping surface (vbn/core/cpd_handle.py:130-193, 276-346; tests/test_cpd_handle.py) ----
def test_cpd_handle_summary_state_and_clone(backend):
    import os

    spec = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mixed_relu.pt"),
                      weights_only=False)["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    for node in spec["nodes"]:
        h = model.get_cpd(node)
        s = h.summary()
        assert s["node"] == node and s["is_fitted"] is True and s["cpd_name"] == spec["cpds"][node]["kind"]
        assert s["input_dim"] == spec["cpds"][node]["input_dim"] and s["output_dim"] == spec["cpds"][node]["output_dim"]
        assert h.name == node and h.parents == spec["parents"][node] and h.x_dim == h.output_dim
        sd = h.state_dict()
        assert sd and all(isinstance(v, torch.Tensor) for v in sd.values())
        cfg = h.export_config()
        assert cfg["node"] == node and cfg["cpd_name"] == h.cpd_name and isinstance(cfg["init_kwargs"], dict)
        clone = h.clone_cpd()
        assert clone is not h.cpd and type(clone) is type(h.cpd)
        pa = torch.randn(3, h.input_dim) if h.input_dim else None
        a = h.cpd.sample(pa, 5, seed=4)
        b = clone.sample(pa, 5, seed=4)
        assert torch.equal(a, b)  # same parameters, same Philox stream
    unfitted = dict(spec["cpds"]["f"], bins_ready=False)  # softmax_nn before fit (softmax_nn.py:178-183)
    model.nodes["f"] = V.cpd_from_spec(unfitted, device=backend.device)
    assert model.get_cpd("f").is_fitted is False and model.get_cpd("f").summary()["is_fitted"] is False


# ---- VBN._posterior_stats / infer_relative (vbn/vbn.py:483-568; tests/test_gaussian_exact_relative.py:40-57)
def _ref_posterior_stats(pdf, samples, eps=1e-12):
    weights = torch.nan_to_num(pdf, nan=0.0, posinf=0.0, neginf=0.0).clamp_min(0.0)
    denom = weights.sum(dim=1, keepdim=True)
    uniform = torch.full_like(weights, 1.0 / max(1, weights.shape[1]))
    weights = torch.where(denom > eps, weights / denom.clamp_min(eps), uniform)
    mean = (weights.unsqueeze(-1) * samples).sum(dim=1)
    var = (weights.unsqueeze(-1) * (samples - mean.unsqueeze(1)) ** 2).sum(dim=1)
    return mean, var.clamp_min(0.0).sqrt(), 1.0 / (weights**2).sum(dim=1).clamp_min(eps)


def test_posterior_stats_match_the_reference_formula(backend):
    g = torch.Generator().manual_seed(0)
    pdf = torch.rand(6, 777, generator=g)
    pdf[1, 5] = float("nan")
    pdf[2, 7] = float("inf")
    pdf[3] = 0.0  # degenerate query -> uniform weights
    pdf[4, 10] = -3.0
    samples = 3.0 + 2.0 * torch.randn(6, 777, 2, generator=g)
    model = _chain(backend.device)
    got = model._posterior_stats(pdf.to(backend.device), samples.to(backend.device))
    mean, std, ess = _ref_posterior_stats(pdf.double(), samples.double())
    torch.testing.assert_close(got["mean"].cpu().double(), mean, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(got["std"].cpu().double(), std, rtol=2e-5, atol=1e-5)
    torch.testing.assert_close(got["ess"].cpu().double(), ess, rtol=2e-5, atol=1e-4)
    with pytest.raises(ValueError):
        model._posterior_stats(pdf[0], samples)


def test_infer_relative_direction(backend):
    model = _chain(backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=4000)
    out = model.infer_relative({"target": "x2", "evidence": {"x0": torch.tensor([[2.0], [3.0]])}})
    assert out["target"] == "x2" and out["delta_mean"].shape == (2, 1)
    assert bool((out["delta_mean"] > 0).all())  # slope-1 chain: raising x0 raises x2 (reference test: delta_mean > 0)
    assert out["query_stats"]["effective_sample_size"].shape == (2,)
    with pytest.raises(ValueError):
        model.infer_relative({"target": "x2", "evidence": {}}, {"target": "x1", "evidence": {}})


# ---- categorical_table (vbn/cpds/categorical_table.py) ---------------------------
def test_categorical_table_strict_support_and_conditional(backend):
    import os

    blob = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "table.pt"),
                      weights_only=False)
    spec = blob["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=16)
    # "wet" is a categorical_table with parents (rain, sprinkler): an off-support rain value must raise
    with pytest.raises(ValueError):
        model.infer_posterior({"target": "slip", "do": {"rain": torch.tensor([[0.5]])}})
    w, s = model.infer_posterior({"target": "slip", "do": {"rain": torch.tensor([[1.0]])}})
    assert set(s.unique().tolist()) <= {0.0, 1.0}
    h = model.get_cpd("wet")
    out = h.conditional({"rain": torch.tensor([[1.0], [0.0]]), "sprinkler": torch.tensor([[0.0], [1.0]])})
    assert out["format"] == "categorical_probs" and out["k"] == 3
    probs = torch.tensor(out["probs"])
    want = torch.softmax(O.ct_logits(spec["cpds"]["wet"], torch.tensor([[1.0, 0.0], [0.0, 1.0]])), dim=-1)
    torch.testing.assert_close(probs, want.reshape(probs.shape), rtol=1e-5, atol=1e-7)


def test_table_only_model_raises_on_off_support_values(backend):
    """A model made of categorical_table CPDs only (no softmax_nn anywhere): off-support evidence or parent values
    must raise in every method and in the CPD handle, with the reference's wording (categorical_table.py:12-21).
    The device flags them and carries on with class 0, so forgetting to read the flag would return a wrong posterior."""
    import os

    full = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "table.pt"),
                      weights_only=False)["spec"]
    keep = ["season", "rain"]  # season -> rain, both categorical_table
    spec = {"nodes": keep, "topo": keep, "parents": {n: [p for p in full["parents"][n] if p in keep] for n in keep},
            "cpds": {n: full["cpds"][n] for n in keep}}
    assert {c["kind"] for c in spec["cpds"].values()} == {"categorical_table"}
    model = V.VBN.from_spec(spec, device=backend.device)
    good = {"target": "season", "evidence": {"rain": torch.tensor([[1.0], [0.0]])}}
    bad_ev = {"target": "season", "evidence": {"rain": torch.tensor([[1.0], [0.5]])}}   # evidence value off support
    bad_pa = {"target": "rain", "do": {"season": torch.tensor([[7.0]])}}                # parent value off support
    for method in ("likelihood_weighting", "importance_sampling", "monte_carlo_marginalization"):
        model.set_inference_method(method, n_samples=32)
        model.infer_posterior(good)
        for q in ((bad_ev, bad_pa) if method != "monte_carlo_marginalization" else (bad_pa,)):
            with pytest.raises(ValueError, match="outside support"):
                model.infer_posterior(q)
    h = model.get_cpd("rain")
    h.log_prob(torch.tensor([[1.0]]), {"season": torch.tensor([[2.0]])})
    with pytest.raises(ValueError, match="outside support"):
        h.log_prob(torch.tensor([[0.5]]), {"season": torch.tensor([[2.0]])})
    with pytest.raises(ValueError, match="outside support"):
        h.log_prob(torch.tensor([[1.0]]), {"season": torch.tensor([[9.0]])})


# ---- categorical_embedded_softmax (vbn/cpds/categorical_embedded_softmax.py) --------
def test_categorical_embedded_softmax_strict_support_conditional_and_frequencies(backend):
    import os

    spec = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "embedded.pt"),
                      weights_only=False)["spec"]
    model = V.VBN.from_spec(spec, device=backend.device)
    model.set_inference_method("likelihood_weighting", n_samples=16)
    with pytest.raises(ValueError):  # "y" embeds (v, u): an off-support v must raise (:36-46)
        model.infer_posterior({"target": "z", "do": {"v": torch.tensor([[0.5]])}})
    pa = torch.tensor([[1.0, 0.0], [3.0, 2.0]])  # parents of y in DAG order
    order = spec["parents"]["y"]
    out = model.get_cpd("y").conditional({order[0]: pa[:, :1], order[1]: pa[:, 1:]})
    assert out["format"] == "categorical_probs" and out["k"] == 3
    probs = torch.tensor(out["probs"])
    want = torch.softmax(O.ces_logits(spec["cpds"]["y"], pa), dim=-1)
    torch.testing.assert_close(probs, want.reshape(probs.shape), rtol=2e-5, atol=1e-7)
    # Philox path: class frequencies of a long draw follow the conditional
    cpd = V.cpd_from_spec(spec["cpds"]["y"], device=backend.device)
    s = cpd.sample(pa[:1], 30000, seed=3).cpu().reshape(-1)
    freq = torch.stack([(s == float(k)).float().mean() for k in range(3)])
    assert (freq - want[0].reshape(-1)).abs().max().item() < 0.015
    unfitted = dict(spec["cpds"]["y"], stats_ready=False)
    with pytest.raises(RuntimeError):  # :136-138
        V.cpd_from_spec(unfitted, device=backend.device).sample(pa, 4)


# ---- "lbp" wrapper (vbn/inference/lbp.py; tests/test_inference.py:27-59) -----------------------------
def test_lbp_wrapper_matches_its_definition(backend):
    model = _chain(backend.device, n=4)
    ev = {"x3": torch.tensor([[0.4], [-1.0]])}
    q = {"target": "x1", "evidence": ev}
    model.set_inference_method("importance_sampling", n_samples=300, ess_threshold=0.0)
    w0, s0 = model.infer_posterior(q, seed=12)
    model.set_inference_method("lbp", n_samples=300)
    model._inference._is.ess_threshold = 0.0
    w1, s1 = model.infer_posterior(q, seed=12)
    assert w1.shape == (2, 300) and s1.shape == (2, 300, 1) and torch.equal(s0, s1)
    # already-normalised weights are a fixed point of the damped update (lbp.py:52-63): converges at once
    torch.testing.assert_close(w1, w0, rtol=1e-5, atol=1e-9)
    with pytest.raises(ValueError):
        V.INFERENCE_REGISTRY["lbp"](damping=1.5)
    with pytest.raises(ValueError):
        V.INFERENCE_REGISTRY["lbp"](fallback="likelihood_weighting")
    model.set_inference_method("lbp", n_samples=64, fallback="monte_carlo_marginalization")
    w2, s2 = model.infer_posterior(q, seed=3)
    torch.testing.assert_close(w2.sum(1).cpu(), torch.ones(2), rtol=1e-4, atol=1e-5)


# ---- Gibbs sampler (vbn/sampling/gibbs.py; tests/test_sampling.py:46-56) -----------
def test_gibbs_sampler_philox_path(backend):
    model = _mixed(backend.device)
    model.set_sampling_method("gibbs", n_samples=7, burn_in=3)
    assert model._sampling.n_candidates == 8 and model._sampling.burn_in == 3 and model._sampling.n_steps == 1
    ev = {"n11": torch.tensor([[0.2], [0.2], [-0.4]])}
    s1 = model.sample({"target": "n5", "evidence": ev}, n_samples=7, seed=5)
    assert s1.shape == (3, 7, 1) and torch.isfinite(s1).all() and not s1.requires_grad
    assert bool((s1 == s1[:, :1]).all())  # final state repeated (views of the live state, gibbs.py:83-91)
    s2 = model.sample({"target": "n5", "evidence": ev}, n_samples=7, seed=5)
    assert torch.equal(s1, s2)
    s3 = model.sample({"target": "n5", "evidence": ev}, n_samples=7, seed=6)
    assert not torch.equal(s1, s3)
    assert s1[0, 0, 0] != s1[1, 0, 0]  # same evidence, independent chains
    # more sweeps move the chain: 1 step vs 40 steps from the same seed end in different states
    model.set_sampling_method("gibbs", n_samples=40, burn_in=0)
    s4 = model.sample({"target": "n5", "evidence": ev}, n_samples=40, seed=5)
    assert s4.shape == (3, 40, 1) and not torch.equal(s4[:, :1], s1[:, :1])
    # a fixed target is returned as given; a chain pulled by strong evidence follows it
    chain = _chain(backend.device, n=3)
    chain.set_sampling_method("gibbs", n_samples=2, burn_in=30)
    hi = chain.sample({"target": "x1", "evidence": {"x2": torch.full((256, 1), 6.0)}}, n_samples=2, seed=1)[:, 0, 0]
    lo = chain.sample({"target": "x1", "evidence": {"x2": torch.full((256, 1), -6.0)}}, n_samples=2, seed=1)[:, 0, 0]
    assert hi.mean().item() > lo.mean().item() + 2.0
    fx = chain.sample({"target": "x2", "evidence": {"x2": torch.tensor([[1.5]])}}, n_samples=2, seed=1)
    assert torch.equal(fx.cpu(), torch.full((1, 2, 1), 1.5))


# ---- resampled_importance_sampling (vbn/inference/resampled_importance_sampling.py) ----------------
def test_resampled_importance_sampling_philox_path(backend):
    spec = S.lg_chain(8)
    model = V.VBN.from_spec(spec, device=backend.device)
    n = 100_000 if backend.name == "cuda" else 6_000
    ev = torch.tensor([[0.5], [3.0]])
    model.set_inference_method("resampled_importance_sampling", n_samples=n)
    w, s = model.infer_posterior({"target": "x3", "evidence": {"x5": ev, "x7": ev + 0.3}}, seed=8)
    assert w.shape == (2, n) and s.shape == (2, n, 1)
    assert model._inference._last_resampled is True and model._inference._last_ess.shape == (2,)
    torch.testing.assert_close(w.sum(1).cpu(), torch.ones(2), rtol=1e-4, atol=1e-4)
    wd, sd = w.double().cpu(), s[..., 0].double().cpu()
    mean = (wd * sd).sum(1)
    em, evar = O.lg_exact_posterior(spec, "x3", {"x5": ev, "x7": ev + 0.3})
    # resampling adds Monte-Carlo noise of order sigma / sqrt(ESS at the resampling step) ~ sigma / sqrt(n/10)
    assert bool(((mean - em).abs() < 8.0 * evar.sqrt() / (n / 10) ** 0.5).all()), (mean, em)
    # resampling off: plain sequential importance sampling, never resamples
    w2, _ = model.infer_posterior({"target": "x3", "evidence": {"x5": ev}}, seed=8, resample=False)
    assert model._inference._last_resampled is False
    torch.testing.assert_close(w2.sum(1).cpu(), torch.ones(2), rtol=1e-4, atol=1e-4)
    # target after the last evidence node, and a fixed target
    w3, s3 = model.infer_posterior({"target": "x7", "evidence": {"x2": ev}}, seed=8)
    assert s3.shape == (2, n, 1) and torch.isfinite(s3).all()
    w4, s4 = model.infer_posterior({"target": "x2", "evidence": {"x2": ev}}, seed=8)
    assert bool((s4[0] == 0.5).all()) and bool((s4[1] == 3.0).all())


# ---- fused weight reduction / summary path (the summaries as an epilogue of the weight pass) ------
def test_summary_path_matches_posterior_stats_of_the_materialised_result(backend):
    """infer_posterior(..., summary=True): the schedule kernel accumulates the weighted moments of the target next to
    the softmax statistics (no [B, S] tensor is materialised); the result must equal VBN._posterior_stats
    (vbn/vbn.py:495-504) of the ordinary (weights, samples) result on the same draws."""
    model = _chain(backend.device, n=6)
    q = {"target": "x2", "evidence": {"x5": torch.tensor([[1.0], [2.5], [-0.5]])}}
    for method, s in (("likelihood_weighting", 1000), ("importance_sampling", 333), ("likelihood_weighting", 37)):
        model.set_inference_method(method, n_samples=s)
        if method == "importance_sampling":
            model._inference.ess_threshold = 0.0  # keep the IS pass (no fallback) so both calls share the draws
        w, x = model.infer_posterior(q, seed=21)
        ref = O_posterior_stats(w.cpu(), x.cpu())
        got = model.infer_posterior(q, seed=21, summary=True)
        assert set(got) >= {"mean", "std", "ess"} and got["mean"].shape == (3, 1) and got["ess"].shape == (3,)
        torch.testing.assert_close(got["mean"].cpu(), ref["mean"], rtol=2e-5, atol=2e-6)
        torch.testing.assert_close(got["std"].cpu(), ref["std"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(got["ess"].cpu(), ref["ess"], rtol=1e-4, atol=1e-3)


def O_posterior_stats(pdf, samples, eps=1e-12):
    weights = torch.nan_to_num(pdf, nan=0.0, posinf=0.0, neginf=0.0).clamp_min(0.0)
    denom = weights.sum(dim=1, keepdim=True)
    weights = torch.where(denom > eps, weights / denom.clamp_min(eps), torch.full_like(weights, 1.0 / weights.shape[1]))
    mean = (weights.unsqueeze(-1) * samples).sum(dim=1)
    var = (weights.unsqueeze(-1) * (samples - mean.unsqueeze(1)) ** 2).sum(dim=1)
    return {"mean": mean, "std": var.clamp_min(0.0).sqrt(), "ess": 1.0 / (weights ** 2).sum(dim=1).clamp_min(eps)}


def test_summary_path_class_histogram(backend):
    """summary={"classes": k}: the weighted class histogram of a discrete target (the benchmark adapter's
    _estimate_discrete_posterior_batch, benchmarking/models/vbn.py:202-242) from the same records."""
    spec = S.alarm_softmax(seed=0)
    model = V.VBN.from_spec(spec, device=backend.device)
    g = torch.Generator().manual_seed(2)
    ev = {n: torch.randint(0, S.ALARM[n][0], (4, 1), generator=g).float() for n in ("HRBP", "BP")}
    q = {"target": "VENTLUNG", "evidence": ev}
    model.set_inference_method("likelihood_weighting", n_samples=500)
    w, x = model.infer_posterior(q, seed=8)
    got = model.infer_posterior(q, seed=8, summary={"classes": 4})
    import numpy as np

    want = np.stack(O.estimate_discrete_posterior_batch(x.cpu()[..., 0], w.cpu(), 4))
    torch.testing.assert_close(got["probs"].cpu(), torch.as_tensor(want, dtype=torch.float32), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(got["probs"].sum(1).cpu(), torch.ones(4), rtol=1e-5, atol=1e-5)


def test_minus_inf_log_weights_follow_softmax(backend):
    """Rows whose log-weight is -inf get weight 0 when the query has any finite row (torch.softmax), whatever their
    position; a query with no finite row gets NaN weights and NaN ESS (and therefore no IS fallback), exactly like
    torch.softmax / 1 / sum w^2 in the reference (importance_sampling.py:82-88)."""
    from vectorizedbayesiannetwork_b200 import engine as E

    dev = backend.device
    logw = torch.randn(3, 70)
    logw[0, 0] = float("-inf")          # first element of a row
    logw[0, 40:45] = float("-inf")
    logw[1, :] = float("-inf")          # nothing finite
    lw = logw.to(dev)
    stats = E.lse_stats(lw)
    w, ess = E.normalize_weights(lw, stats)
    want = torch.softmax(logw, dim=1)
    torch.testing.assert_close(w.cpu()[[0, 2]], want[[0, 2]], rtol=2e-5, atol=1e-9)
    assert bool((w.cpu()[0, 40:45] == 0).all()) and w.cpu()[0, 0] == 0
    assert bool(torch.isnan(w.cpu()[1]).all()) and bool(torch.isnan(ess.cpu()[1]))
    torch.testing.assert_close(ess.cpu()[[0, 2]], 1.0 / (want[[0, 2]] ** 2).sum(1), rtol=1e-4, atol=1e-4)


def test_barren_node_pruning_is_exact(backend):
    """prune=True drops unobserved nodes without an observed / queried descendant from the schedule; the emitted
    nodes keep their place in the random streams, so weights and samples equal the full walk's BIT FOR BIT."""
    cases = [(S.random_dag_lg_mdn(60, seed=4), "n30", ["n57", "n58", "n59"], None),
             (S.alarm_softmax(seed=0), "LVFAILURE", ["HRBP", "BP"], S.ALARM)]
    for spec, target, ev_nodes, cards in cases:
        model = V.VBN.from_spec(spec, device=backend.device)
        g = torch.Generator().manual_seed(6)
        ev = {n: (torch.randint(0, cards[n][0], (3, 1), generator=g).float() if cards else 0.3 * torch.randn(3, 1, generator=g))
              for n in ev_nodes}
        q = {"target": target, "evidence": ev}
        for method in ("likelihood_weighting", "importance_sampling", "monte_carlo_marginalization"):
            model.set_inference_method(method, n_samples=257)
            w0, s0 = model.infer_posterior(q, seed=5)
            w1, s1 = model.infer_posterior(q, seed=5, prune=True)
            assert torch.equal(s0, s1) and torch.equal(w0, w1), (target, method)
            if method != "monte_carlo_marginalization":  # (its one-CPD fast path has nothing to prune)
                runner = model._inference._runner
                n_ops = sorted(len(p.program.ops) for p in runner._cache.values())
                assert n_ops[0] < n_ops[-1], n_ops  # the pruned schedule really is shorter
