"""The reference arm of bench.py: the UNMODIFIED reference package (baseline/_ref or /root/reference) loaded with the
parameters of a model spec through its own CPD classes (oracle/reference_arm.py).  Checked here: the round trip
spec -> reference modules -> spec is lossless, and the reference's own infer_posterior / log_prob on those modules
equals the oracle on the round-tripped spec bit for bit under the same torch seed (so the reference arm, the oracle
and -- through the parity tests -- the B200 arm all evaluate the same model)."""
import pytest
import torch

from oracle import reference_arm as R
from oracle import vbn_oracle as O
from vectorizedbayesiannetwork_b200 import synthetic as S

pytestmark = pytest.mark.skipif(R.reference_path() is None, reason="reference package not present")


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return torch.equal(a.float(), torch.as_tensor(b).float().reshape(a.shape))
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, float):
        return abs(a - float(b)) <= 1e-12 * max(1.0, abs(a))
    return a == b


@pytest.mark.parametrize("name", ["cfg5", "cfg2", "cfg3", "cfg4"])
def test_spec_round_trips_through_reference_modules(name):
    spec = {"cfg5": lambda: S.random_dag_lg_mdn(40, seed=3), "cfg2": lambda: S.lg_chain(6),
            "cfg3": lambda: S.alarm_softmax(seed=0), "cfg4": lambda: S.kde_pair(300)}[name]()
    model = R.reference_from_spec(spec)
    back = O.spec_from_reference(model)
    assert set(back["topo"]) == set(spec["topo"])
    for node in spec["topo"]:
        assert list(back["parents"][node]) == list(spec["parents"][node])
        for key, val in spec["cpds"][node].items():
            assert _same(val, back["cpds"][node][key]), (node, key)


def test_reference_inference_on_loaded_modules_equals_oracle():
    spec = S.random_dag_lg_mdn(40, seed=3)
    model = R.reference_from_spec(spec)
    back = O.spec_from_reference(model)
    g = torch.Generator().manual_seed(0)
    q = {"target": "n20", "evidence": {n: 0.3 * torch.randn(3, 1, generator=g) for n in spec["nodes"][-2:]}}
    for method, fn in (("importance_sampling", O.importance_sampling), ("likelihood_weighting", O.likelihood_weighting)):
        model.set_inference_method(method, n_samples=64)
        torch.manual_seed(11)
        pdf, smp = model.infer_posterior(q)
        torch.manual_seed(11)
        ow, os_ = fn(back, q, 64)
        assert torch.equal(pdf, ow) and torch.equal(smp, os_), method


def test_reference_kde_log_prob_on_loaded_modules_equals_oracle():
    spec = S.kde_pair(500)
    model = R.reference_from_spec(spec)
    g = torch.Generator().manual_seed(1)
    x, p = torch.randn(64, 1, generator=g), torch.randn(64, 1, generator=g)
    got = model.get_cpd("y").log_prob(x, {"p": p})
    want = O.kde_log_prob(spec["cpds"]["y"], x, p)
    assert torch.equal(got.reshape(-1), want.reshape(-1))
