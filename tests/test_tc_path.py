"""The tcgen05 (tensor-core) schedule kernel against the FFMA schedule kernel on identical Philox
streams: same schedule, same draws, so every output must agree to fp32 parity tolerance (the
3xTF32 split keeps ~2^-21 per product).  The FFMA kernel itself is pinned to the reference by
tests/test_parity_golden.py (which also runs through the tensor-core kernel on the GPU, because the
golden models use the reference's default hidden_dims [32, 32])."""
import numpy as np
import pytest
import torch

import vectorizedbayesiannetwork_b200 as V
from vectorizedbayesiannetwork_b200 import _lib as L
from vectorizedbayesiannetwork_b200 import synthetic as S
from vectorizedbayesiannetwork_b200.cpds import _core_matrix_image, _split_tf32, pack_mlp_tc
from vectorizedbayesiannetwork_b200.plan import Role, compile_schedule


def test_tf32_split_is_exact_and_representable():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(4096) * np.exp(rng.uniform(-20, 20, 4096))).astype(np.float32)
    hi, lo = _split_tf32(w)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)
    # hi is the nearest tf32; the two-term split captures w to ~2^-22
    assert np.all(np.abs(w - hi) <= np.abs(w) * 2.0**-11 * 1.0001)
    assert np.all(np.abs(w - (hi.astype(np.float64) + lo)) <= np.abs(w) * 2.0**-21)


def test_core_matrix_image_layout():
    n, k = 16, 8
    w = np.arange(n * k, dtype=np.float32).reshape(n, k)
    img = _core_matrix_image(w)
    for r in range(n):
        for c in range(k):
            off = ((r // 8) * (k // 4) + c // 4) * 32 + (r % 8) * 4 + c % 4  # floats
            assert img[off] == w[r, c]


def test_pack_mlp_tc_eligibility_and_size():
    g = torch.Generator().manual_seed(0)
    ok = pack_mlp_tc(S.mlp_layers(g, 3, (32, 32), 9), 3)
    assert ok is not None
    blob, k1, n3 = ok
    # every layer's K is extended by one 8-wide step that carries the bias (column K, then zeros)
    assert (k1, n3) == (8, 16) and blob.size * 4 == 4 * 2 * (32 * (8 + 8) + 32 * (32 + 8) + 16 * (32 + 8))
    layers = S.mlp_layers(torch.Generator().manual_seed(0), 3, (32, 32), 9)
    w1hi = blob[: 32 * 16].reshape(4, 4, 8, 4).transpose(0, 2, 1, 3).reshape(32, 16)  # undo the core-matrix image
    w1lo = blob[32 * 16: 2 * 32 * 16].reshape(4, 4, 8, 4).transpose(0, 2, 1, 3).reshape(32, 16)
    torch.testing.assert_close(torch.from_numpy(w1hi + w1lo)[:, :3], layers[0][0], rtol=0, atol=1e-6)  # lo is itself rounded to tf32: 2^-21 relative
    torch.testing.assert_close(torch.from_numpy(w1hi + w1lo)[:, 8], layers[0][1], rtol=0, atol=1e-6)
    assert not (w1hi + w1lo)[:, 3:8].any() and not (w1hi + w1lo)[:, 9:].any()
    # gaussian_nn / mdn nodes with <= 4 parent dims: first layer on the FP32 pipe -> plain W1^T[4][32], b1[32] block
    blob2, k1b, n3b = pack_mlp_tc(layers, 3, l1_fma=True)
    assert (k1b, n3b) == (0, 16) and blob2.size * 4 == 640 + 4 * 2 * (32 * 40 + 16 * 40)
    plain = blob2[:160].reshape(5, 32)
    torch.testing.assert_close(torch.from_numpy(plain[:3]), layers[0][0].t(), rtol=0, atol=0)
    torch.testing.assert_close(torch.from_numpy(plain[4]), layers[0][1], rtol=0, atol=0)
    assert not plain[3].any()
    np.testing.assert_array_equal(blob2[160:], blob[2 * 32 * 16:])       # the W2 / W3 images are the same
    assert pack_mlp_tc(S.mlp_layers(g, 5, (32, 32), 9), 5, l1_fma=True)[1] == 8   # > 4 parent dims: MMA first layer
    assert pack_mlp_tc(S.mlp_layers(g, 3, (16, 16), 9), 3) is None       # other hidden sizes: FFMA path
    assert pack_mlp_tc(S.mlp_layers(g, 40, (32, 32), 9), 40) is None     # too many parent dims
    assert pack_mlp_tc(S.mlp_layers(g, 3, (32, 32), 40), 3) is None      # too many outputs


def test_plan_marks_tensor_core_ops(monkeypatch):
    spec = S.random_dag_lg_mdn(12, seed=3)
    cpds = {n: V.cpd_from_spec(c, device="cpu") for n, c in spec["cpds"].items()}
    roles = {n: Role() for n in spec["topo"]}
    prog = compile_schedule(spec["topo"], spec["parents"], cpds, roles, use_tc=True)
    assert prog.tc
    for op, n in zip(prog.ops, prog.nodes):
        want = spec["cpds"][n]["kind"] == "mdn" and spec["cpds"][n]["input_dim"] > 0
        assert bool(op["tc"][0]) == want
        if want:
            assert op["tc"][2] == 0 and op["tc"][3] == 16  # K1 = 0: first layer on the FP32 pipe
            if op["flags"] & L.F_MDNFAST:  # short descriptor: out_slot | tail count << 16 ride in tc[1]
                assert op["tc"][1] & 0xFFFF == op["out_slot"] and op["tc"][1] >> 16 == op["layer_dim"][7]
            else:
                assert op["tc"][1] % 4 == 0
    assert all(off % 4 == 0 for off, _ in prog.tc_list)
    assert any(op["flags"] & L.F_MDNFAST for op in prog.ops)
    assert not compile_schedule(spec["topo"], spec["parents"], cpds, roles, use_tc=False).tc


def _run(spec, q, method, n_samples, seed, device):
    model = V.VBN.from_spec(spec, device=device)
    model.set_inference_method(method, n_samples=n_samples)
    w, s = model.infer_posterior(q, seed=seed)
    plan = next(iter(model._inference._runner._cache.values()))
    torch.cuda.synchronize()
    return w.cpu(), s.cpu(), plan.program.tc


@pytest.mark.gpu
@pytest.mark.parametrize("n_queries,n_samples", [(1, 200), (8, 2048), (3, 1000)])
def test_tensor_core_kernel_matches_ffma_kernel(monkeypatch, n_queries, n_samples):
    dev = torch.device("cuda", 0)
    spec = S.random_dag_lg_mdn(60, seed=1)
    g = torch.Generator().manual_seed(5)
    q = {"target": "n30", "evidence": {n: 0.3 * torch.randn(n_queries, 1, generator=g) for n in spec["nodes"][-3:]}}
    monkeypatch.setenv("VBN_TC", "0")
    w0, s0, tc0 = _run(spec, q, "likelihood_weighting", n_samples, 1234, dev)
    monkeypatch.setenv("VBN_TC", "1")
    w1, s1, tc1 = _run(spec, q, "likelihood_weighting", n_samples, 1234, dev)
    assert not tc0 and tc1
    assert torch.isfinite(s1).all() and torch.isfinite(w1).all()
    # a component pick can flip when u sits within rounding of a CDF edge: allow a vanishing fraction
    bad = ((s1 - s0).abs() > 1e-5 + 1e-5 * s0.abs()).float().mean().item()
    assert bad < 1e-4, f"fraction of mismatching samples {bad}"
    torch.testing.assert_close(w1.sum(1), torch.ones(n_queries), rtol=1e-4, atol=1e-4)
    badw = ((w1 - w0).abs() > 1e-7 + 2e-4 * w0.abs()).float().mean().item()
    assert badw < 1e-3, f"fraction of mismatching weights {badw}"


@pytest.mark.gpu
def test_long_runs_between_mlp_ops_leave_the_descriptor_tail(monkeypatch):
    """MLP ops followed by more ops than a ring slot can carry get no descriptor tail (plan.py: all or nothing): the
    walk's running descriptor pointer then continues in the global op list and returns to a ring slot at the next
    MLP op with a tail (TcMlp::layers_end, both branches)."""
    import random
    dev = torch.device("cuda", 0)
    rng, gen = random.Random(7), torch.Generator().manual_seed(7)
    names = [f"n{i}" for i in range(400)]
    parents, cpds = {}, {}
    for i, name in enumerate(names):
        ps = sorted(rng.sample(range(max(0, i - 20), i), min(rng.randint(1, 3), i))) if i else []
        parents[name] = [names[p] for p in ps]
        cpds[name] = S.mdn_cpd(gen, len(ps)) if i in (5, 200, 390) else S.lg_cpd(gen, len(ps))
    spec = {"nodes": names, "parents": parents, "topo": list(names), "cpds": cpds}
    g = torch.Generator().manual_seed(5)
    q = {"target": "n395", "evidence": {n: 0.3 * torch.randn(4, 1, generator=g) for n in names[-2:]}}
    monkeypatch.setenv("VBN_TC", "0")
    w0, s0, tc0 = _run(spec, q, "likelihood_weighting", 1024, 99, dev)
    monkeypatch.setenv("VBN_TC", "1")
    model = V.VBN.from_spec(spec, device=dev)
    model.set_inference_method("likelihood_weighting", n_samples=1024)
    w1, s1 = model.infer_posterior(q, seed=99)
    prog = next(iter(model._inference._runner._cache.values())).program
    tails = [int(op["layer_dim"][7]) for op in prog.ops if op["tc"][0]]
    assert prog.tc and not tc0 and tails[:2] == [0, 0] and tails[2] > 0, tails
    w1, s1 = w1.cpu(), s1.cpu()
    assert torch.isfinite(s1).all() and torch.isfinite(w1).all()
    bad = ((s1 - s0).abs() > 1e-5 + 1e-5 * s0.abs()).float().mean().item()
    assert bad < 1e-4, f"fraction of mismatching samples {bad}"
    badw = ((w1 - w0).abs() > 1e-7 + 2e-4 * w0.abs()).float().mean().item()
    assert badw < 1e-3, f"fraction of mismatching weights {badw}"
