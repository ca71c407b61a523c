"""The Philox-only fast paths, pinned to the oracle on IDENTICAL draws.

Reference-recorded noise cannot reach the kernels' fast paths (plain linear-Gaussian / MDN / table ops, op-embedded
root mixtures, the tcgen05 MLP with its register-resident mixture tail): they only exist where the draws come from
the device generator.  So the test turns the direction round: it runs the product, asks the library for the draws
that run made (``vbn_stream_draws``: same generator code, same counters), and replays them through the CPU oracle's
restatement of the reference algorithm (``oracle.device_stream.DeviceStreamNoise``).  Samples must agree within 1e-5
relative, weights within 5e-5 -- no mismatch fractions.  The only rows left out are those where a categorical pick's
uniform lies within 2e-6 of a CDF edge (the two sides may legitimately round the edge differently, which changes the
row completely); they are counted and bounded.

Separate known-answer tests pin the generator itself: raw words against the oracle's numpy Philox4x32-10 (which is
checked against the Random123 vectors in test_host_side.py), the uniform map bit-for-bit, Box-Muller within the
approximation error of the device intrinsics, and the law of the normals (moments, tails) over >= 1e8 draws.
"""
import math

import numpy as np
import pytest
import torch

import vectorizedbayesiannetwork_b200 as V
from backends import backend  # noqa: F401  (fixture)
from oracle import philox as P
from oracle import vbn_oracle as O
from oracle.device_stream import DeviceStreamNoise
from vectorizedbayesiannetwork_b200 import _lib as L
from vectorizedbayesiannetwork_b200 import engine as E
from vectorizedbayesiannetwork_b200 import synthetic as S
from vectorizedbayesiannetwork_b200.core import model_cpds

SEED = 0x1234_5678_9ABC_DEF0


# ----------------------------------------------------------------------------------------------------------
# generator known answers
# ----------------------------------------------------------------------------------------------------------
def _counters(n_blocks, b, s, *, tag, shared, q_off=0, s_off=0, call=0):
    blk, bb, ss = np.meshgrid(np.arange(n_blocks), np.arange(1 if shared else b), np.arange(s), indexing="ij")
    ctr = np.stack([s_off + ss, np.full_like(bb, 0xFFFFFFFF) if shared else q_off + bb,
                    blk | (tag << 30), np.full_like(bb, call)], axis=-1).astype(np.uint32)
    return ctr.reshape(-1, 4)


@pytest.mark.parametrize("shared", [False, True])
def test_device_generator_words_match_oracle_philox(backend, shared):
    b, s, nb = 3, 37, 5
    kw = dict(seed=SEED, shared=shared, call_offset=7, query_offset=11, sample_offset=1000, device=backend.device)
    for kind, tag in (("bits", 0), ("uniform", 1)):
        got = E.stream_draws(kind, 4 * nb, b, s, **kw).cpu().numpy()
        tag_full = tag + (2 if shared else 0)
        ctr = _counters(nb, b, s, tag=tag_full, shared=shared, q_off=11, s_off=1000, call=7)
        want = P.philox4x32_10(ctr, SEED & 0xFFFFFFFF, SEED >> 32).reshape(nb, -1, s, 4)  # [blk, b, s, word]
        want = np.moveaxis(want, -1, 1).reshape(4 * nb, -1, s)                              # [4 blk + word, b, s]
        if kind == "bits":
            np.testing.assert_array_equal(got.view(np.uint32), want)
        else:
            np.testing.assert_array_equal(got, P.u01(want))  # (x >> 8 + 0.5) * 2^-24, exact in fp32


def test_device_normals_are_box_muller_of_the_same_words(backend):
    b, s, nb = 2, 4096, 4
    kw = dict(seed=SEED, call_offset=0, device=backend.device)
    bits = E.stream_draws("bits", 4 * nb, b, s, **kw).cpu().numpy().view(np.uint32).reshape(nb, 4, b, s)
    z = E.stream_draws("normal", 4 * nb, b, s, **kw).cpu().numpy().reshape(nb, 4, b, s)

    def bm(a, c):  # csrc/vbn_device.cuh box_muller, evaluated in float64
        f1 = ((a >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32).astype(np.float64)
        f2 = ((c >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32).astype(np.float64)
        u1 = (f1.astype(np.float32) - np.float32(0.99999994)).astype(np.float64)
        ang = 2 * math.pi * (f2 - 1.5)
        r = np.sqrt(-2.0 * np.log(u1))
        return r * np.cos(ang), r * np.sin(ang)

    if backend.name == "emu":  # the host build keeps the textbook form on u01(a), u01(b) - 0.5
        u1 = P.u01(bits[:, 0]).astype(np.float64), P.u01(bits[:, 2]).astype(np.float64)
        u2 = P.u01(bits[:, 1]).astype(np.float64) - 0.5, P.u01(bits[:, 3]).astype(np.float64) - 0.5
        want = []
        for a, c in zip(u1, u2):
            r = np.sqrt(-2.0 * np.log(a))
            want += [r * np.cos(2 * math.pi * c), r * np.sin(2 * math.pi * c)]
    else:
        w0, w1 = bm(bits[:, 0], bits[:, 1])
        w2, w3 = bm(bits[:, 2], bits[:, 3])
        want = [w0, w1, w2, w3]
    want = np.stack(want, axis=1)
    # lg2 / sqrt / sin / cos .approx against float64: the angle functions carry ~1e-6 absolute error, scaled by the
    # radius (<= 5.8); measured maximum on the B200 7e-6
    np.testing.assert_allclose(z, want, rtol=2e-6, atol=1.5e-5)


@pytest.mark.gpu
def test_device_normal_law_moments_and_tails():
    """>= 1e8 device normals: mean, variance, kurtosis and tail mass of N(0,1).  The generator floors u1 at 2^-24
    (|z| <= 5.77): P(|z| > 5.77) = 8e-9 is the only part of the law that is cut."""
    dev = torch.device("cuda", 0)
    L._lib = None
    n = 0
    s1 = s2 = s4 = 0.0
    tail4 = tail5 = 0
    zmax = 0.0
    for call in range(4):
        z = E.stream_draws("normal", 64, 64, 8192, seed=SEED + call, call_offset=call, device=dev).double().flatten()
        n += z.numel()
        s1 += z.sum().item()
        s2 += (z * z).sum().item()
        s4 += (z ** 4).sum().item()
        tail4 += int((z.abs() > 4).sum())
        tail5 += int((z.abs() > 5).sum())
        zmax = max(zmax, float(z.abs().max()))
    assert n >= 1e8
    mean, var, kurt = s1 / n, s2 / n, s4 / n
    assert abs(mean) < 5 / math.sqrt(n)                      # 5 sigma
    assert abs(var - 1) < 5 * math.sqrt(2 / n)
    assert abs(kurt - 3) < 5 * math.sqrt(96 / n)
    p4, p5 = 6.334e-5, 5.733e-7                              # P(|z| > 4), P(|z| > 5)
    assert abs(tail4 - n * p4) < 5 * math.sqrt(n * p4)
    assert abs(tail5 - n * p5) < 5 * math.sqrt(n * p5) + 1
    assert zmax <= 5.78


# ----------------------------------------------------------------------------------------------------------
# replay of a run's draws through the oracle
# ----------------------------------------------------------------------------------------------------------
def _scope(model, plan, spec, seed, b, s, device):
    """DeviceStreamNoise scope for one pass: stream table of the compiled schedule + the draws of that pass."""
    prog = plan.program
    cpds = model_cpds(model)
    streams, n_norm, n_uni, any_shared = {}, 1, 1, False
    for op, node in zip(prog.ops, prog.nodes):
        if int(op["kind"]) == L.OP_NONE or (int(op["flags"]) & 3) != L.SRC_SAMPLE:
            continue
        pk = cpds[node].pack()
        d = int(op["dim"])
        shared = bool(int(op["flags"]) & L.F_SHARED)
        any_shared |= shared
        c = spec["cpds"][node]
        wb_u = None
        if c["kind"] == "softmax_nn" and c.get("within_bin", "uniform") != "gaussian" and pk.n_uniforms == 2 * d:
            wb_u = int(op["u_off"]) + d
        streams[node] = {"n_off": int(op["n_off"]), "u_off": int(op["u_off"]), "shared": shared, "wb_u": wb_u}
        n_norm = max(n_norm, int(op["n_off"]) + pk.n_normals)
        n_uni = max(n_uni, int(op["u_off"]) + pk.n_uniforms)
    kw = dict(seed=seed, device=device)
    sc = {"streams": streams, "B": b, "S": s,
          "normals": E.stream_draws("normal", n_norm, b, s, **kw).cpu(),
          "uniforms": E.stream_draws("uniform", n_uni, b, s, **kw).cpu(),
          "normals_shared": E.stream_draws("normal", n_norm, b, s, shared=True, **kw).cpu() if any_shared else None,
          "uniforms_shared": E.stream_draws("uniform", n_uni, b, s, shared=True, **kw).cpu() if any_shared else None}
    return sc


def _plan_of(method):
    return next(iter(method._runner._cache.values()))


def _compare(w, smp, ow, os_, suspect, max_suspect_frac, what):
    ok = ~suspect
    frac = suspect.float().mean().item()
    assert frac <= max_suspect_frac, f"{what}: {frac:.2e} of the rows sit on a CDF edge"
    torch.testing.assert_close(smp[ok], os_[ok], rtol=1e-5, atol=1e-6)
    if not suspect.any():
        torch.testing.assert_close(w, ow, rtol=5e-5, atol=1e-9)
    else:  # weights are normalised per query: compare the un-normalised ratios on the clean rows of clean queries
        clean_q = ~suspect.any(dim=1)
        torch.testing.assert_close(w[clean_q], ow[clean_q], rtol=5e-5, atol=1e-9)
    return frac


def _run_lw(spec, q, s, device, seed=SEED):
    model = V.VBN.from_spec(spec, device=device)
    model.set_inference_method("likelihood_weighting", n_samples=s)
    w, smp = model.infer_posterior(q, seed=seed)
    b = next(iter(q["evidence"].values())).shape[0]
    sc = _scope(model, _plan_of(model._inference), spec, seed, b, s, device)
    noise = DeviceStreamNoise({"lw": sc})
    ow, os_ = O.likelihood_weighting(spec, q, s, noise=noise)
    return w.cpu(), smp.cpu(), ow, os_, noise.suspect["lw"], _plan_of(model._inference).program


def _run_is(spec, q, s, device, seed=SEED):
    model = V.VBN.from_spec(spec, device=device)
    model.set_inference_method("importance_sampling", n_samples=s)
    w, smp = model.infer_posterior(q, seed=seed)
    inf = model._inference
    b = next(iter(q["evidence"].values())).shape[0]
    scopes = {"is": _scope(model, _plan_of(inf), spec, seed, b, s, device)}
    if inf._last_fallback:  # the whole batch was redone by likelihood weighting with seed + 1 (inference.py)
        scopes["lw"] = _scope(model, _plan_of(inf._lw), spec, seed + 1, b, s, device)
    noise = DeviceStreamNoise(scopes)
    ow, os_, info = O.importance_sampling(spec, q, s, noise=noise, return_info=True)
    assert bool(info["fallback"]) == bool(inf._last_fallback)
    clean_q = ~noise.suspect["is"].any(dim=1)
    torch.testing.assert_close(inf._last_ess.cpu()[clean_q], info["ess"][clean_q], rtol=1e-4, atol=1e-3)
    suspect = noise.suspect["lw"] if inf._last_fallback else noise.suspect["is"]
    return w.cpu(), smp.cpu(), ow, os_, suspect, _plan_of(inf).program, bool(inf._last_fallback)


def _flag_count(prog, flag):
    return int(((prog.ops["flags"] & flag) != 0).sum())


def test_lg_chain_plain_ops_replay(backend):
    """cfg2 shape: 50-node linear-Gaussian chain, importance sampling (VBN_F_LGPLAIN on every drawn node)."""
    spec = S.lg_chain(50)
    g = torch.Generator().manual_seed(3)
    q = {"target": "x25", "evidence": {"x49": 4.9 + 3.6 * (torch.rand(6, 1, generator=g) - 0.5)}}
    w, smp, ow, os_, suspect, prog, fb = _run_is(spec, q, 2048, backend.device)
    assert _flag_count(prog, L.F_LGPLAIN) >= 48
    assert not suspect.any()  # no categorical draw anywhere
    _compare(w, smp, ow, os_, suspect, 0.0, "lg chain")


def test_alarm_plain_table_ops_replay(backend):
    """cfg3 shape: ALARM with softmax_nn CPDs compiled to lookup tables, likelihood weighting (VBN_F_TABPLAIN)."""
    spec = S.alarm_softmax(seed=0)
    g = torch.Generator().manual_seed(4)
    ev = {n: torch.randint(0, S.ALARM[n][0], (5, 1), generator=g).float() for n in ("HRBP", "BP", "EXPCO2", "PRESS")}
    q = {"target": "LVFAILURE", "evidence": ev}
    w, smp, ow, os_, suspect, prog = _run_lw(spec, q, 4096, backend.device)
    assert _flag_count(prog, L.F_TABPLAIN) >= 25
    frac = _compare(w, smp, ow, os_, suspect, 2e-3, "alarm")  # 33 picks per row, each within 2e-6 of <= 3 edges
    # discrete target: the un-suspect rows agree EXACTLY
    assert torch.equal(smp[~suspect], os_[~suspect]), frac


@pytest.mark.parametrize("n_nodes,b,s", [(60, 4, 1024), (200, 3, 512)])
def test_cfg5_shaped_dag_replay(backend, n_nodes, b, s):
    """cfg5 shape: random DAG, even nodes linear_gaussian, odd nodes mdn(K=3, [32,32]); importance sampling.  On the
    GPU this is the tcgen05 kernel with every fast path on (LGPLAIN + MDNPLAIN + MDNROOT, first layer on the FP32
    pipe, descriptor tails); in the host emulation the FFMA kernel's LGPLAIN + MDNROOT paths."""
    spec = S.random_dag_lg_mdn(n_nodes, seed=2)
    g = torch.Generator().manual_seed(5)
    q = {"target": f"n{n_nodes // 2}",
         "evidence": {n: 0.3 * torch.randn(b, 1, generator=g) for n in spec["nodes"][-3:]}}
    w, smp, ow, os_, suspect, prog, fb = _run_is(spec, q, s, backend.device)
    assert _flag_count(prog, L.F_LGPLAIN) > n_nodes // 4 and _flag_count(prog, L.F_MDNROOT) > 0
    if backend.name == "cuda":
        assert prog.tc and _flag_count(prog, L.F_MDNPLAIN) > n_nodes // 4
    # ~n/2 mixture picks per row, each within 2e-6 of one of two CDF edges
    _compare(w, smp, ow, os_, suspect, 4 * n_nodes * 2e-6 + 2.0 / (b * s), f"dag{n_nodes}")


@pytest.mark.gpu
def test_cfg5_full_dag_replay_on_the_tensor_core_kernel():
    """The BASELINE cfg5 model itself (1000 nodes) on a few queries: every sample of every clean row within 1e-5 of
    the oracle on the same draws, for both tile geometries of the tcgen05 kernel."""
    import os

    dev = torch.device("cuda", 0)
    L._lib = None
    spec = S.random_dag_lg_mdn(1000, seed=0)
    g = torch.Generator().manual_seed(1)
    b, s = 2, 256
    q = {"target": "n500", "evidence": {n: 0.3 * torch.randn(b, 1, generator=g) for n in spec["nodes"][-5:]}}
    old = os.environ.get("VBN_TC_SHAPE")
    try:
        for shape in ("4x1", "2x2"):
            os.environ["VBN_TC_SHAPE"] = shape
            w, smp, ow, os_, suspect, prog, fb = _run_is(spec, q, s, dev)
            assert prog.tc and _flag_count(prog, L.F_MDNPLAIN) > 300
            _compare(w, smp, ow, os_, suspect, 0.02, f"cfg5 {shape}")
    finally:
        if old is None:
            os.environ.pop("VBN_TC_SHAPE", None)
        else:
            os.environ["VBN_TC_SHAPE"] = old
