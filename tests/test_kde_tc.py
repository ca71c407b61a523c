"""KDE conditional log-density with the pairwise |x - y|^2 term on the tensor cores (csrc/vbn_k_kde_tc.cu,
vbn_kde_log_prob_tc) against the oracle's restatement of vbn/cpds/kde.py:105-149.  The GEMM form
|x|^2 + |y|^2 - 2 x.y cancels, so the kernel splits every operand into three tf32 pieces (fp32-exact products); the
bar is the same 1e-5 as everywhere else."""
import pytest
import torch

import vectorizedbayesiannetwork_b200 as V
from oracle import vbn_oracle as O


def _kde(n, dp, dx, bw=0.5, pbw=0.6, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n, dp, generator=g) if dp else None
    y = (torch.sin(p.sum(1, keepdim=True)) if dp else 0.0) + 0.4 * torch.randn(n, dx, generator=g)
    return {"kind": "kde", "input_dim": dp, "output_dim": dx, "bandwidth": bw, "parent_bandwidth": pbw,
            "min_scale": 1e-4, "parents": p, "targets": y}


def _check(c, rows, monkeypatch, seed=1, spread=1.0, rtol=1e-5, atol=1e-5):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    x = spread * torch.randn(rows, c["output_dim"], generator=g)
    p = spread * torch.randn(rows, c["input_dim"], generator=g) if c["input_dim"] else None
    cpd = V.cpd_from_spec(c, device=dev)
    monkeypatch.setenv("VBN_KDE_TC", "1")
    got = cpd.log_prob(x, p).cpu().reshape(-1)
    want = O.kde_log_prob(c, x, p).reshape(-1)
    assert torch.isfinite(got).all()
    torch.testing.assert_close(got, want, rtol=rtol, atol=atol)
    return cpd, x, p, got


@pytest.mark.gpu
def test_kde_tensor_core_matches_oracle_dp7_dx1_50k_points(monkeypatch):
    _check(_kde(50_000, 7, 1), 4096, monkeypatch)


@pytest.mark.gpu
@pytest.mark.parametrize("n,dp,dx", [(1000, 7, 1), (777, 12, 2), (4097, 0, 8), (65, 30, 6), (64, 3, 1)])
def test_kde_tensor_core_shapes(monkeypatch, n, dp, dx):
    """ragged tile tails (N not a multiple of 64), root KDE (no parents), the widest supported dims, a row count that
    is not a multiple of the 128-row CTA tile"""
    _check(_kde(n, dp, dx), 4099, monkeypatch)


@pytest.mark.gpu
def test_kde_tensor_core_far_queries_take_the_exact_pass(monkeypatch):
    """Queries many kernel widths from every stored point: the GEMM form is not trusted there (scaled squared norm
    above the kernel's bound) and the fixed-shift sums underflow, so these rows are redone with direct differences and
    the online-max accumulation.  The density is then a difference of two logsumexps of size ~1e3-1e4, which fp32
    itself only resolves to ~1e-7 of that size -- the oracle's own fp32 evaluation included -- so the bar here is a few
    fp32 ulps of the logsumexp magnitude, measured against a float64 evaluation."""
    c = _kde(2000, 7, 1, bw=0.05, pbw=0.05)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    x, p = 3.0 * torch.randn(4096, 1, generator=g), 3.0 * torch.randn(4096, 7, generator=g)
    cpd = V.cpd_from_spec(c, device=dev)
    monkeypatch.setenv("VBN_KDE_TC", "1")
    got = cpd.log_prob(x, p).cpu().reshape(-1).double()
    sy, sp = 0.05 + 1e-4, 0.05 + 1e-4
    dpar = ((p.double()[:, None, :] - c["parents"].double()[None]) / sp) ** 2
    dy = ((x.double()[:, None, :] - c["targets"].double()[None]) / sy) ** 2
    lkp = -0.5 * dpar.sum(-1)
    lky = -0.5 * dy.sum(-1) - 0.5 * (torch.log(torch.tensor(2 * torch.pi, dtype=torch.float64)) + 2 * torch.log(torch.tensor(sy, dtype=torch.float64)))
    num, den = torch.logsumexp(lkp + lky, dim=1), torch.logsumexp(lkp, dim=1)
    exact = num - den
    ulp = 1.2e-7 * (num.abs() + den.abs())
    assert torch.isfinite(got).all()
    assert bool(((got - exact).abs() <= 8 * ulp + 2e-5).all()), ((got - exact).abs() / ulp).max()
    oracle = O.kde_log_prob(c, x, p).reshape(-1).double()   # the reference's own fp32 arithmetic: same envelope
    assert bool(((oracle - exact).abs() <= 8 * ulp + 2e-5).all())


@pytest.mark.gpu
def test_kde_tensor_core_equals_fp32_pipe_kernel(monkeypatch):
    c = _kde(20_000, 3, 1)
    cpd, x, p, got_tc = _check(c, 8192, monkeypatch)
    monkeypatch.setenv("VBN_KDE_TC", "0")
    got_fp = cpd.log_prob(x, p).cpu().reshape(-1)
    torch.testing.assert_close(got_tc, got_fp, rtol=1e-5, atol=1e-5)
