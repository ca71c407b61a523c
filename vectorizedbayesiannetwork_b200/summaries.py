"""Posterior summaries computed right after the inference path: the helpers the reference's
benchmark adapter runs on the host -- a Python double loop over (query, sample) for discrete targets
(benchmarking/models/vbn.py:202-242) and numpy moments for continuous ones (:381-423).  Here they
are one kernel launch over the [B,S] tensors already resident on the GPU; only the [B,k] / [B,2] result
crosses to the host.  Same names, arguments and return formats as the reference functions."""
from __future__ import annotations

import math
from typing import Any, List, Optional

import torch

from . import engine as E


def _normalize_probs(probs) -> List[float]:
    """benchmarking/models/vbn.py:116-121."""
    arr = torch.as_tensor(list(probs), dtype=torch.float64)
    total = float(arr.sum())
    if not math.isfinite(total) or total <= 0:
        return (torch.ones_like(arr) / arr.numel()).tolist()
    return (arr / total).tolist()


def estimate_discrete_posterior_tensor(samples: torch.Tensor, weights: torch.Tensor, k: int) -> torch.Tensor:
    """[B,k] class probabilities on the device (no host copy): what a GPU-resident caller wants."""
    if samples.dim() == 3:
        samples = samples[:, :, :1]  # the reference reads dim 0 (samples[:, :, 0], :207)
    if samples.dim() not in (2, 3):
        raise ValueError(f"Expected samples with 2D shape, got {tuple(samples.shape)}")
    if weights.dim() != 2:
        raise ValueError(f"Expected weights with 2D shape, got {tuple(weights.shape)}")
    if samples.shape[0] != weights.shape[0]:
        raise ValueError("Samples/weights batch size mismatch")
    if samples.shape[1] != weights.shape[1]:
        raise ValueError("Samples/weights sample count mismatch")
    return E.weighted_histogram(samples, weights, int(k))


def _estimate_discrete_posterior_batch(samples: torch.Tensor, weights: torch.Tensor, k: int) -> List[List[float]]:
    """benchmarking/models/vbn.py:226-242 -- same list-of-lists result, one launch instead of B*S Python steps."""
    if samples.dim() == 3:
        samples = samples[:, :, 0]
    if samples.dim() != 2:
        raise ValueError(f"Expected samples with 2D shape, got {tuple(samples.shape)}")
    return estimate_discrete_posterior_tensor(samples, weights, k).cpu().tolist()


def _estimate_discrete_posterior(samples: torch.Tensor, weights: torch.Tensor, k: int) -> List[float]:
    """benchmarking/models/vbn.py:202-223: first query of the batch only (samples[0], weights[0])."""
    if samples.dim() == 3:
        samples = samples[:, :, 0]
    if samples.dim() == 2:
        samples = samples[0]
    if weights.dim() == 2:
        weights = weights[0]
    return _estimate_discrete_posterior_batch(samples.reshape(1, -1), weights.reshape(1, -1), k)[0]


def _extract_samples_1d(samples: torch.Tensor) -> torch.Tensor:
    """benchmarking/models/vbn.py:365-379."""
    arr = torch.as_tensor(samples)
    if arr.dim() == 0:
        return arr.reshape(1)
    if arr.dim() == 3:
        arr = arr.reshape(-1, arr.shape[-1])
    if arr.dim() == 2:
        if arr.shape[1] == 1:
            return arr[:, 0]
        raise ValueError("Multivariate continuous targets are unsupported")
    if arr.dim() == 1:
        return arr
    raise ValueError(f"Unsupported sample shape {tuple(arr.shape)}")


def _continuous_from_samples(samples: Any, weights: Optional[Any] = None) -> dict:
    """benchmarking/models/vbn.py:381-423: weighted mean / std (population) of a continuous target, the
    first 2048 samples, format "normal_params" (or "samples_1d" when the moments are not finite)."""
    vals = _extract_samples_1d(samples)
    if vals.numel() == 0:
        raise ValueError("No samples returned for continuous target")
    n = int(vals.numel())
    wts = None
    if weights is not None:
        w = torch.as_tensor(weights).reshape(-1)
        if w.numel() == n and bool(torch.isfinite(w).any()):
            wts = w
    mean = std = None
    if wts is not None and bool(torch.isfinite(wts).all()):
        # hot path: every weight finite -> clip at 0, normalise, two-pass moments (vbn_posterior_stats);
        # a non-positive total falls back to uniform weights exactly like :399-406
        st = E.posterior_stats(wts.reshape(1, n), vals.reshape(1, n, 1), 0.0)
        mean, std = float(st["mean"][0, 0]), float(st["std"][0, 0])
    elif wts is not None:
        # nan / inf weights: the reference's numpy arithmetic verbatim (rare; not worth a kernel)
        v64, w64 = vals.double(), wts.double().clamp(0.0, float("inf"))
        total = float(w64.sum())
        if total > 0:
            w64 = w64 / total
            mean = float((w64 * v64).sum())
            std = math.sqrt(float((w64 * (v64 - mean) ** 2).sum()))
    if mean is None:
        st = E.posterior_stats(torch.ones(1, n, device=vals.device), vals.reshape(1, n, 1), 0.0)
        mean, std = float(st["mean"][0, 0]), float(st["std"][0, 0])
    keep = [float(x) for x in vals[: min(n, 2048)].cpu().tolist()]
    if math.isfinite(mean) and math.isfinite(std):
        return {"format": "normal_params", "mean": mean, "std": std, "n_samples": n, "samples": keep}
    return {"format": "samples_1d", "samples": keep, "n_samples": n}
