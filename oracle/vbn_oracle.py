"""CPU oracle for the VBN batched posterior-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``vectorizedbayesiannetwork_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or the
timed CPU baseline -- never as the product path.

It is a functional torch-CPU restatement of the reference algorithm (the reference
itself is pure Python + torch, so torch on the host IS the reference's arithmetic;
SURVEY.md section 8c).  Every function cites the reference ``file:line`` it follows
(paths relative to /root/reference).  The model is a plain ``spec`` dict of tensors
(see ``spec_from_reference``), not ``nn.Module`` objects.

Parity pinning: ``tests/test_oracle_pin.py`` checks (in the build container, where
/root/reference exists) that this oracle reproduces the unmodified reference
BIT-FOR-BIT under the same ``torch.manual_seed`` for every CPD kind and every
inference method on the path, and ``tests/golden/*.pt`` holds input/noise/output
vectors recorded from the reference itself (generator: ``tests/golden/make_golden.py``)
that travel to the GPU box.

Randomness is routed through a *noise source* so the same draw can be
  * taken from torch's global generator exactly like the reference does (``TorchNoise``),
  * recorded (``RecordingNoise``) and
  * replayed / injected (``ReplayNoise``) -- the CUDA path accepts the same tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2 * math.pi)

# --------------------------------------------------------------------------------------
# noise sources
# --------------------------------------------------------------------------------------


class TorchNoise:
    """Draws from torch's global generator with the same calls the reference makes
    (SURVEY.md Appendix B), so a run under ``torch.manual_seed(k)`` consumes the stream
    identically to the reference."""

    def normal(self, key, like: torch.Tensor) -> torch.Tensor:
        # torch.randn_like(scale)  (linear_gaussian.py:195, gaussian_nn.py:260, mdn.py:234, kde.py:180)
        return torch.randn_like(like)

    def normal_shape(self, key, shape, like: torch.Tensor) -> torch.Tensor:
        # Normal(loc, scale).sample((b, S)) == loc + scale * randn(shape)  (gaussian_nn.py:254)
        return torch.randn(shape, dtype=like.dtype, device=like.device)

    def uniform(self, key, like: torch.Tensor) -> torch.Tensor:
        # torch.rand_like(center)  (softmax_nn.py:666, 669)
        return torch.rand_like(like)

    def categorical_probs(self, key, probs: torch.Tensor) -> torch.Tensor:
        # Categorical(probs=...).sample()  (mdn.py:229)
        return torch.distributions.Categorical(probs=probs).sample()

    def categorical_logits(self, key, logits: torch.Tensor) -> torch.Tensor:
        # Categorical(logits=...).sample()  (softmax_nn.py:651-652)
        return torch.distributions.Categorical(logits=logits).sample()

    def multinomial(self, key, weights: torch.Tensor) -> torch.Tensor:
        # torch.multinomial(weights, 1).squeeze(-1)  (kde.py:178)
        return torch.multinomial(weights, num_samples=1).squeeze(-1)

    def randint(self, key, n: int, count: int, device) -> torch.Tensor:
        # torch.randint(0, n, (count,))  (kde.py:170)
        return torch.randint(0, n, (count,), device=device)

    def resample(self, key, weights: torch.Tensor, n: int) -> torch.Tensor:
        # torch.multinomial(weights, num_samples=n, replacement=True)  (resampled_importance_sampling.py:38)
        return torch.multinomial(weights, num_samples=n, replacement=True)


class RecordingNoise(TorchNoise):
    """TorchNoise that also stores every draw under its key (list per key, call order)."""

    def __init__(self) -> None:
        self.log: Dict[Tuple, List[torch.Tensor]] = {}

    def _rec(self, key, value):
        self.log.setdefault(tuple(key), []).append(value.clone())
        return value

    def normal(self, key, like):
        return self._rec(key + ("eps",), super().normal(key, like))

    def normal_shape(self, key, shape, like):
        return self._rec(key + ("eps",), super().normal_shape(key, shape, like))

    def uniform(self, key, like):
        return self._rec(key + ("u",), super().uniform(key, like))

    def categorical_probs(self, key, probs):
        return self._rec(key + ("idx",), super().categorical_probs(key, probs))

    def categorical_logits(self, key, logits):
        return self._rec(key + ("idx",), super().categorical_logits(key, logits))

    def multinomial(self, key, weights):
        return self._rec(key + ("idx",), super().multinomial(key, weights))

    def randint(self, key, n, count, device):
        return self._rec(key + ("idx",), super().randint(key, n, count, device))

    def resample(self, key, weights, n):
        return self._rec(key + ("idx",), super().resample(key, weights, n))


class ReplayNoise:
    """Replays a ``RecordingNoise.log`` (or a hand-built dict) in call order."""

    def __init__(self, log: Dict[Tuple, Sequence[torch.Tensor]]) -> None:
        self.log = {tuple(k): list(v) for k, v in log.items()}
        self.pos: Dict[Tuple, int] = {}

    def _next(self, key):
        key = tuple(key)
        i = self.pos.get(key, 0)
        self.pos[key] = i + 1
        return self.log[key][i]

    def normal(self, key, like):
        return self._next(key + ("eps",)).to(like.dtype).reshape(like.shape)

    def normal_shape(self, key, shape, like):
        return self._next(key + ("eps",)).to(like.dtype).reshape(shape)

    def uniform(self, key, like):
        return self._next(key + ("u",)).to(like.dtype).reshape(like.shape)

    def categorical_probs(self, key, probs):
        return self._next(key + ("idx",)).reshape(probs.shape[:-1])

    def categorical_logits(self, key, logits):
        return self._next(key + ("idx",)).reshape(logits.shape[:-1])

    def multinomial(self, key, weights):
        return self._next(key + ("idx",)).reshape(weights.shape[:-1])

    def randint(self, key, n, count, device):
        return self._next(key + ("idx",)).reshape(count)

    def resample(self, key, weights, n):
        return self._next(key + ("idx",)).reshape(weights.shape[0], n).long()


# --------------------------------------------------------------------------------------
# shape helpers (vbn/core/utils.py:56-82)
# --------------------------------------------------------------------------------------


def ensure_2d(x: torch.Tensor) -> torch.Tensor:
    if x.dim() == 1:
        return x.unsqueeze(-1)
    if x.dim() == 2:
        return x
    raise ValueError(f"Expected 1D or 2D tensor, got shape {tuple(x.shape)}")


def broadcast_samples(x: torch.Tensor, n: int) -> torch.Tensor:
    if x.dim() == 2:
        return x.unsqueeze(1).expand(-1, n, -1)
    if x.dim() == 3:
        return x
    raise ValueError(f"Expected 2D or 3D tensor, got shape {tuple(x.shape)}")


def _x3(x: torch.Tensor) -> torch.Tensor:
    if x.dim() <= 2:
        x = ensure_2d(x)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    return x


_ACT = {"relu": F.relu, "tanh": torch.tanh, "gelu": F.gelu, "elu": F.elu}


def mlp_forward(layers, activation: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Sequential(Linear, act, ..., Linear)  (gaussian_nn.py:16-34 ``_build_mlp``)."""
    act = _ACT[activation]
    last = len(layers) - 1
    for i, (w, b) in enumerate(layers):
        x = F.linear(x, w, b)
        if i != last:
            x = act(x)
    return x


def softplus_min(x: torch.Tensor, min_val: float) -> torch.Tensor:
    """safe_softplus (vbn/cpds/utils.py:6-7): F.softplus (beta=1, threshold=20) + min."""
    return F.softplus(x) + float(min_val)


def gaussian_logpdf_sum(x, loc, scale):
    """-0.5*sum_d[(x-loc)^2/scale^2 + 2 ln scale + ln 2pi]  (linear_gaussian.py:214-217)."""
    var = scale**2
    return -0.5 * (((x - loc) ** 2) / var + 2 * torch.log(scale) + LOG_2PI).sum(dim=-1)


# --------------------------------------------------------------------------------------
# linear_gaussian  (vbn/cpds/linear_gaussian.py:163-217)
# --------------------------------------------------------------------------------------


def lg_scale(c) -> torch.Tensor:
    return torch.sqrt(c["var"].clamp(min=float(c["min_scale"]) ** 2))  # :163-165


def lg_params(c, parents: Optional[torch.Tensor]):
    if c["input_dim"] == 0:
        return c["bias"], lg_scale(c)
    if parents.dim() == 2:
        parents = parents.unsqueeze(1)
    b, s, dp = parents.shape
    mu = parents.reshape(b * s, dp) @ c["weight"] + c["bias"]  # :180
    loc = mu.reshape(b, s, c["output_dim"])
    scale = lg_scale(c).view(1, 1, -1).expand(b, s, -1)
    return loc, scale


def lg_sample(c, parents, n, noise, key):
    if c["input_dim"] == 0:
        b = 1 if parents is None else parents.shape[0]  # :187
        loc = c["bias"].view(1, 1, -1).expand(b, n, -1)
        scale = lg_scale(c).view(1, 1, -1).expand(b, n, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        loc, scale = lg_params(c, broadcast_samples(parents, n))
    eps = noise.normal(key, scale)  # :195
    return loc + eps * scale


def lg_log_prob(c, x, parents):
    x = _x3(x)
    if c["input_dim"] == 0:
        b, s, _ = x.shape
        loc = c["bias"].view(1, 1, -1).expand(b, s, -1)
        scale = lg_scale(c).view(1, 1, -1).expand(b, s, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        loc, scale = lg_params(c, broadcast_samples(parents, x.shape[1]))
    return gaussian_logpdf_sum(x, loc, scale)


# --------------------------------------------------------------------------------------
# gaussian_nn  (vbn/cpds/gaussian_nn.py:105-119, 215-288)
# --------------------------------------------------------------------------------------


def gnn_params(c, parents: Optional[torch.Tensor]):
    d = c["output_dim"]
    mean_y = c["mean_y"].view(1, 1, -1)
    std_y = c["std_y"].view(1, 1, -1)
    if c["input_dim"] == 0:
        loc = c["loc"].view(1, 1, -1)
        scale = softplus_min(c["log_scale"], c["min_scale"]).view(1, 1, -1)
        return loc * std_y + mean_y, scale * std_y  # :114-119
    if parents.dim() == 2:
        parents = parents.unsqueeze(1)
    z = (parents - c["mean_x"].view(1, 1, -1)) / c["std_x"].view(1, 1, -1)  # :105-112
    b, s, dp = z.shape
    out = mlp_forward(c["layers"], c["activation"], z.reshape(b * s, dp)).reshape(b, s, 2 * d)
    loc = out[..., :d]
    scale = softplus_min(out[..., d:], c["min_scale"])
    return loc * std_y + mean_y, scale * std_y


def gnn_sample(c, parents, n, noise, key):
    if c["input_dim"] == 0:
        b = 1 if parents is None else parents.shape[0]
        loc, scale = gnn_params(c, None)
        loc, scale = loc.squeeze(0).squeeze(0), scale.squeeze(0).squeeze(0)
        # Normal(loc, scale).sample((b, n))  (:244-254)
        eps = noise.normal_shape(key, (b, n) + tuple(loc.shape), loc)
        return loc + eps * scale
    if parents is None:
        raise ValueError("parents cannot be None when input_dim > 0")
    loc, scale = gnn_params(c, broadcast_samples(parents, n))
    eps = noise.normal(key, scale)  # :260
    return loc + eps * scale


def gnn_log_prob(c, x, parents):
    x = _x3(x)
    if c["input_dim"] == 0:
        loc, scale = gnn_params(c, None)
        loc, scale = loc.squeeze(0).squeeze(0), scale.squeeze(0).squeeze(0)
        # torch.distributions.Normal.log_prob  (:263-278)
        var = scale**2
        lp = -((x - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))
        return lp.sum(dim=-1)
    if parents is None:
        raise ValueError("parents cannot be None when input_dim > 0")
    loc, scale = gnn_params(c, broadcast_samples(parents, x.shape[1]))
    return gaussian_logpdf_sum(x, loc, scale)


# --------------------------------------------------------------------------------------
# rff_gaussian  (vbn/cpds/rff_gaussian.py:131-146, 181-206, 254-291)  -- SURVEY 8f row 3
# --------------------------------------------------------------------------------------


def rff_scale(c) -> torch.Tensor:
    return torch.sqrt(c["var"].clamp(min=float(c["min_scale"]) ** 2))  # :181-183


def rff_features(c, parents: torch.Tensor) -> torch.Tensor:
    z = (parents - c["mean_x"].view(1, -1)) / c["std_x"].view(1, -1)  # :131-136
    proj = z @ c["rff_w"].t() + c["rff_b"]  # :144
    return math.sqrt(2.0 / float(c["n_features"])) * torch.cos(proj)  # :145-146


def rff_params(c, parents: Optional[torch.Tensor]):
    if c["input_dim"] == 0:
        return c["mean_y"], rff_scale(c)  # :188-191
    if parents is None:
        raise ValueError("parents cannot be None when input_dim > 0")
    if parents.dim() == 2:
        parents = parents.unsqueeze(1)
    b, s, dp = parents.shape
    feats = rff_features(c, parents.reshape(b * s, dp))
    loc_norm = (feats @ c["coef"] + c["bias"]).reshape(b, s, c["output_dim"])  # :198-199
    loc = loc_norm * c["std_y"].view(1, 1, -1) + c["mean_y"].view(1, 1, -1)
    return loc, rff_scale(c).view(1, 1, -1).expand(b, s, -1)


def _rff_check(c):
    if not c["stats_ready"]:
        raise RuntimeError("RFFGaussianCPD is not fitted yet.")  # :76-78


def rff_sample(c, parents, n, noise, key):
    _rff_check(c)
    if c["input_dim"] == 0:
        b = 1 if parents is None else parents.shape[0]
        loc = c["mean_y"].view(1, 1, -1).expand(b, n, -1)
        scale = rff_scale(c).view(1, 1, -1).expand(b, n, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        loc, scale = rff_params(c, broadcast_samples(parents, n))
    eps = noise.normal(key, scale)  # torch.randn_like(scale)  (:266)
    return loc + eps * scale


def rff_log_prob(c, x, parents):
    _rff_check(c)
    x = _x3(x)
    if c["input_dim"] == 0:
        b, s, _ = x.shape
        loc = c["mean_y"].view(1, 1, -1).expand(b, s, -1)
        scale = rff_scale(c).view(1, 1, -1).expand(b, s, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        loc, scale = rff_params(c, broadcast_samples(parents, x.shape[1]))
    return gaussian_logpdf_sum(x, loc, scale)  # :287-291


# --------------------------------------------------------------------------------------
# mdn  (vbn/cpds/mdn.py:185-272)
# --------------------------------------------------------------------------------------


def mdn_params(c, parents: Optional[torch.Tensor]):
    k, d = c["n_components"], c["output_dim"]
    if c["input_dim"] == 0:
        return c["logits"], c["loc"], softplus_min(c["log_scale"], c["min_scale"])
    if parents.dim() == 2:
        parents = parents.unsqueeze(1)
    b, s, dp = parents.shape
    out = mlp_forward(c["layers"], c["activation"], parents.reshape(b * s, dp)).reshape(b, s, -1)
    logits = out[..., :k]
    rest = out[..., k:].reshape(b, s, k, 2 * d)  # :200-203
    loc = rest[..., :d]
    scale = softplus_min(rest[..., d:], c["min_scale"])
    return logits, loc, scale


def mdn_mixture_weights(logits):
    pi = torch.softmax(logits, dim=-1).clamp_min(1e-5)  # :227 / :269
    return pi / pi.sum(dim=-1, keepdim=True).clamp_min(1e-12)


def mdn_sample(c, parents, n, noise, key):
    k, d = c["n_components"], c["output_dim"]
    if c["input_dim"] == 0:
        b = 1 if parents is None else parents.shape[0]
        logits = c["logits"].view(1, 1, -1).expand(b, n, -1)
        loc = c["loc"].view(1, 1, k, d).expand(b, n, -1, -1)
        scale = softplus_min(c["log_scale"], c["min_scale"]).view(1, 1, k, d).expand(b, n, -1, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits, loc, scale = mdn_params(c, broadcast_samples(parents, n))
    b, s, _ = logits.shape
    pi = mdn_mixture_weights(logits)
    comps = noise.categorical_probs(key, pi.reshape(b * s, -1)).reshape(b, s)  # :229
    idx = comps.unsqueeze(-1).unsqueeze(-1).expand(-1, -1, 1, d)
    loc = loc.gather(dim=2, index=idx).squeeze(2)
    scale = scale.gather(dim=2, index=idx).squeeze(2)
    eps = noise.normal(key, loc)  # :234
    return loc + eps * scale


def mdn_log_prob(c, x, parents):
    k, d = c["n_components"], c["output_dim"]
    x = _x3(x)
    if c["input_dim"] == 0:
        b, s, _ = x.shape
        logits = c["logits"].view(1, 1, -1).expand(b, s, -1)
        loc = c["loc"].view(1, 1, k, d).expand(b, s, -1, -1)
        scale = softplus_min(c["log_scale"], c["min_scale"])
        log_scale = torch.log(scale.view(1, 1, k, d).expand(b, s, -1, -1))
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits, loc, scale = mdn_params(c, broadcast_samples(parents, x.shape[1]))
        log_scale = torch.log(scale)
    x_exp = x.unsqueeze(2).expand(-1, -1, k, -1)
    var = torch.exp(2 * log_scale)  # :265
    log_comp = -0.5 * (((x_exp - loc) ** 2) / var + 2 * log_scale + LOG_2PI).sum(dim=-1)
    log_pi = torch.log(mdn_mixture_weights(logits))
    return torch.logsumexp(log_pi + log_comp, dim=-1)  # :272


# --------------------------------------------------------------------------------------
# softmax_nn  (vbn/cpds/softmax_nn.py:581-759)
# --------------------------------------------------------------------------------------


def snn_logits(c, parents: Optional[torch.Tensor], b: int, s: int) -> torch.Tensor:
    d, k = c["output_dim"], c["n_classes"]
    t = float(c.get("temperature", 1.0))
    if c["input_dim"] == 0:
        if c["root_ready"]:  # :636-641
            logits = c["root_log_probs"].view(1, 1, d, k).expand(b, s, -1, -1)
            return torch.log_softmax(logits / t, dim=-1)
        return c["logits"].view(1, 1, d, k).expand(b, s, -1, -1) / t
    if parents.dim() == 2:
        parents = parents.unsqueeze(1)
    pb, ps, dp = parents.shape
    out = mlp_forward(c["layers"], c["activation"], parents.reshape(pb * ps, dp))
    return out.reshape(pb, ps, d, k) / t  # :581-587


def snn_gather_bin_edges(c, indices: torch.Tensor):
    """(:589-610)"""
    d, k = c["output_dim"], c["n_classes"]
    edges = c["bin_edges"]
    if indices.dim() == 3:
        d_idx = torch.arange(d).view(1, 1, -1).expand(indices.shape[0], indices.shape[1], -1)
    else:
        d_idx = torch.arange(d).view(1, -1).expand(indices.shape[0], -1)
    idx = indices.clamp(min=0, max=k - 1)
    left = edges[d_idx, idx]
    right = edges[d_idx, (idx + 1).clamp(max=k)]
    width = torch.clamp(right - left, min=float(c["min_bin_width"]))
    center = 0.5 * (left + right)
    return left, right, width, center


def snn_x_to_bin(c, x: torch.Tensor) -> torch.Tensor:
    """(:612-630)"""
    if not c["bins_ready"]:
        raise RuntimeError("Bins not initialized. Call fit(...) before sampling.")
    k = c["n_classes"]
    edges = c["bin_edges"].to(dtype=x.dtype)
    flat = x.reshape(-1, x.shape[-1])
    cont = (flat.unsqueeze(-1) >= edges.unsqueeze(0)).sum(dim=-1) - 1
    cont = cont.clamp(min=0, max=k - 1)
    is_disc = c["is_discrete"]
    if bool(is_disc.any()):
        cv = c["class_values"].to(dtype=x.dtype)
        match = flat.unsqueeze(-1) == cv.unsqueeze(0)
        disc = match.long().argmax(dim=-1)
        missing = (~match.any(dim=-1)) & is_disc.unsqueeze(0)
        if bool(missing.any()):
            raise ValueError("Found values outside discrete class set.")
        bins = torch.where(is_disc.unsqueeze(0), disc, cont)
    else:
        bins = cont
    return bins.reshape(*x.shape)


def snn_sample(c, parents, n, noise, key):
    if not c["bins_ready"]:
        raise RuntimeError("Bins not initialized. Call fit(...) before sampling.")
    d, k = c["output_dim"], c["n_classes"]
    if c["input_dim"] == 0:
        b = 1 if parents is None else parents.shape[0]
        logits = snn_logits(c, None, b, n)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = snn_logits(c, broadcast_samples(parents, n), 0, 0)
    indices = noise.categorical_logits(key, logits)  # :651-652
    values = c["sample_values"].to(dtype=logits.dtype).view(1, 1, d, k)
    values = values.expand(indices.shape[0], indices.shape[1], -1, -1)
    disc_values = values.gather(-1, indices.unsqueeze(-1)).squeeze(-1)
    left, right, width, center = snn_gather_bin_edges(c, indices)
    wb = c["within_bin"]
    if wb == "uniform":
        u = noise.uniform(key, center)
        cont = left + u * width
    elif wb == "triangular":
        u = noise.uniform(key, center)
        lv = left + width * torch.sqrt(torch.clamp(u * 0.5, min=0.0))
        rv = right - width * torch.sqrt(torch.clamp((1.0 - u) * 0.5, min=0.0))
        cont = torch.where(u < 0.5, lv, rv)
    elif wb == "gaussian":
        sigma = torch.clamp(float(c["within_bin_scale"]) * width, min=float(c["min_bin_width"]))
        cont = center + noise.normal(key, center) * sigma
    else:
        raise ValueError(f"Unknown within_bin '{wb}'")
    if c["within_bin_clip"]:
        cont = cont.clamp(min=left, max=right)
    is_disc = c["is_discrete"]
    if bool(is_disc.any()):
        return torch.where(is_disc.view(1, 1, -1), disc_values, cont)
    return cont


def snn_log_prob(c, x, parents):
    x = _x3(x)
    if c["input_dim"] == 0:
        b, s, _ = x.shape
        logits = snn_logits(c, None, b, s)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = snn_logits(c, broadcast_samples(parents, x.shape[1]), 0, 0)
    bins = snn_x_to_bin(c, x).long()
    log_probs = torch.log_softmax(logits, dim=-1)
    log_bin = log_probs.gather(-1, bins.unsqueeze(-1)).squeeze(-1)
    left, right, width, center = snn_gather_bin_edges(c, bins)
    clip = bool(c["within_bin_clip"])
    x_use = x.clamp(min=left, max=right) if clip else x
    wb = c["within_bin"]
    mbw = float(c["min_bin_width"])
    neg_inf = float("-inf")
    if wb == "uniform":
        log_within = -torch.log(width)
        if not clip:
            inside = (x >= left) & (x <= right)
            log_within = torch.where(inside, log_within, torch.full_like(log_within, neg_inf))
    elif wb == "triangular":
        denom_l = torch.clamp(width * (center - left), min=mbw**2)
        denom_r = torch.clamp(width * (right - center), min=mbw**2)
        left_pdf = 2.0 * (x_use - left) / denom_l
        right_pdf = 2.0 * (right - x_use) / denom_r
        pdf = torch.clamp(torch.where(x_use <= center, left_pdf, right_pdf), min=0.0)
        log_within = torch.log(torch.clamp(pdf, min=1e-12))
        if not clip:
            inside = (x >= left) & (x <= right)
            log_within = torch.where(inside, log_within, torch.full_like(log_within, neg_inf))
    elif wb == "gaussian":
        sigma = torch.clamp(float(c["within_bin_scale"]) * width, min=mbw)
        var = sigma**2
        log_within = -((x_use - center) ** 2) / (2 * var) - sigma.log() - math.log(math.sqrt(2 * math.pi))
    else:
        raise ValueError(f"Unknown within_bin '{wb}'")
    is_disc = c["is_discrete"]
    if bool(is_disc.any()):
        mask = (~is_disc).view(1, 1, -1)
        log_within = torch.where(mask, log_within, torch.zeros_like(log_within))
    return (log_bin + log_within).sum(dim=-1)


# --------------------------------------------------------------------------------------
# categorical_table  (vbn/cpds/categorical_table.py:12-21, 232-255, 359-417)  -- SURVEY 8f row 3
# spec: counts [D, n_cfg, C] (smoothing already applied by fit), class_values [D, C],
#       class_mask [D, C], parent_values [tensor per parent dim], parent_strides [int per parent dim]
# --------------------------------------------------------------------------------------


def ct_map_values(values: torch.Tensor, support: torch.Tensor) -> torch.Tensor:
    support = support.contiguous()
    values = values.contiguous()
    idx = torch.searchsorted(support, values)  # :12-21
    if torch.any(idx < 0) or torch.any(idx >= support.shape[0]):
        raise ValueError("Found values outside support.")
    if not torch.all(support[idx] == values):
        raise ValueError("Found values outside support.")
    return idx


def ct_parents_to_index(c, parents: torch.Tensor) -> torch.Tensor:
    if c["input_dim"] == 0:  # :232-245
        return torch.zeros(parents.shape[0], dtype=torch.long)
    idx = torch.zeros(parents.shape[0], dtype=torch.long)
    for d, support in enumerate(c["parent_values"]):
        idx = idx + ct_map_values(parents[:, d], support.to(dtype=parents.dtype)) * int(c["parent_strides"][d])
    return idx


def ct_logits(c, parents: Optional[torch.Tensor]) -> torch.Tensor:
    counts = c["counts"]
    d = c["output_dim"]
    if parents is None:  # :371-378
        probs = counts / counts.sum(dim=-1, keepdim=True).clamp_min(1e-12)
        return torch.log(probs.clamp_min(1e-12)).view(1, 1, d, -1)
    if parents.dim() == 2:  # :359-369
        parents = parents.unsqueeze(1)
    b, s, dp = parents.shape
    parent_idx = ct_parents_to_index(c, parents.reshape(b * s, dp))
    probs = counts[:, parent_idx, :]
    probs = probs / probs.sum(dim=-1, keepdim=True).clamp_min(1e-12)
    logits = torch.log(probs.clamp_min(1e-12))
    return logits.permute(1, 0, 2).reshape(b, s, d, -1)


def ct_sample(c, parents, n, noise, key):
    d = c["output_dim"]
    if c["input_dim"] == 0:  # :379-396
        b = 1 if parents is None else parents.shape[0]
        logits = ct_logits(c, None).expand(b, n, -1, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = ct_logits(c, broadcast_samples(parents, n))
    indices = noise.categorical_logits(key, logits)
    values = c["class_values"].to(dtype=logits.dtype).view(1, 1, d, -1)
    values = values.expand(indices.shape[0], indices.shape[1], -1, -1)
    return values.gather(-1, indices.unsqueeze(-1)).squeeze(-1)


def ct_log_prob(c, x, parents):
    if x.dim() <= 2:  # :398-417
        x = ensure_2d(x)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    d = c["output_dim"]
    if c["input_dim"] == 0:
        logits = ct_logits(c, None).expand(x.shape[0], x.shape[1], -1, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = ct_logits(c, broadcast_samples(parents, x.shape[1]))
    log_probs = torch.log_softmax(logits, dim=-1)
    x_flat = x.reshape(-1, x.shape[-1])
    targets = torch.zeros(x_flat.shape[0], d, dtype=torch.long)
    for k in range(d):  # :246-255
        support = c["class_values"][k][c["class_mask"][k]].to(dtype=x_flat.dtype)
        targets[:, k] = ct_map_values(x_flat[:, k], support)
    targets = targets.reshape(x.shape[0], x.shape[1], d)
    return log_probs.gather(-1, targets.unsqueeze(-1)).squeeze(-1).sum(dim=-1)


# --------------------------------------------------------------------------------------
# categorical_embedded_softmax  (vbn/cpds/categorical_embedded_softmax.py:36-46, 280-329, 469-511)
# -- SURVEY 8f row 3.  spec: embeddings [tensor [card_d, E] per parent dim], layers, activation,
#    logits [D, C] (root), class_values / class_mask [D, C], parent_values [tensor per parent dim]
# --------------------------------------------------------------------------------------


def ces_logits(c, parents: torch.Tensor) -> torch.Tensor:
    if parents.dim() == 2:  # :316-329
        parents = parents.unsqueeze(1)
    b, s, dp = parents.shape
    flat = parents.reshape(b * s, dp)
    d = c["output_dim"]
    if c["input_dim"] == 0:
        logits = c["logits"].view(1, 1, d, -1).expand(b, s, -1, -1)
    else:
        idx = torch.zeros(flat.shape[0], c["input_dim"], dtype=torch.long)  # :280-293
        for k, support in enumerate(c["parent_values"]):
            idx[:, k] = ct_map_values(flat[:, k], support.to(dtype=flat.dtype))
        feats = torch.cat([emb[idx[:, k]] for k, emb in enumerate(c["embeddings"])], dim=-1)  # :306-314
        logits = mlp_forward(c["layers"], c["activation"], feats).reshape(b, s, d, -1)
    return logits.masked_fill(~c["class_mask"].view(1, 1, d, -1), -1e9)


def ces_logits_or_root(c, parents: Optional[torch.Tensor]) -> torch.Tensor:
    """What the exact / Rao-Blackwellized methods read: cpd._logits for a root (no mask), else
    _logits_from_parents (rao_blackwellized_marginalization.py:163-175, categorical_exact.py:48-71)."""
    if parents is None:
        return c["logits"].view(1, 1, c["output_dim"], -1)
    return ces_logits(c, parents)


def _ces_check(c):
    if not c["stats_ready"]:
        raise RuntimeError("CategoricalEmbeddedSoftmaxCPD is not fitted yet.")  # :136-138


def ces_sample(c, parents, n, noise, key):
    _ces_check(c)
    d = c["output_dim"]
    if c["input_dim"] == 0:  # :471-475 (root logits are NOT masked here)
        b = 1 if parents is None else parents.shape[0]
        logits = c["logits"].view(1, 1, d, -1).expand(b, n, -1, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = ces_logits(c, broadcast_samples(parents, n))
    indices = noise.categorical_logits(key, logits)
    values = c["class_values"].to(dtype=logits.dtype).view(1, 1, d, -1)
    values = values.expand(indices.shape[0], indices.shape[1], -1, -1)
    return values.gather(-1, indices.unsqueeze(-1)).squeeze(-1)


def ces_log_prob(c, x, parents):
    _ces_check(c)
    if x.dim() <= 2:  # :490-511
        x = ensure_2d(x)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    d = c["output_dim"]
    if c["input_dim"] == 0:
        logits = c["logits"].view(1, 1, d, -1).expand(x.shape[0], x.shape[1], -1, -1)
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        logits = ces_logits(c, broadcast_samples(parents, x.shape[1]))
    log_probs = torch.log_softmax(logits, dim=-1)
    x_flat = x.reshape(-1, x.shape[-1])
    targets = torch.zeros(x_flat.shape[0], d, dtype=torch.long)
    for k in range(d):  # :295-304
        support = c["class_values"][k][c["class_mask"][k]].to(dtype=x_flat.dtype)
        targets[:, k] = ct_map_values(x_flat[:, k], support)
    targets = targets.reshape(x.shape[0], x.shape[1], d)
    return log_probs.gather(-1, targets.unsqueeze(-1)).squeeze(-1).sum(dim=-1)


# --------------------------------------------------------------------------------------
# kde  (vbn/cpds/kde.py:105-182)
# --------------------------------------------------------------------------------------


def kde_kernel_log(c, diff: torch.Tensor, bandwidth: float) -> torch.Tensor:
    scale = max(float(bandwidth), 1e-3) + float(c["min_scale"])  # :105-109
    return -0.5 * ((diff / scale) ** 2 + LOG_2PI + 2 * math.log(scale))


def kde_log_prob(c, x, parents, chunk: int = 512):
    if c.get("targets") is None:
        raise RuntimeError("KDECPD is not fitted yet.")
    x = _x3(x)
    targets = c["targets"]
    b, s, dx = x.shape
    n = targets.shape[0]
    flat_x = x.reshape(b * s, dx)
    flat_p = None
    if c["input_dim"] != 0:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        flat_p = broadcast_samples(parents, s).reshape(b * s, c["input_dim"])
    out = []
    for start in range(0, flat_x.shape[0], chunk):
        end = min(start + chunk, flat_x.shape[0])
        diff_y = flat_x[start:end].unsqueeze(1) - targets.unsqueeze(0)
        log_ky = kde_kernel_log(c, diff_y, c["bandwidth"]).sum(dim=-1)
        if c["input_dim"] == 0:
            out.append(torch.logsumexp(log_ky, dim=1) - math.log(float(n)))
        else:
            diff_p = flat_p[start:end].unsqueeze(1) - c["parents"].unsqueeze(0)
            log_kp = kde_kernel_log(c, diff_p, c["parent_bandwidth"]).sum(dim=-1)
            out.append(torch.logsumexp(log_kp + log_ky, dim=1) - torch.logsumexp(log_kp, dim=1))
    return torch.cat(out, dim=0).reshape(b, s)


def kde_sample(c, parents, n_samples, noise, key, chunk: int = 512):
    if c.get("targets") is None:
        raise RuntimeError("KDECPD is not fitted yet.")
    targets = c["targets"]
    n = targets.shape[0]
    b = 1 if parents is None else parents.shape[0]
    flat_out = torch.empty(b * n_samples, c["output_dim"], dtype=targets.dtype)
    flat_p = None
    if c["input_dim"] != 0:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        flat_p = broadcast_samples(parents, n_samples).reshape(b * n_samples, c["input_dim"])
    bw = max(float(c["bandwidth"]), 1e-3)
    for start in range(0, flat_out.shape[0], chunk):
        end = min(start + chunk, flat_out.shape[0])
        if c["input_dim"] == 0:
            idx = noise.randint(key, n, end - start, targets.device)  # :170
        else:
            diff_p = flat_p[start:end].unsqueeze(1) - c["parents"].unsqueeze(0)
            log_kp = kde_kernel_log(c, diff_p, c["parent_bandwidth"]).sum(dim=-1)
            idx = noise.multinomial(key, torch.softmax(log_kp, dim=-1))  # :177-178
        selected = targets[idx]
        flat_out[start:end] = selected + noise.normal(key, selected) * (bw + float(c["min_scale"]))
    return flat_out.reshape(b, n_samples, c["output_dim"])


# --------------------------------------------------------------------------------------
# CPD dispatch (BaseCPD.sample / log_prob / forward, vbn/core/base.py:28-59)
# --------------------------------------------------------------------------------------

_SAMPLE = {
    "linear_gaussian": lg_sample,
    "gaussian_nn": gnn_sample,
    "mdn": mdn_sample,
    "softmax_nn": snn_sample,
    "kde": kde_sample,
    "categorical_table": ct_sample,
    "rff_gaussian": rff_sample,
    "categorical_embedded_softmax": ces_sample,
}
_LOG_PROB = {
    "linear_gaussian": lg_log_prob,
    "gaussian_nn": gnn_log_prob,
    "mdn": mdn_log_prob,
    "softmax_nn": snn_log_prob,
    "kde": kde_log_prob,
    "categorical_table": ct_log_prob,
    "rff_gaussian": rff_log_prob,
    "categorical_embedded_softmax": ces_log_prob,
}


# CPDs whose conditional is one Gaussian: read through _weight/_bias/_var or a callable _params by the
# exact / Rao-Blackwellized methods and CPDHandle.conditional (cpd_handle.py:40-70)
_GAUSSIAN_PARAMS = {"linear_gaussian": lg_params, "gaussian_nn": gnn_params, "rff_gaussian": rff_params}
# categorical CPDs over strictly discrete parents: logits of the row's parent configuration
_TABLE_LOGITS = {"categorical_table": ct_logits, "categorical_embedded_softmax": ces_logits_or_root}


def cpd_sample(c, parents, n_samples: int, noise=None, key=("cpd",)):
    return _SAMPLE[c["kind"]](c, parents, int(n_samples), noise or TorchNoise(), tuple(key))


def cpd_log_prob(c, x, parents):
    return _LOG_PROB[c["kind"]](c, x, parents)


def cpd_forward(c, parents, n_samples: int, noise=None, key=("cpd",)):
    samples = cpd_sample(c, parents, n_samples, noise, key)
    lp = cpd_log_prob(c, samples, parents)
    return samples, lp, torch.exp(lp)


# --------------------------------------------------------------------------------------
# inference state (vbn/inference/_core.py:57-135, vbn/utils/__init__.py:46-61)
# --------------------------------------------------------------------------------------


def infer_batch_size(evidence, do) -> int:
    evidence = evidence or {}
    do = do or {}
    if evidence:
        batch = int(next(iter(evidence.values())).shape[0])
        if do and int(next(iter(do.values())).shape[0]) != batch:
            raise ValueError("Evidence and do batch sizes must match.")
        return batch
    if do:
        return int(next(iter(do.values())).shape[0])
    return 1


def clamp_evidence(x: torch.Tensor) -> torch.Tensor:
    x = torch.nan_to_num(x, nan=0.0, posinf=1e6, neginf=-1e6)  # _core.py:112-114
    return x.clamp(min=-1e6, max=1e6)


class _State:
    def __init__(self, spec, query):
        topo = list(spec["topo"])
        self.topo = topo
        n2i = {n: i for i, n in enumerate(topo)}
        self.node_to_idx = n2i
        self.parent_idx = [tuple(n2i[p] for p in spec["parents"][n]) for n in topo]
        self.evidence_mask = [n in query["evidence"] for n in topo]
        self.do_mask = [n in query.get("do", {}) for n in topo]
        self.target_idx = n2i[query["target"]]
        self.slices = []
        off = 0
        for n in topo:
            d = int(spec["cpds"][n]["output_dim"])
            self.slices.append(slice(off, off + d))
            off += d
        self.total_dim = off


def _fixed_values(query, st: _State, dtype, clamp_obs: bool):
    vals: List[Optional[torch.Tensor]] = [None] * len(st.topo)
    for node, v in (query.get("do") or {}).items():  # _core.py:117-135: do first, then evidence
        vals[st.node_to_idx[node]] = ensure_2d(v).to(dtype)
    for node, v in query["evidence"].items():
        v = ensure_2d(v).to(dtype)
        vals[st.node_to_idx[node]] = clamp_evidence(v) if clamp_obs else v
    return vals


def _gather_parents(samples, st: _State, idx: int):
    pidx = st.parent_idx[idx]
    if not pidx:
        return None
    return torch.cat([samples[..., st.slices[p]] for p in pidx], dim=-1)


def _norm_query(query):
    return {
        "target": query["target"],
        "evidence": dict(query.get("evidence") or {}),
        "do": dict(query.get("do") or {}),
    }


# --------------------------------------------------------------------------------------
# likelihood weighting (vbn/inference/likelihood_weighting.py:24-82)
# --------------------------------------------------------------------------------------


def likelihood_weighting(spec, query, n_samples: int, noise=None, scope="lw", normalize=True,
                         eps: float = 1e-12, return_all: bool = False):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    dtype = torch.float32
    samples = torch.zeros(b, n_samples, st.total_dim, dtype=dtype)
    logw = torch.zeros(b, n_samples, dtype=dtype)
    fixed = _fixed_values(query, st, dtype, clamp_obs=True)
    for idx, node in enumerate(st.topo):
        c = spec["cpds"][node]
        if fixed[idx] is not None:
            value = fixed[idx].unsqueeze(1).expand(b, n_samples, -1)
            samples[..., st.slices[idx]] = value
            if st.evidence_mask[idx]:
                logw = logw + cpd_log_prob(c, value, _gather_parents(samples, st, idx))
            continue
        samples[..., st.slices[idx]] = cpd_sample(
            c, _gather_parents(samples, st, idx), n_samples, noise, (scope, node)
        )
    target = samples[..., st.slices[st.target_idx]]
    if normalize:
        w = torch.softmax(logw, dim=1)
    else:
        lw = logw - logw.max(dim=-1, keepdim=True).values
        w = torch.exp(lw).clamp_min(eps)
    if return_all:
        return w, target, samples, logw
    return w, target


# --------------------------------------------------------------------------------------
# resampled importance sampling (vbn/inference/resampled_importance_sampling.py:13-105)
# --------------------------------------------------------------------------------------


def resampled_importance_sampling(spec, query, n_samples: int, noise=None, ess_threshold: float = 0.5,
                                  resample: bool = True, clamp_obs: bool = True, return_info: bool = False):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    dtype = torch.float32
    samples = torch.zeros(b, n_samples, st.total_dim, dtype=dtype)
    logw = torch.zeros(b, n_samples, dtype=dtype)
    fixed = _fixed_values(query, st, dtype, clamp_obs=clamp_obs)
    threshold = max(1.0, ess_threshold * float(n_samples)) if ess_threshold <= 1.0 else float(ess_threshold)  # :62-65
    resampled, last_ess = False, None
    for idx, node in enumerate(st.topo):
        c = spec["cpds"][node]
        if fixed[idx] is not None:  # :72-91
            value = fixed[idx].unsqueeze(1).expand(b, n_samples, -1)
            samples[..., st.slices[idx]] = value
            if st.evidence_mask[idx]:
                logw = logw + cpd_log_prob(c, value, _gather_parents(samples, st, idx))
                if resample:
                    weights = torch.softmax(logw, dim=1)
                    ess = 1.0 / (weights**2).sum(dim=1)
                    last_ess = ess
                    if torch.any(ess < threshold):
                        w = torch.softmax(logw, dim=1)  # _resample :33-41
                        pick = noise.resample(("ris", "__resample__"), w, n_samples)
                        samples = samples[torch.arange(b).unsqueeze(1), pick]
                        logw = torch.zeros_like(logw)
                        resampled = True
            continue
        samples[..., st.slices[idx]] = cpd_sample(c, _gather_parents(samples, st, idx), n_samples, noise, ("ris", node))
    w = torch.softmax(logw, dim=1)
    target = samples[..., st.slices[st.target_idx]]
    if return_info:
        return w, target, {"resampled": resampled, "ess": last_ess}
    return w, target


# --------------------------------------------------------------------------------------
# rao_blackwellized_marginalization (vbn/inference/rao_blackwellized_marginalization.py:15-324)
# --------------------------------------------------------------------------------------


def _descendants(spec, node: str) -> set:
    children: Dict[str, List[str]] = {n: [] for n in spec["nodes"]}
    for n in spec["nodes"]:
        for p in spec["parents"][n]:
            children[p].append(n)
    out, stack = set(), [node]
    while stack:
        for c in children[stack.pop()]:
            if c not in out:
                out.add(c)
                stack.append(c)
    return out


def rb_normalized_weights(logw: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    logw = torch.nan_to_num(logw, nan=-1e30, posinf=1e30, neginf=-1e30)  # :68-77
    logw = logw - logw.max(dim=1, keepdim=True).values
    w = torch.exp(logw)
    denom = w.sum(dim=1, keepdim=True)
    uniform = torch.full_like(w, 1.0 / max(1, w.shape[1]))
    return torch.where(denom > eps, w / denom.clamp_min(eps), uniform)


def rao_blackwellized_marginalization(spec, query, n_samples: int, noise=None, n_particles: Optional[int] = None,
                                      stddevs: float = 4.0, min_scale: float = 1e-6, return_info: bool = False):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    n_samples = max(1, int(n_samples))
    n_particles = max(1, int(n_particles if n_particles is not None else n_samples))
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    t = st.target_idx
    tnode = st.topo[t]

    def fallback(reason):
        out = likelihood_weighting(spec, query, n_samples, noise=noise)
        return (*out, {"fallback": True, "reason": reason}) if return_info else out

    def done(*out):
        return (*out, {"fallback": False, "reason": None}) if return_info else out

    desc = {st.topo.index(n) for n in _descendants(spec, tnode)}
    if any(st.evidence_mask[i] or st.do_mask[i] for i in desc):  # :213-220
        return fallback("target has observed/intervened descendants")
    fixed = _fixed_values(query, st, torch.float32, clamp_obs=True)
    if fixed[t] is not None:  # :223-227
        return done(torch.ones(b, 1), fixed[t].unsqueeze(1).expand(b, 1, -1))
    skip = set(desc) | {t}
    samples = torch.zeros(b, n_particles, st.total_dim)
    logw = torch.zeros(b, n_particles)
    for idx, node in enumerate(st.topo):  # :237-256
        if idx in skip:
            continue
        c = spec["cpds"][node]
        parents = _gather_parents(samples, st, idx)
        if fixed[idx] is not None:
            value = fixed[idx].unsqueeze(1).expand(b, n_particles, -1)
            samples[..., st.slices[idx]] = value
            if st.evidence_mask[idx]:
                logw = logw + cpd_log_prob(c, value, parents)
            continue
        samples[..., st.slices[idx]] = cpd_sample(c, parents, n_particles, noise, ("rb", node))
    w = rb_normalized_weights(logw)
    parents = _gather_parents(samples, st, t)
    c = spec["cpds"][tnode]
    if c["kind"] in ("softmax_nn",) + tuple(_TABLE_LOGITS):  # _target_categorical_probs :155-194
        if c["kind"] == "softmax_nn":
            if parents is None:
                raw = c["root_log_probs"] if c["root_ready"] else c["logits"]
                logits = (torch.log_softmax(raw, dim=-1) if c["root_ready"] else raw).view(1, 1, c["output_dim"], -1)
                logits = logits.expand(b, 1, -1, -1)
            else:
                logits = snn_logits(c, parents, b, n_particles)
            support = c["sample_values"][0]
        else:
            logits = _TABLE_LOGITS[c["kind"]](c, parents)
            if parents is None:
                logits = logits.expand(b, 1, -1, -1)
            support = c["class_values"][0]
        probs = torch.softmax(logits, dim=-1)
        if probs.dim() == 4 and probs.shape[2] == 1:
            probs = probs[:, :, 0, :]
        if probs.dim() == 3 and probs.shape[-1] == int(c["n_classes"]):
            if probs.shape[1] != n_particles:
                if probs.shape[1] != 1:
                    return fallback("categorical conditional shape mismatch")
                probs = probs.expand(-1, n_particles, -1)
            marginal = (w.unsqueeze(-1) * probs).sum(dim=1)  # :277-279
            return done(marginal, support.to(dtype=marginal.dtype).view(1, -1, 1).expand(b, -1, 1))
    if c["kind"] in _GAUSSIAN_PARAMS and c["output_dim"] == 1:  # _target_gaussian_params :92-153
        loc, scale = _GAUSSIAN_PARAMS[c["kind"]](c, parents)
        loc = loc.reshape(1, 1, -1) if loc.dim() == 1 else loc
        scale = scale.reshape(1, 1, -1) if scale.dim() == 1 else scale
        if loc.shape[0] == 1 and b > 1:
            loc = loc.expand(b, -1, -1)
        if scale.shape[0] == 1 and b > 1:
            scale = scale.expand(b, -1, -1)
        if scale.shape[1] == 1 and loc.shape[1] > 1:
            scale = scale.expand(-1, loc.shape[1], -1)
        scale = torch.nan_to_num(scale, nan=min_scale, posinf=min_scale, neginf=min_scale).abs().clamp_min(min_scale)
        if loc.shape[1] != n_particles:
            if loc.shape[1] != 1:
                return fallback("gaussian conditional shape mismatch")
            loc = loc.expand(-1, n_particles, -1)
            scale = scale.expand(-1, n_particles, -1)
        comp_var = scale.squeeze(-1) ** 2  # :296-317
        comp_mean = loc.squeeze(-1)
        mix_mean = (w * comp_mean).sum(dim=1)
        second = (w * (comp_var + comp_mean**2)).sum(dim=1)
        mix_std = (second - mix_mean**2).clamp_min(min_scale**2).sqrt()
        z = torch.linspace(0.0, 1.0, n_samples).view(1, n_samples, 1)
        lo = (mix_mean - stddevs * mix_std).view(b, 1, 1)
        hi = (mix_mean + stddevs * mix_std).view(b, 1, 1)
        grid = lo + (hi - lo) * z
        x = grid.squeeze(-1).unsqueeze(1)
        mu = loc.squeeze(-1).unsqueeze(-1)
        sigma = scale.squeeze(-1).unsqueeze(-1).clamp_min(min_scale)
        zn = (x - mu) / sigma
        comp_pdf = torch.exp(-0.5 * zn**2) / (math.sqrt(2.0 * math.pi) * sigma)
        return done((w.unsqueeze(-1) * comp_pdf).sum(dim=1), grid)
    return fallback("unsupported target CPD for RB marginalization")


# --------------------------------------------------------------------------------------
# exact methods: closed form only when every parent of the target is fixed,
# otherwise the configured fallback method runs (vbn/inference/gaussian_exact.py:134-183,
# vbn/inference/categorical_exact.py:89-128)
# --------------------------------------------------------------------------------------


def _exact_parent_tensor(st: "_State", fixed, b: int):
    pidx = st.parent_idx[st.target_idx]
    if not all(fixed[p] is not None for p in pidx):
        return False, None
    if not pidx:
        return True, None
    return True, torch.cat([fixed[p].unsqueeze(1).expand(b, 1, -1) for p in pidx], dim=-1)


def gaussian_exact(spec, query, n_samples: int, noise=None, stddevs: float = 4.0, min_scale: float = 1e-6,
                   return_info: bool = False):
    query = _norm_query(query)
    n_samples = max(1, int(n_samples))
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    fixed = _fixed_values(query, st, torch.float32, clamp_obs=True)
    t = st.target_idx
    c = spec["cpds"][st.topo[t]]

    def fallback():
        out = likelihood_weighting(spec, query, n_samples, noise=noise)
        return (*out, {"exact": False}) if return_info else out

    if c["output_dim"] != 1:  # :146-147
        return fallback()
    if fixed[t] is not None:  # :149-153
        out = (torch.ones(b, 1), fixed[t].unsqueeze(1).expand(b, 1, -1))
        return (*out, {"exact": True}) if return_info else out
    ok, parents = _exact_parent_tensor(st, fixed, b)
    if not ok or c["kind"] not in _GAUSSIAN_PARAMS:  # :155-170 (_gaussian_params -> None)
        return fallback()
    loc, scale = _GAUSSIAN_PARAMS[c["kind"]](c, parents)
    loc = loc.reshape(-1, 1, 1) if loc.dim() < 3 else loc  # _to_3d (:52-64)
    scale = scale.reshape(-1, 1, 1) if scale.dim() < 3 else scale
    if loc.shape[0] == 1 and b > 1:
        loc = loc.expand(b, -1, -1)
    if scale.shape[0] == 1 and b > 1:
        scale = scale.expand(b, -1, -1)
    scale = torch.nan_to_num(scale, nan=min_scale, posinf=min_scale, neginf=min_scale).abs().clamp_min(min_scale)
    loc, scale = loc[:, :1, :], scale[:, :1, :]
    z = torch.linspace(-stddevs, stddevs, n_samples, dtype=torch.float32).view(1, n_samples, 1)  # :172-176
    samples = loc + scale * z
    log_pdf = -0.5 * (z**2 + 2.0 * torch.log(scale) + math.log(2.0 * math.pi)).sum(dim=-1)
    out = (torch.exp(log_pdf), samples)
    return (*out, {"exact": True}) if return_info else out


def categorical_exact(spec, query, n_samples: int = 512, noise=None, return_info: bool = False):
    query = _norm_query(query)
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    fixed = _fixed_values(query, st, torch.float32, clamp_obs=True)
    t = st.target_idx
    c = spec["cpds"][st.topo[t]]

    def fallback():
        out = likelihood_weighting(spec, query, n_samples, noise=noise)
        return (*out, {"exact": False}) if return_info else out

    if fixed[t] is not None:  # :101-104
        out = (torch.ones(b, 1), fixed[t].unsqueeze(1).expand(b, 1, -1))
        return (*out, {"exact": True}) if return_info else out
    ok, parents = _exact_parent_tensor(st, fixed, b)
    if not ok or c["kind"] not in ("softmax_nn",) + tuple(_TABLE_LOGITS):  # :106-121
        return fallback()
    if c["kind"] == "softmax_nn":
        logits = snn_logits(c, parents, b, 1)  # includes the temperature and the root_ready branch (:48-71)
        support = c["sample_values"][0]
    else:
        logits = _TABLE_LOGITS[c["kind"]](c, parents)
        support = c["class_values"][0]
    probs = torch.softmax(logits, dim=-1)
    if probs.dim() == 4:
        probs = probs.reshape(probs.shape[0], -1, probs.shape[-1]) if parents is None else probs.squeeze(1)
    if probs.dim() != 3 or probs.shape[1] != 1 or c["output_dim"] != 1:  # :77-80, :123-124
        return fallback()
    probs = probs[:, 0, :]
    if probs.shape[0] == 1 and b > 1:
        probs = probs.expand(b, -1)
    out = (probs, support.to(dtype=probs.dtype).view(1, -1, 1).expand(b, -1, 1))
    return (*out, {"exact": True}) if return_info else out


# --------------------------------------------------------------------------------------
# importance sampling (vbn/inference/importance_sampling.py:24-93)
# --------------------------------------------------------------------------------------


def importance_sampling(spec, query, n_samples: int, noise=None, ess_threshold: float = 0.1,
                        return_info: bool = False):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    dtype = torch.float32
    samples = torch.zeros(b, n_samples, st.total_dim, dtype=dtype)
    logw = torch.zeros(b, n_samples, dtype=dtype)
    fixed = _fixed_values(query, st, dtype, clamp_obs=False)

    def sample_node(c, node, parent_tensor):
        if b <= 1:
            return cpd_sample(c, parent_tensor, n_samples, noise, ("is", node))
        rows = []  # per-query-row independent draws (:37-54)
        for bi in range(b):
            p_i = None if parent_tensor is None else parent_tensor[bi : bi + 1]
            rows.append(cpd_sample(c, p_i, n_samples, noise, ("is", node)))
        return torch.cat(rows, dim=0)

    for idx, node in enumerate(st.topo):
        c = spec["cpds"][node]
        if fixed[idx] is not None:
            value = fixed[idx].unsqueeze(1).expand(b, n_samples, -1)
            samples[..., st.slices[idx]] = value
            if st.evidence_mask[idx]:
                logw = logw + cpd_log_prob(c, value, _gather_parents(samples, st, idx))
            continue
        samples[..., st.slices[idx]] = sample_node(c, node, _gather_parents(samples, st, idx))

    w = torch.softmax(logw, dim=1)
    ess = 1.0 / (w**2).sum(dim=1)  # :82-84
    threshold = max(1.0, ess_threshold * float(n_samples))
    fallback = bool(torch.any(ess < threshold))
    if fallback:
        w, target = likelihood_weighting(spec, query, n_samples, noise, scope="lw")  # :85-88
    else:
        target = samples[..., st.slices[st.target_idx]]
    if return_info:
        return w, target, {"ess": ess, "fallback": fallback, "logw": logw, "samples": samples}
    return w, target


# --------------------------------------------------------------------------------------
# ancestral sampling (vbn/sampling/ancestral.py:13-65)
# --------------------------------------------------------------------------------------


def ancestral_sample_tensor(spec, query, n_samples: int, noise=None, scope="anc"):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query) if query["target"] else None
    if st is None:
        q2 = dict(query)
        q2["target"] = spec["topo"][0]
        st = _State(spec, q2)
    samples = torch.zeros(b, n_samples, st.total_dim, dtype=torch.float32)
    fixed = _fixed_values(query, st, torch.float32, clamp_obs=False)
    for idx, node in enumerate(st.topo):
        if fixed[idx] is not None:
            samples[..., st.slices[idx]] = fixed[idx].unsqueeze(1).expand(b, n_samples, -1)
            continue
        samples[..., st.slices[idx]] = cpd_sample(
            spec["cpds"][node], _gather_parents(samples, st, idx), n_samples, noise, (scope, node)
        )
    return samples, st


def ancestral_sample(spec, query, n_samples: int, noise=None):
    samples, st = ancestral_sample_tensor(spec, query, n_samples, noise)
    out = {n: samples[..., st.slices[i]] for i, n in enumerate(st.topo)}
    return out[query["target"]] if query.get("target") else out


# --------------------------------------------------------------------------------------
# "lbp" (vbn/inference/lbp.py:11-70): a wrapper -- one IS / MCM pass, a damped re-normalisation of the
# weights until the largest change is below tol, else a fresh IS pass
# --------------------------------------------------------------------------------------


def lbp(spec, query, n_samples: int, noise=None, n_iters: int = 10, damping: float = 0.5, tol: float = 1e-4,
        fallback: str = "importance_sampling"):
    eps = 1e-12
    if fallback == "monte_carlo_marginalization":  # :43-47
        pdf, target = monte_carlo_marginalization(spec, query, n_samples, noise)
        weights = pdf / (pdf.sum(dim=-1, keepdim=True) + eps)
    else:
        weights, target = importance_sampling(spec, query, n_samples, noise)
    converged = False
    for _ in range(max(int(n_iters), 0)):  # :52-63
        w_new = torch.clamp(weights, min=eps)
        w_new = w_new / (w_new.sum(dim=-1, keepdim=True) + eps)
        msg = damping * w_new + (1.0 - damping) * weights
        msg = msg / (msg.sum(dim=-1, keepdim=True) + eps)
        delta = (msg - weights).abs().max().item()
        weights = msg
        if delta < tol:
            converged = True
            break
    if not converged:  # :65-66
        return importance_sampling(spec, query, n_samples, noise)
    return weights, target


# --------------------------------------------------------------------------------------
# Gibbs sampler (vbn/sampling/gibbs.py:23-92) -- SURVEY 8f row 4
# --------------------------------------------------------------------------------------


def spec_children(spec) -> Dict[str, List[str]]:
    """Children of every node in dag.edges() order (_core.py:84-93).  Specs written before the key existed
    fall back to (child in node order, parent in parent order), which is networkx's order for DAGs built
    with add_edges_from in that order."""
    if "children" in spec:
        return spec["children"]
    out: Dict[str, List[str]] = {n: [] for n in spec["nodes"]}
    for p in spec["nodes"]:
        for c in spec["nodes"]:
            if p in spec["parents"][c]:
                out[p].append(c)
    return out


def gibbs_sample(spec, query, n_samples: int, noise=None, burn_in: int = 10, n_steps: int = 1,
                 n_candidates: int = 8):
    """One chain per query.  Per step and latent node: n_candidates proposals from the node's own CPD given the
    current parents, scored by their own log-density plus the log-density of every child's current value with the
    proposal substituted, one of them drawn by softmax (gibbs.py:38-82).  The target is collected every
    max(n_steps, 1) steps after burn_in -- as views of the live state, so the returned [B, n, D] tensor holds the
    chain's final target value n times (reference behaviour, reproduced)."""
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    current, st = ancestral_sample_tensor(spec, query, 1, noise, scope="gibbs_init")  # :32
    fixed = _fixed_values(query, st, torch.float32, clamp_obs=False)
    children = spec_children(spec)
    latent = [i for i in range(len(st.topo)) if fixed[i] is None]
    total_steps = int(burn_in) + int(n_samples) * max(int(n_steps), 1)
    k = int(n_candidates)
    collected = []
    for step in range(total_steps):
        for idx in latent:
            node = st.topo[idx]
            c = spec["cpds"][node]
            parents = _gather_parents(current, st, idx)
            if parents is not None and parents.shape[1] != k:
                parents = parents.expand(b, k, -1)
            cand = cpd_sample(c, parents, k, noise, ("gibbs", node))
            score = cpd_log_prob(c, cand, parents)
            for child in children[node]:
                ci = st.node_to_idx[child]
                cval = current[..., st.slices[ci]].expand(b, k, -1)
                parts = [cand if p == idx else current[..., st.slices[p]].expand(b, k, -1) for p in st.parent_idx[ci]]
                score = score + cpd_log_prob(spec["cpds"][child], cval, torch.cat(parts, dim=-1) if parts else None)
            weights = torch.softmax(score, dim=1)
            choice = noise.multinomial(("gibbs", node, "choice"), weights)
            chosen = cand[torch.arange(b), choice]
            current[..., st.slices[idx]] = chosen.unsqueeze(1)
        if step >= burn_in and (step - burn_in) % max(int(n_steps), 1) == 0:
            # QUIRK kept on purpose: the reference appends a VIEW of `current` (gibbs.py:83-87), which the later
            # in-place updates keep rewriting -- every collected entry ends up equal to the FINAL state
            collected.append(current[..., st.slices[st.target_idx]])
    if not collected:
        return current[..., st.slices[st.target_idx]]
    return torch.cat(collected, dim=1)


# --------------------------------------------------------------------------------------
# monte-carlo marginalisation (vbn/inference/monte_carlo_marginalization.py:18-92)
# --------------------------------------------------------------------------------------


def monte_carlo_marginalization(spec, query, n_samples: int, noise=None):
    query = _norm_query(query)
    noise = noise or TorchNoise()
    b = infer_batch_size(query["evidence"], query["do"])
    st = _State(spec, query)
    dtype = torch.float32
    fixed = _fixed_values(query, st, dtype, clamp_obs=False)
    t = st.target_idx
    target_node = st.topo[t]
    c_t = spec["cpds"][target_node]
    pidx = st.parent_idx[t]

    if target_node in query["do"]:  # :33-37
        tv = fixed[t].unsqueeze(1).expand(b, n_samples, -1)
        return torch.ones(b, n_samples, dtype=dtype), tv

    if all(fixed[p] is not None for p in pidx):  # :39-58
        if pidx:
            ptensor = torch.cat(
                [fixed[p].unsqueeze(1).expand(b, n_samples, -1) for p in pidx], dim=-1
            )
        else:
            ptensor = None
        if fixed[t] is not None:
            ts = fixed[t].unsqueeze(1).expand(b, n_samples, -1)
        else:
            ts = cpd_sample(c_t, ptensor, n_samples, noise, ("mcm", target_node))
        return torch.exp(cpd_log_prob(c_t, ts, ptensor)), ts

    samples = torch.zeros(b, n_samples, st.total_dim, dtype=dtype)
    for idx, node in enumerate(st.topo):  # :60-78
        if fixed[idx] is not None:
            samples[..., st.slices[idx]] = fixed[idx].unsqueeze(1).expand(b, n_samples, -1)
            continue
        samples[..., st.slices[idx]] = cpd_sample(
            spec["cpds"][node], _gather_parents(samples, st, idx), n_samples, noise, ("mcm", node)
        )
    ts = samples[..., st.slices[t]]
    lp = cpd_log_prob(c_t, ts, _gather_parents(samples, st, t))  # :80-92
    return torch.exp(lp), ts


# --------------------------------------------------------------------------------------
# closed-form linear-Gaussian posterior (builder-supplied check)
# --------------------------------------------------------------------------------------


# ---------------------------------------------------------------------------------------------
# Posterior summaries of the benchmark adapter, restated with numpy float64 like
# the reference (benchmarking/models/vbn.py).
# ---------------------------------------------------------------------------------------------
def normalize_probs(hist):
    """benchmarking/models/vbn.py:116-121."""
    import numpy as np

    arr = np.asarray(list(hist), dtype=float)
    total = float(arr.sum())
    if not math.isfinite(total) or total <= 0:
        return (np.ones_like(arr) / len(arr)).tolist()
    return (arr / total).tolist()


def estimate_discrete_posterior_batch(samples: torch.Tensor, weights: torch.Tensor, k: int):
    """benchmarking/models/vbn.py:202-242: per query, hist[round(x)] += w over finite weights with the
    class in range; Python round() is round-half-to-even (= numpy rint); float64 accumulation."""
    import numpy as np

    if samples.dim() == 3:
        samples = samples[:, :, 0]
    if samples.dim() != 2:
        raise ValueError(f"Expected samples with 2D shape, got {tuple(samples.shape)}")
    if weights.dim() != 2:
        raise ValueError(f"Expected weights with 2D shape, got {tuple(weights.shape)}")
    if samples.shape[0] != weights.shape[0]:
        raise ValueError("Samples/weights batch size mismatch")
    out = []
    for b in range(samples.shape[0]):
        vals = samples[b].detach().cpu().numpy().reshape(-1).astype(float)
        wts = weights[b].detach().cpu().numpy().reshape(-1).astype(float)
        idx = np.rint(vals)
        ok = np.isfinite(wts) & (idx >= 0) & (idx < k)
        hist = np.zeros(int(k), dtype=float)
        np.add.at(hist, idx[ok].astype(int), wts[ok])
        out.append(normalize_probs(hist))
    return out


def continuous_from_samples(samples, weights=None) -> dict:
    """benchmarking/models/vbn.py:365-423 (_extract_samples_1d + _continuous_from_samples)."""
    import numpy as np

    arr = np.asarray(samples.detach().cpu().numpy() if isinstance(samples, torch.Tensor) else samples)
    if arr.ndim == 0:
        arr = arr.reshape(1)
    if arr.ndim == 3:
        arr = arr.reshape(-1, arr.shape[-1])
    if arr.ndim == 2:
        if arr.shape[1] != 1:
            raise ValueError("Multivariate continuous targets are unsupported")
        arr = arr[:, 0]
    vals = arr.astype(float)
    if vals.size == 0:
        raise ValueError("No samples returned for continuous target")
    wts = None
    if weights is not None:
        w = np.asarray(weights.detach().cpu().numpy() if isinstance(weights, torch.Tensor) else weights).reshape(-1)
        if w.size == vals.size and np.isfinite(w).any():
            wts = w.astype(float)
    mean = std = None
    if wts is not None:
        wts = np.clip(wts, 0.0, np.inf)
        total = float(wts.sum())
        if total > 0:
            wts = wts / total
            mean = float(np.sum(wts * vals))
            std = math.sqrt(float(np.sum(wts * (vals - mean) ** 2)))
    if mean is None:
        mean, std = float(np.mean(vals)), float(np.std(vals, ddof=0))
    keep = [float(x) for x in vals[: min(int(vals.size), 2048)]]
    if math.isfinite(mean) and math.isfinite(std):
        return {"format": "normal_params", "mean": mean, "std": std, "n_samples": int(vals.size), "samples": keep}
    return {"format": "samples_1d", "samples": keep, "n_samples": int(vals.size)}


def lg_exact_posterior(spec, target: str, evidence: Dict[str, torch.Tensor],
                       do: Optional[Dict[str, torch.Tensor]] = None):
    """Exact N(mean, var) of ``target`` given evidence for an all-linear_gaussian DAG with
    scalar nodes.  float64.  Returns (mean[B], var[B])."""
    do = do or {}
    topo = list(spec["topo"])
    n = len(topo)
    n2i = {k: i for i, k in enumerate(topo)}
    b = infer_batch_size(evidence, do)
    A = torch.zeros(n, n, dtype=torch.float64)
    bias = torch.zeros(b, n, dtype=torch.float64)
    d = torch.zeros(n, dtype=torch.float64)
    for node in topo:
        c = spec["cpds"][node]
        assert c["kind"] == "linear_gaussian" and c["output_dim"] == 1
        i = n2i[node]
        if node in do:
            bias[:, i] = ensure_2d(do[node]).double()[:, 0]
            continue
        for j, p in enumerate(spec["parents"][node]):
            A[i, n2i[p]] = c["weight"][j, 0].double()
        bias[:, i] = c["bias"][0].double()
        d[i] = c["var"][0].double().clamp(min=float(c["min_scale"]) ** 2)
    M = torch.linalg.inv(torch.eye(n, dtype=torch.float64) - A)
    mu = bias @ M.T
    Sigma = M @ torch.diag(d) @ M.T
    t = n2i[target]
    e_idx = [n2i[k] for k in evidence]
    if not e_idx:
        return mu[:, t], Sigma[t, t].expand(b)
    e_val = torch.cat([ensure_2d(evidence[k]).double() for k in evidence], dim=1)
    S_ee = Sigma[e_idx][:, e_idx]
    S_te = Sigma[t, e_idx]
    gain = torch.linalg.solve(S_ee, S_te)
    mean = mu[:, t] + (e_val - mu[:, e_idx]) @ gain
    var = Sigma[t, t] - S_te @ gain
    return mean, var.expand(b)


# --------------------------------------------------------------------------------------
# spec extraction from a live reference model (used only where /root/reference exists)
# --------------------------------------------------------------------------------------


def _layers_of(net) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    return [
        (m.weight.detach().clone(), m.bias.detach().clone())
        for m in net
        if hasattr(m, "weight") and hasattr(m, "bias")
    ]


def cpd_spec_from_reference(cpd) -> dict:
    """Read a reference CPD object's parameters/buffers into the oracle's spec dict."""
    name = type(cpd).__name__
    base = {"input_dim": int(cpd.input_dim), "output_dim": int(cpd.output_dim)}
    g = lambda t: t.detach().clone()
    if name == "LinearGaussianCPD":
        return {**base, "kind": "linear_gaussian", "min_scale": float(cpd.min_scale),
                "weight": g(cpd._weight), "bias": g(cpd._bias), "var": g(cpd._var)}
    if name == "GaussianNNCPD":
        out = {**base, "kind": "gaussian_nn", "min_scale": float(cpd.min_scale),
               "activation": cpd.activation, "mean_x": g(cpd.mean_x), "std_x": g(cpd.std_x),
               "mean_y": g(cpd.mean_y), "std_y": g(cpd.std_y)}
        if cpd.input_dim == 0:
            out.update(loc=g(cpd._loc), log_scale=g(cpd._log_scale))
        else:
            out["layers"] = _layers_of(cpd.net)
        return out
    if name == "RFFGaussianCPD":
        return {**base, "kind": "rff_gaussian", "min_scale": float(cpd.min_scale), "n_features": int(cpd.n_features),
                "mean_x": g(cpd.mean_x), "std_x": g(cpd.std_x), "mean_y": g(cpd.mean_y), "std_y": g(cpd.std_y),
                "rff_w": g(cpd._rff_w), "rff_b": g(cpd._rff_b), "coef": g(cpd._coef), "bias": g(cpd._bias),
                "var": g(cpd._var), "stats_ready": bool(cpd._stats_ready.item())}
    if name == "MDNCPD":
        out = {**base, "kind": "mdn", "min_scale": float(cpd.min_scale),
               "activation": cpd.activation, "n_components": int(cpd.n_components)}
        if cpd.input_dim == 0:
            out.update(logits=g(cpd._logits), loc=g(cpd._loc), log_scale=g(cpd._log_scale))
        else:
            out["layers"] = _layers_of(cpd.net)
        return out
    if name == "SoftmaxNNCPD":
        out = {**base, "kind": "softmax_nn", "n_classes": int(cpd.n_classes),
               "activation": cpd.activation, "temperature": float(cpd.temperature),
               "min_bin_width": float(cpd.min_bin_width), "within_bin": cpd.within_bin,
               "within_bin_scale": float(cpd.within_bin_scale),
               "within_bin_clip": bool(cpd.within_bin_clip),
               "bin_edges": g(cpd._bin_edges), "class_values": g(cpd._class_values),
               "sample_values": g(cpd._sample_values), "is_discrete": g(cpd._is_discrete),
               "bins_ready": bool(cpd._bins_ready.item()),
               "root_ready": bool(cpd._root_ready.item()),
               "root_log_probs": g(cpd._root_log_probs)}
        if cpd.input_dim == 0:
            out["logits"] = g(cpd._logits)
        else:
            out["layers"] = _layers_of(cpd.net)
        return out
    if name == "KDECPD":
        return {**base, "kind": "kde", "bandwidth": float(cpd.bandwidth),
                "parent_bandwidth": float(cpd.parent_bandwidth), "min_scale": float(cpd.min_scale),
                "parents": None if cpd._parents is None else g(cpd._parents),
                "targets": None if cpd._targets is None else g(cpd._targets)}
    if name == "CategoricalTableCPD":
        return {**base, "kind": "categorical_table", "n_classes": int(cpd.n_classes),
                "counts": g(cpd._counts), "class_values": g(cpd._class_values), "class_mask": g(cpd._class_mask),
                "parent_values": [g(v) for v in (cpd._parent_values or [])],
                "parent_strides": [int(v) for v in (cpd._parent_strides or [])]}
    if name == "CategoricalEmbeddedSoftmaxCPD":
        root = cpd.input_dim == 0
        return {**base, "kind": "categorical_embedded_softmax", "n_classes": int(cpd.n_classes),
                "activation": cpd.activation, "stats_ready": bool(cpd._stats_ready.item()),
                "class_values": g(cpd._class_values), "class_mask": g(cpd._class_mask),
                "sample_values": g(cpd._sample_values),
                "parent_values": [g(v) for v in (cpd._parent_values or [])],
                "logits": g(cpd._logits) if root else None,
                "embeddings": None if root else [g(e.weight) for e in cpd.embeddings],
                "layers": None if root else _layers_of(cpd.net)}
    raise ValueError(f"CPD type '{name}' is outside the hot-path scope")


def spec_from_reference(vbn) -> dict:
    """Model spec of a fitted reference ``VBN`` (vbn/vbn.py:184; dags.py:24-45)."""
    nodes = list(vbn.dag.nodes())
    return {
        "nodes": nodes,
        "parents": {n: list(vbn.dag.parents(n)) for n in nodes},
        "topo": list(vbn.dag.topological_order()),
        "cpds": {n: cpd_spec_from_reference(vbn.nodes[n]) for n in nodes},
        # children in dag.edges() order (what _core.py:84-93 iterates; only the Gibbs sampler reads it)
        "children": {n: [c for p, c in vbn.dag.edges() if p == n] for n in nodes},
    }
