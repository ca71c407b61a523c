"""User-facing per-node handle, mirror of vbn/core/cpd_handle.py:121-428 for the hot-path
methods: sample / log_prob / pdf / forward / conditional / conditional_samples /
conditional_log_prob / conditional_pdf / conditional_mean_std.  All numerics run on the GPU via
the CPD objects; outputs are detached tensors like the reference."""
from __future__ import annotations

from typing import Optional

import torch

from .core import CPDOutput, ensure_2d, ensure_tensor
from .cpds import TABLE_KINDS, wrap_cpd


def _to_serializable(obj, *, max_tensor_elems: int = 2048):
    """vbn/core/utils.py:102-128."""
    if obj is None or isinstance(obj, (str, int, float, bool)):
        return obj
    if isinstance(obj, (torch.device, torch.dtype)):
        return str(obj)
    if isinstance(obj, torch.Tensor):
        t = obj.detach()
        if int(t.numel()) <= max_tensor_elems:
            return t.cpu().tolist()
        return {"type": "tensor", "shape": list(t.shape), "dtype": str(t.dtype),
                "device": str(t.device), "numel": int(t.numel())}
    if isinstance(obj, dict):
        return {str(k): _to_serializable(v, max_tensor_elems=max_tensor_elems) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_to_serializable(v, max_tensor_elems=max_tensor_elems) for v in obj]
    return str(obj)


class CPDHandle:
    def __init__(self, vbn, node: str) -> None:
        if node not in vbn.nodes:
            raise ValueError(f"Unknown node '{node}'.")
        self._vbn = vbn
        self._node = node
        self._cpd = wrap_cpd(vbn.nodes[node], vbn.device)
        self._parents = list(vbn.dag.parents(node))

    node = property(lambda self: self._node)
    name = property(lambda self: self._node)
    cpd = property(lambda self: self._cpd)
    cpd_name = property(lambda self: self._cpd.kind)
    cpd_type = property(lambda self: type(self._cpd).__name__)
    parents = property(lambda self: list(self._parents))
    device = property(lambda self: self._vbn.device)
    x_dim = property(lambda self: int(self._cpd.output_dim))
    output_dim = x_dim
    parents_dim = property(lambda self: int(self._cpd.input_dim))
    input_dim = parents_dim

    def _parents_tensor(self, parents) -> Optional[torch.Tensor]:
        """vbn/core/cpd_handle.py:195-243."""
        if self.parents_dim == 0:
            if parents is None or (isinstance(parents, dict) and not parents):
                return None
            if isinstance(parents, torch.Tensor):
                t = ensure_tensor(parents, device=self._vbn.device)
                if t.dim() == 2 and t.shape[-1] == 0:
                    return t
            raise ValueError(f"Node '{self._node}' has no parents.")
        if parents is None:
            raise ValueError(f"Parents required for node '{self._node}'.")
        if isinstance(parents, dict):
            tensors = []
            for parent in self._parents:
                if parent not in parents:
                    raise ValueError(f"Missing parent '{parent}' for node '{self._node}'.")
                tensors.append(ensure_2d(ensure_tensor(parents[parent], device=self._vbn.device)))
            t = torch.cat(tensors, dim=-1)
            if t.shape[-1] != self.parents_dim:
                raise ValueError(f"Expected parents_dim {self.parents_dim}, got {t.shape[-1]}")
            return t
        if isinstance(parents, torch.Tensor):
            t = ensure_tensor(parents, device=self._vbn.device)
            if t.dim() == 1:
                t = ensure_2d(t)
            if t.dim() not in (2, 3):
                raise ValueError(f"Expected parents with 2D or 3D shape, got {tuple(t.shape)}")
            if t.shape[-1] != self.parents_dim:
                raise ValueError(f"Expected parents_dim {self.parents_dim}, got {t.shape[-1]}")
            return t
        raise TypeError("parents must be a tensor or dict")

    def _x_tensor(self, x) -> torch.Tensor:
        t = ensure_tensor(x, device=self._vbn.device)
        if t.dim() == 1:
            t = ensure_2d(t)
        if t.dim() not in (2, 3):
            raise ValueError(f"Expected x with 2D or 3D shape, got {tuple(t.shape)}")
        return t

    def sample(self, parents, n_samples: int) -> torch.Tensor:
        return self._cpd.sample(self._parents_tensor(parents), int(n_samples)).detach()

    def log_prob(self, x, parents) -> torch.Tensor:
        return self._cpd.log_prob(self._x_tensor(x), self._parents_tensor(parents)).detach()

    def pdf(self, x, parents) -> torch.Tensor:
        return torch.exp(self.log_prob(x, parents))

    def forward(self, parents, n_samples: int) -> CPDOutput:
        out = self._cpd.forward(self._parents_tensor(parents), int(n_samples))
        return CPDOutput(samples=out.samples.detach(), log_prob=out.log_prob.detach(), pdf=out.pdf.detach())

    @property
    def is_fitted(self) -> bool:
        """vbn/core/cpd_handle.py:154-171: the fitted flags the reference probes (`_targets`, `_bins_ready`,
        `_stats_ready`), else "has parameters" -- which holds for every parameter holder here."""
        for flag in ("_targets", "_bins_ready", "_stats_ready"):
            if hasattr(self._cpd, flag):
                value = getattr(self._cpd, flag)
                if isinstance(value, torch.Tensor):
                    value = value.numel() > 0 if value.dim() > 0 else bool(value.item())
                if value is not None and bool(value):
                    return True
                if flag != "_targets":
                    return False
        return True

    def state_dict(self) -> dict:
        """vbn/core/cpd_handle.py:308-311: the tensors of the CPD's spec (what the packed device blob is built from)."""
        out = {}

        def walk(prefix, v):
            if isinstance(v, torch.Tensor):
                out[prefix] = v
            elif isinstance(v, (list, tuple)):
                for i, x in enumerate(v):
                    walk(f"{prefix}.{i}", x)

        for k, v in self._cpd.to_spec().items():
            walk(k, v)
        return out

    def export_config(self) -> dict:
        """vbn/core/cpd_handle.py:288-306: the non-tensor part of the spec plays the role of init_kwargs."""
        spec = self._cpd.to_spec()
        init = {k: v for k, v in spec.items() if not isinstance(v, (torch.Tensor, list, tuple)) and k != "kind"}
        return {"node": self.node, "parents": self.parents, "cpd_name": self.cpd_name, "cpd_type": self.cpd_type,
                "init_kwargs": _to_serializable(init), "extra_state": None}

    def clone_cpd(self, detach: bool = True):
        """vbn/core/cpd_handle.py:313-346: an independent copy of the CPD (parameters cloned; nothing here tracks
        gradients, so ``detach`` has nothing left to do)."""
        from .cpds import cpd_from_spec

        def cp(v):
            if isinstance(v, torch.Tensor):
                return v.detach().clone()
            if isinstance(v, (list, tuple)):
                return type(v)(cp(x) for x in v)
            return v

        return cpd_from_spec({k: cp(v) for k, v in self._cpd.to_spec().items()}, device=self._cpd.device)

    def summary(self) -> dict:
        return {"node": self.node, "parents": self.parents, "cpd_name": self.cpd_name,
                "cpd_type": self.cpd_type, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "device": str(self.device), "is_fitted": self.is_fitted}

    def conditional(self, parents, *, n_samples: int = 1024) -> dict:
        """vbn/core/cpd_handle.py:348-402: normal_params (linear_gaussian, gaussian_nn), mixture_params
        (mdn), categorical_probs (softmax_nn) from the GPU parameter read-out (VBN_F_OUT_PARAMS);
        empirical_samples for kde."""
        ptensor = self._parents_tensor(parents)
        base = {"node": self.node, "parents": self.parents, "cpd_name": self.cpd_name,
                "cpd_type": self.cpd_type, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "conditioning": _to_serializable(ptensor)}
        cpd, d = self._cpd, self.x_dim
        if cpd.kind in TABLE_KINDS:  # the CPD's parameters are the table: a host-side lookup
            rows = cpd.probs(ptensor)
            return {**base, "format": "categorical_probs", "probs": _to_serializable(rows),
                    "k": int(cpd.n_classes), "support": _to_serializable(cpd._sample_values)}
        if cpd.param_width() > 0:
            out = cpd.params(ptensor).detach()  # [B, S, width] from the GPU read-out
            if cpd.kind in ("linear_gaussian", "gaussian_nn", "rff_gaussian"):
                return {**base, "format": "normal_params", "mean": _to_serializable(out[..., :d]),
                        "std": _to_serializable(out[..., d:])}
            if cpd.kind == "mdn":
                k = cpd.n_components
                b, s_ = out.shape[0], out.shape[1]
                return {**base, "format": "mixture_params", "weights": _to_serializable(out[..., :k]),
                        "loc": _to_serializable(out[..., k:k + k * d].reshape(b, s_, k, d)),
                        "scale": _to_serializable(out[..., k + k * d:].reshape(b, s_, k, d))}
            if cpd.kind == "softmax_nn":
                b, s_ = out.shape[0], out.shape[1]
                return {**base, "format": "categorical_probs",
                        "probs": _to_serializable(out.reshape(b, s_, d, cpd.n_classes)),
                        "k": int(cpd.n_classes), "support": _to_serializable(cpd._sample_values)}
        samples = self._cpd.sample(ptensor, int(n_samples)).detach()
        return {**base, "format": "empirical_samples", "samples": _to_serializable(samples),
                "mean": _to_serializable(samples.mean(dim=1)),
                "std": _to_serializable(samples.std(dim=1, unbiased=False)), "n_samples": int(n_samples)}

    def conditional_samples(self, parents, n_samples: int = 1024) -> torch.Tensor:
        return self.sample(parents, n_samples)

    def conditional_log_prob(self, x, parents) -> torch.Tensor:
        return self.log_prob(x, parents)

    def conditional_pdf(self, x, parents) -> torch.Tensor:
        return self.pdf(x, parents)

    def conditional_mean_std(self, parents, n_samples: int = 1024) -> dict:
        samples = self._cpd.sample(self._parents_tensor(parents), int(n_samples)).detach()
        return {"format": "empirical_samples", "mean": samples.mean(dim=1),
                "std": samples.std(dim=1, unbiased=False)}
