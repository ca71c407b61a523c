"""TEST / BENCH INFRASTRUCTURE ONLY -- the UNMODIFIED reference package as the CPU arm.

``load_reference()`` imports ``vbn`` 0.3.0 from ``baseline/_ref`` (installed there from /root/reference with
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref``: git-ignored, but it travels to the GPU
box with the snapshot) or, in the build container, straight from /root/reference.

``reference_from_spec(spec)`` builds a reference ``VBN`` whose nodes are the reference's OWN CPD classes
(``vbn.core.registry.CPD_REGISTRY``) holding exactly the tensors of a model spec (the inverse of
``oracle.vbn_oracle.cpd_spec_from_reference``), so the reference arm of bench.py and the B200 arm evaluate the same
parameters.  Inference then runs through the reference's public API and stock code path:
``VBN.set_inference_method`` / ``VBN.infer_posterior`` (vbn/vbn.py:257-335, 474-481), ``VBN.get_cpd().log_prob``.
Nothing in ``vectorizedbayesiannetwork_b200/`` imports this module.
"""
from __future__ import annotations

import os
import sys
from typing import Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("VBN_REFERENCE_ROOT", "/root/reference"))


def reference_path() -> Optional[str]:
    for p in CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "vbn")):
            return p
    return None


def load_reference():
    """The reference package (module ``vbn``), or None when neither location exists."""
    p = reference_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    import vbn  # noqa: F401

    return vbn


def _set(t: torch.Tensor, value: torch.Tensor) -> None:
    with torch.no_grad():
        t.copy_(value.to(dtype=t.dtype, device=t.device).reshape(t.shape))


def _load_layers(net, layers) -> None:
    linears = [m for m in net if hasattr(m, "weight") and hasattr(m, "bias")]
    if len(linears) != len(layers):
        raise ValueError(f"MLP depth mismatch: reference {len(linears)}, spec {len(layers)}")
    for m, (w, b) in zip(linears, layers):
        _set(m.weight, w)
        _set(m.bias, b)


def reference_cpd(vbn_mod, c: dict, device="cpu"):
    """One reference CPD object carrying the parameters of spec entry ``c``."""
    from vbn.core.registry import CPD_REGISTRY

    dev = torch.device(device)
    kind = c["kind"]
    cls = CPD_REGISTRY[kind]
    dp, d = int(c["input_dim"]), int(c["output_dim"])
    hidden = tuple(int(w.shape[0]) for w, _ in c["layers"][:-1]) if c.get("layers") else (32, 32)
    if kind == "linear_gaussian":
        m = cls(input_dim=dp, output_dim=d, device=dev, min_scale=float(c["min_scale"]))
        _set(m._weight, c["weight"]); _set(m._bias, c["bias"]); _set(m._var, c["var"])
    elif kind == "gaussian_nn":
        m = cls(input_dim=dp, output_dim=d, device=dev, hidden_dims=hidden, activation=c["activation"],
                min_scale=float(c["min_scale"]))
        for name in ("mean_x", "std_x", "mean_y", "std_y"):
            _set(getattr(m, name), c[name])
        if dp == 0:
            _set(m._loc, c["loc"]); _set(m._log_scale, c["log_scale"])
        else:
            _load_layers(m.net, c["layers"])
    elif kind == "mdn":
        m = cls(input_dim=dp, output_dim=d, device=dev, n_components=int(c["n_components"]), hidden_dims=hidden,
                activation=c["activation"], min_scale=float(c["min_scale"]))
        if dp == 0:
            _set(m._logits, c["logits"]); _set(m._loc, c["loc"]); _set(m._log_scale, c["log_scale"])
        else:
            _load_layers(m.net, c["layers"])
    elif kind == "softmax_nn":
        m = cls(input_dim=dp, output_dim=d, device=dev, n_classes=int(c["n_classes"]), hidden_dims=hidden,
                activation=c["activation"], min_bin_width=float(c["min_bin_width"]), within_bin=c["within_bin"],
                within_bin_scale=float(c["within_bin_scale"]), within_bin_clip=bool(c["within_bin_clip"]))
        m.temperature = float(c["temperature"])
        _set(m._bin_edges, c["bin_edges"]); _set(m._class_values, c["class_values"])
        _set(m._sample_values, c["sample_values"]); _set(m._is_discrete, c["is_discrete"])
        _set(m._root_log_probs, c["root_log_probs"])
        m._bins_ready.fill_(bool(c["bins_ready"])); m._root_ready.fill_(bool(c["root_ready"]))
        if dp == 0:
            _set(m._logits, c["logits"])
        else:
            _load_layers(m.net, c["layers"])
    elif kind == "kde":
        n = int(c["targets"].shape[0])
        m = cls(input_dim=dp, output_dim=d, device=dev, bandwidth=float(c["bandwidth"]),
                parent_bandwidth=float(c["parent_bandwidth"]), max_points=n, min_scale=float(c["min_scale"]))
        m._parents = None if c.get("parents") is None else c["parents"].to(dev).float()
        m._targets = c["targets"].to(dev).float()
    else:
        raise ValueError(f"reference arm: CPD kind '{kind}' is not part of a BASELINE workload")
    m.eval()
    return m


def reference_from_spec(spec: dict, device="cpu"):
    """Reference ``VBN`` (vbn/vbn.py:184) over the spec's DAG with reference CPD modules holding the spec's tensors."""
    import networkx as nx

    vbn_mod = load_reference()
    if vbn_mod is None:
        raise RuntimeError("the reference package is not available (baseline/_ref or /root/reference)")
    g = nx.DiGraph()
    g.add_nodes_from(spec["topo"])  # insertion order fixes topological_order() like the spec's
    for node in spec["topo"]:
        for p in spec["parents"][node]:
            g.add_edge(p, node)
    model = vbn_mod.VBN(g, seed=0, device=device)
    for node in spec["topo"]:
        model.nodes[node] = reference_cpd(vbn_mod, spec["cpds"][node], device)
    return model
