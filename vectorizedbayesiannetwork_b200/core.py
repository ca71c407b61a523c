"""Host-side mirror of the reference's registry-driven API for the hot path
(vbn/core/{base,registry,dags,cpd_handle}.py and the inference/sampling part of vbn/vbn.py),
so the package is usable stand-alone on a box without the reference installed and so the
parity tests read like the reference's own tests.  Training, update policies, persistence,
plotting and the YAML config system are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import torch

from .cpds import BaseCPD, cpd_from_spec, wrap_cpd

# ---------------------------------------------------------------------------------------------
# vbn/core/base.py:11-25
# ---------------------------------------------------------------------------------------------


@dataclass
class CPDOutput:
    samples: torch.Tensor  # [B, S, Dx]
    log_prob: torch.Tensor  # [B, S]
    pdf: torch.Tensor  # [B, S]


@dataclass
class Query:
    target: str
    evidence: Dict[str, torch.Tensor]
    do: Dict[str, torch.Tensor] = field(default_factory=dict)


# ---------------------------------------------------------------------------------------------
# vbn/core/registry.py:7-43 (inference + sampling registries only)
# ---------------------------------------------------------------------------------------------
INFERENCE_REGISTRY: Dict[str, type] = {}
SAMPLING_REGISTRY: Dict[str, type] = {}


def _register(registry: Dict[str, type], name: str):
    key = name.lower().strip()

    def decorator(cls):
        if key in registry:
            raise ValueError(f"Duplicate registry key '{key}' for {cls.__name__}")
        registry[key] = cls
        return cls

    return decorator


def register_inference(name: str):
    return _register(INFERENCE_REGISTRY, name)


def register_sampling(name: str):
    return _register(SAMPLING_REGISTRY, name)


@dataclass
class ConfigItem:  # vbn/vbn.py:37-51
    name: str
    params: Dict
    kind: Optional[str] = None


def _default_config() -> SimpleNamespace:
    """Packaged YAML defaults of the hot-path methods (vbn/configs/inference/*.yaml,
    vbn/configs/sampling/ancestral.yaml)."""
    inference = SimpleNamespace(
        monte_carlo_marginalization=ConfigItem("monte_carlo_marginalization", {"n_samples": 1024}, "inference"),
        importance_sampling=ConfigItem("importance_sampling", {"n_samples": 1024}, "inference"),
        likelihood_weighting=ConfigItem(
            "likelihood_weighting", {"n_samples": 1024, "eps": 1e-12, "normalize": True}, "inference"),
    )
    sampling = SimpleNamespace(ancestral=ConfigItem("ancestral", {"n_samples": 512}, "sampling"))
    return SimpleNamespace(inference=inference, sampling=sampling)


# ---------------------------------------------------------------------------------------------
# vbn/core/utils.py:48-61, vbn/utils/__init__.py:46-61
# ---------------------------------------------------------------------------------------------


def ensure_tensor(x, device, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        if x.dtype == dtype and x.device == device:  # the common case on the small-call latency path
            return x
        return x.to(device=device, dtype=dtype)
    return torch.tensor(x, device=device, dtype=dtype)


def ensure_2d(x: torch.Tensor) -> torch.Tensor:
    if x.dim() == 1:
        return x.unsqueeze(-1)
    if x.dim() == 2:
        return x
    raise ValueError(f"Expected 1D or 2D tensor, got shape {tuple(x.shape)}")


def infer_batch_size(evidence, do=None) -> int:
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


# ---------------------------------------------------------------------------------------------
# vbn/core/dags.py:24-45
# ---------------------------------------------------------------------------------------------


class StaticDAG:
    """Accepts a networkx.DiGraph (like the reference) or an explicit (nodes, parents) pair."""

    def __init__(self, graph=None, *, nodes: Optional[List[str]] = None,
                 parents: Optional[Dict[str, List[str]]] = None, topo: Optional[List[str]] = None,
                 children: Optional[Dict[str, List[str]]] = None):
        if graph is not None:
            import networkx as nx

            if not nx.is_directed_acyclic_graph(graph):
                raise ValueError("DAG must be acyclic")
            self._nodes = list(graph.nodes)
            self._topo = list(nx.topological_sort(graph))
            self._parents = {n: list(graph.predecessors(n)) for n in graph.nodes}
            self._edges = list(graph.edges)
            self.graph = graph
        else:
            self._nodes = list(nodes)
            self._parents = {n: list(parents.get(n, [])) for n in self._nodes}
            if children is not None:  # dag.edges() order of the reference (what its Gibbs sampler iterates)
                self._edges = [(p, c) for p in self._nodes for c in children.get(p, [])]
            else:
                self._edges = [(p, n) for p in self._nodes for n in self._nodes if p in self._parents[n]]
            self._topo = list(topo) if topo is not None else self._toposort()
            self.graph = None

    def _toposort(self) -> List[str]:
        indeg = {n: len(self._parents[n]) for n in self._nodes}
        children: Dict[str, List[str]] = {n: [] for n in self._nodes}
        for p, c in self._edges:
            children[p].append(c)
        ready = [n for n in self._nodes if indeg[n] == 0]
        out: List[str] = []
        while ready:
            n = ready.pop(0)
            out.append(n)
            for c in children[n]:
                indeg[c] -= 1
                if indeg[c] == 0:
                    ready.append(c)
        if len(out) != len(self._nodes):
            raise ValueError("DAG must be acyclic")
        return out

    def nodes(self) -> List[str]:
        return list(self._nodes)

    def edges(self):
        return list(self._edges)

    def parents(self, node: str) -> List[str]:
        return self._parents.get(node, [])

    def topological_order(self) -> List[str]:
        return list(self._topo)


# ---------------------------------------------------------------------------------------------
# model view used by the inference classes: works on our VBN and on a reference vbn.VBN
# ---------------------------------------------------------------------------------------------


def _module_tensors(obj) -> list:
    sd = obj.state_dict() if hasattr(obj, "state_dict") else {}
    extra = obj.get_extra_state() if hasattr(obj, "get_extra_state") else None
    ts = [v for v in sd.values() if isinstance(v, torch.Tensor)]
    if isinstance(extra, dict):
        ts += [v for v in extra.values() if isinstance(v, torch.Tensor)]
    return ts


def _module_fingerprint(obj, memo=None, refresh: bool = True) -> tuple:
    """Identity + versions of a reference module's tensors.  In-place changes (optimizer steps, ``copy_``,
    ``load_state_dict``) bump ``_version``, ``.to()`` moves ``data_ptr``: both are seen by reading the tensors the
    module had last time (``memo``: ~2 us per module instead of ~17 us for a fresh ``state_dict()`` walk -- 15 ms per
    call on a 1000-node model).  Re-bound tensors (``m.weight = nn.Parameter(...)``, KDE refits) are caught by the
    full walk, which ``model_cpds`` forces every 16th call."""
    if isinstance(obj, BaseCPD):
        return (obj._uid, obj._version)
    ent = None if memo is None or refresh else memo.get(id(obj))
    if ent is None or ent[0] is not obj:
        ent = (obj, _module_tensors(obj))
        if memo is not None:
            memo[id(obj)] = ent
    return (id(obj), tuple((t.data_ptr(), t._version) for t in ent[1]))


def model_cpds(vbn) -> Dict[str, BaseCPD]:
    """Our CPD objects for every node of ``vbn`` (ours or the reference's), re-wrapped when a
    reference module's tensors changed (e.g. after ``update``)."""
    cache = vbn.__dict__.setdefault("_b200_cpd_cache", {})
    memo = vbn.__dict__.setdefault("_b200_tensor_memo", {})
    calls = vbn.__dict__["_b200_calls"] = vbn.__dict__.get("_b200_calls", 0) + 1
    refresh = calls % 16 == 1
    out: Dict[str, BaseCPD] = {}
    for node, obj in vbn.nodes.items():
        if isinstance(obj, BaseCPD):
            out[node] = obj
            continue
        fp = _module_fingerprint(obj, memo, refresh)
        hit = cache.get(node)
        if hit is None or hit[0] != fp:
            hit = (fp, wrap_cpd(obj, vbn.device))
            cache[node] = hit
        out[node] = hit[1]
    return out


# ---------------------------------------------------------------------------------------------
# VBN facade (vbn/vbn.py:184-335, 474-481, 570-642) -- inference/sampling side only
# ---------------------------------------------------------------------------------------------


class VBN:
    def __init__(self, dag, seed: Optional[int] = None, device=None) -> None:
        from .engine import require_cuda

        self.device = require_cuda(device)
        if seed is not None:
            torch.manual_seed(seed)
        self.seed = seed
        self.dag = dag if isinstance(dag, StaticDAG) else StaticDAG(dag)
        self.nodes: Dict[str, BaseCPD] = {}
        self.config = _default_config()
        self._inference = None
        self._sampling = None
        self._inference_config: Optional[Dict[str, Any]] = None
        self._sampling_config: Optional[Dict[str, Any]] = None

    # ----- construction from fitted parameters (fit itself is out of scope) -----------------
    @classmethod
    def from_spec(cls, spec: dict, device=None, seed: Optional[int] = None) -> "VBN":
        dag = StaticDAG(nodes=spec["nodes"], parents=spec["parents"], topo=spec.get("topo"),
                        children=spec.get("children"))
        model = cls(dag, seed=seed, device=device)
        model.nodes = {n: cpd_from_spec(spec["cpds"][n], model.device) for n in spec["nodes"]}
        return model

    @classmethod
    def from_reference(cls, ref_vbn, device=None) -> "VBN":
        """Snapshot of a fitted reference ``vbn.VBN`` (same DAG order, same parameters)."""
        nodes = list(ref_vbn.dag.nodes())
        dag = StaticDAG(nodes=nodes, parents={n: list(ref_vbn.dag.parents(n)) for n in nodes},
                        topo=list(ref_vbn.dag.topological_order()),
                        children={n: [c for p, c in ref_vbn.dag.edges() if p == n] for n in nodes})
        model = cls(dag, device=device)
        model.nodes = {n: wrap_cpd(ref_vbn.nodes[n], model.device) for n in nodes}
        return model

    def to_device(self, device) -> None:
        """vbn/vbn.py:621-631: move the model to another CUDA device.  The CPD objects only hold host-side
        parameters; their packed blobs and cached plans are per device, so they are dropped and rebuilt lazily."""
        from .engine import require_cuda

        self.device = require_cuda(device)
        for cpd in self.nodes.values():
            cpd.device = self.device
            cpd.invalidate()

    def set_cpd(self, node: str, cpd: BaseCPD) -> None:
        if node not in self.dag.nodes():
            raise ValueError(f"Unknown node '{node}'.")
        self.nodes[node] = cpd

    def to_spec(self) -> dict:
        nodes = self.dag.nodes()
        return {"nodes": nodes, "parents": {n: list(self.dag.parents(n)) for n in nodes},
                "topo": self.dag.topological_order(), "cpds": {n: self.nodes[n].to_spec() for n in nodes}}

    # ----- method selection (vbn/vbn.py:257-335) --------------------------------------------
    @staticmethod
    def _resolve(method, registry, what: str):
        if isinstance(method, dict):
            name = method.get("name") or method.get("method")
            if name is None:
                raise TypeError("method dict must include a 'name' field")
            if not isinstance(name, str):
                raise TypeError("method name must be a string")
            key = name.lower().strip()
            if key not in registry:
                raise ValueError(f"Unknown {what} method '{name}'. Available: {list(registry.keys())}")
            return key, {k: v for k, v in method.items() if k not in {"name", "method"}}
        if isinstance(method, ConfigItem) or (hasattr(method, "name") and hasattr(method, "params")):
            return method.name, dict(method.params)
        if isinstance(method, str):
            key = method.lower().strip()
            if key not in registry:
                raise ValueError(f"Unknown {what} method '{method}'. Available: {list(registry.keys())}")
            return key, {}
        return None, None

    def set_inference_method(self, method, **kwargs) -> None:
        name, base = self._resolve(method, INFERENCE_REGISTRY, "inference")
        if name is None:
            if callable(method) or hasattr(method, "infer_posterior"):
                self._inference = method
                self._inference_config = {"callable": True, "name": getattr(method, "__qualname__", str(method))}
                return
            raise TypeError("method must be a string, ConfigItem, or callable")
        params = {**base, **kwargs}
        self._inference = INFERENCE_REGISTRY[name](**params)
        self._inference_config = {"name": name, "params": params}

    def set_sampling_method(self, method, **kwargs) -> None:
        name, base = self._resolve(method, SAMPLING_REGISTRY, "sampling")
        if name is None:
            if callable(method) or hasattr(method, "sample"):
                self._sampling = method
                self._sampling_config = {"callable": True, "name": getattr(method, "__qualname__", str(method))}
                return
            raise TypeError("method must be a string, ConfigItem, or callable")
        params = {**base, **kwargs}
        self._sampling = SAMPLING_REGISTRY[name](**params)
        self._sampling_config = {"name": name, "params": params}

    # ----- inference / sampling (vbn/vbn.py:474-481, 570-618) ----------------------------------
    def infer_posterior(self, query, **kwargs):
        if self._inference is None:
            raise RuntimeError("Call set_inference_method(...) before infer_posterior().")
        q = self._normalize_query(query)
        out = self._inference.infer_posterior(self, q, **kwargs)
        if isinstance(out, dict):  # summary=True (extension): per-query posterior summary reduced on the device
            return {k: v.detach() for k, v in out.items()}
        pdf, samples = out
        return pdf.detach(), samples.detach()

    def sample(self, query, n_samples: int = 200, **kwargs):
        if self._sampling is None:
            raise RuntimeError("Call set_sampling_method(...) before sample().")
        q = self._normalize_query(query)
        samples = self._sampling.sample(self, q, n_samples=n_samples, **kwargs)
        if isinstance(samples, dict):
            return {k: v.detach() for k, v in samples.items()}
        return samples.detach()

    def _normalize_query(self, query) -> Query:
        if isinstance(query, Query) or (hasattr(query, "target") and hasattr(query, "evidence")):
            target = query.target
            evidence_src = query.evidence or {}
            do_src = getattr(query, "do", None) or {}
        else:
            if not isinstance(query, dict):
                raise TypeError("query must be a dict or Query")
            target = query.get("target") or query.get("target_feature")
            if target is None:
                raise ValueError("query must contain 'target'")
            evidence_src = query.get("evidence") or {}
            do_src = query.get("do") or {}
        evidence = {k: ensure_2d(ensure_tensor(v, device=self.device)) for k, v in evidence_src.items()}
        do = {k: ensure_2d(ensure_tensor(v, device=self.device)) for k, v in do_src.items()}
        nodes = set(self.dag.nodes())
        if target not in nodes:
            raise ValueError(f"Unknown target node '{target}'.")
        unknown = (set(evidence) | set(do)) - nodes
        if unknown:
            raise ValueError(f"Unknown query nodes: {sorted(unknown)}")
        overlap = set(evidence) & set(do)
        if overlap:
            raise ValueError(f"Nodes cannot be in both evidence and do: {sorted(overlap)}")
        infer_batch_size(evidence, do)
        return Query(target=target, evidence=evidence, do=do)

    # ----- posterior summaries (vbn/vbn.py:483-568) ------------------------------------------------
    def _posterior_stats(self, pdf: torch.Tensor, samples: torch.Tensor, *, eps: float = 1e-12):
        from .engine import posterior_stats

        return posterior_stats(pdf, samples, eps)

    posterior_stats = _posterior_stats

    @staticmethod
    def _broadcast_batch(a: torch.Tensor, b: torch.Tensor):
        if a.shape[0] == b.shape[0]:
            return a, b
        if a.shape[0] == 1:
            return a.expand(b.shape[0], *a.shape[1:]), b
        if b.shape[0] == 1:
            return a, b.expand(a.shape[0], *b.shape[1:])
        raise ValueError("Query and reference batch sizes must match, unless one of them is 1.")

    def infer_relative(self, query, reference_query=None, *, eps: float = 1e-12, **kwargs):
        q = self._normalize_query(query)
        if reference_query is None:
            reference_query = Query(target=q.target, evidence={}, do={})
        rq = self._normalize_query(reference_query)
        if rq.target != q.target:
            raise ValueError("query and reference_query must have the same target node.")
        qs = self._posterior_stats(*self.infer_posterior(q, **kwargs), eps=eps)
        rs = self._posterior_stats(*self.infer_posterior(rq, **kwargs), eps=eps)
        qm, rm = self._broadcast_batch(qs["mean"], rs["mean"])
        qsd, rsd = self._broadcast_batch(qs["std"], rs["std"])
        qe, re_ = self._broadcast_batch(qs["ess"], rs["ess"])
        dm, dsd = qm - rm, qsd - rsd
        return {
            "target": q.target,
            "query_stats": {"mean": qm, "std": qsd, "effective_sample_size": qe},
            "reference_stats": {"mean": rm, "std": rsd, "effective_sample_size": re_},
            "delta_mean": dm, "delta_std": dsd,
            "relative_mean_change": dm / rm.abs().clamp_min(eps),
            "relative_std_change": dsd / rsd.abs().clamp_min(eps),
        }

    # ----- CPD access (vbn/vbn.py:633-642) -------------------------------------------------------
    def cpd(self, node: str):
        from .cpd_handle import CPDHandle

        return CPDHandle(self, node)

    get_cpd = cpd

    def get_cpds(self):
        from .cpd_handle import CPDHandle

        return {node: CPDHandle(self, node) for node in self.dag.nodes()}
