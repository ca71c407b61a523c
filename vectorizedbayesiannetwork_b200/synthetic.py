"""Seeded synthetic models for the BASELINE.json configs (spec dicts of CPU tensors, the same
format the oracle and ``VBN.from_spec`` read).  No reference code is involved: parameters are
random-initialised with the reference's architectures (nn.Linear default init for the MLPs), as
the bench contract allows when fitted checkpoints cannot travel to the GPU box."""
from __future__ import annotations

import math
import random
from typing import Dict, List, Tuple

import torch


def _linear_init(gen: torch.Generator, out_dim: int, in_dim: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.nn.Linear default init: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    bound = 1.0 / math.sqrt(max(in_dim, 1))
    w = (torch.rand(out_dim, in_dim, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_dim, generator=gen) * 2 - 1) * bound
    return w, b


def mlp_layers(gen, in_dim: int, hidden: Tuple[int, ...], out_dim: int):
    dims = [in_dim, *hidden, out_dim]
    return [_linear_init(gen, dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def lg_cpd(gen, dp: int, d: int = 1, weight_scale: float = 0.5, var: float = 0.25) -> dict:
    w = torch.randn(dp, d, generator=gen) * (weight_scale / math.sqrt(max(dp, 1)))
    return {"kind": "linear_gaussian", "input_dim": dp, "output_dim": d, "min_scale": 1e-4,
            "weight": w, "bias": 0.1 * torch.randn(d, generator=gen), "var": torch.full((d,), var)}


def mdn_cpd(gen, dp: int, d: int = 1, k: int = 3, hidden=(32, 32)) -> dict:
    c = {"kind": "mdn", "input_dim": dp, "output_dim": d, "min_scale": 1e-4, "activation": "relu",
         "n_components": k}
    if dp == 0:
        c.update(logits=0.3 * torch.randn(k, generator=gen), loc=torch.randn(k, d, generator=gen),
                 log_scale=-0.5 + 0.2 * torch.randn(k, d, generator=gen))
    else:
        c["layers"] = mlp_layers(gen, dp, tuple(hidden), k * (2 * d + 1))
    return c


def random_dag_lg_mdn(n_nodes: int = 1000, window: int = 20, max_parents: int = 3, seed: int = 0,
                      k: int = 3) -> dict:
    """BASELINE config 5: node i draws k~U{0..max_parents} parents from the previous ``window``
    nodes (random.Random(seed)); even nodes linear_gaussian, odd nodes mdn(K, hidden [32,32])."""
    rng = random.Random(seed)
    gen = torch.Generator().manual_seed(seed)
    names = [f"n{i}" for i in range(n_nodes)]
    parents: Dict[str, List[str]] = {}
    cpds: Dict[str, dict] = {}
    for i, name in enumerate(names):
        lo = max(0, i - window)
        kk = min(rng.randint(0, max_parents), i - lo)
        ps = sorted(rng.sample(range(lo, i), kk)) if kk else []
        parents[name] = [names[p] for p in ps]
        cpds[name] = lg_cpd(gen, len(ps)) if i % 2 == 0 else mdn_cpd(gen, len(ps), k=k)
    return {"nodes": names, "parents": parents, "topo": list(names), "cpds": cpds}


def lg_chain(n_nodes: int = 50, slope: float = 1.0, bias: float = 0.1, var: float = 0.25) -> dict:
    """BASELINE config 2: x0 -> ... -> x49, all linear_gaussian (x_{i+1} = x_i + 0.1 + 0.5 eps)."""
    names = [f"x{i}" for i in range(n_nodes)]
    parents = {names[0]: []}
    cpds = {names[0]: {"kind": "linear_gaussian", "input_dim": 0, "output_dim": 1, "min_scale": 1e-4,
                       "weight": torch.zeros(0, 1), "bias": torch.zeros(1), "var": torch.ones(1)}}
    for a, b in zip(names[:-1], names[1:]):
        parents[b] = [a]
        cpds[b] = {"kind": "linear_gaussian", "input_dim": 1, "output_dim": 1, "min_scale": 1e-4,
                   "weight": torch.full((1, 1), slope), "bias": torch.full((1,), bias),
                   "var": torch.full((1,), var)}
    return {"nodes": names, "parents": parents, "topo": list(names), "cpds": cpds}


# ALARM structure (bnlearn, 37 nodes / 46 arcs): NODE(cardinality) <- parents  
ALARM = {
    "HISTORY": (2, ["LVFAILURE"]), "CVP": (3, ["LVEDVOLUME"]), "PCWP": (3, ["LVEDVOLUME"]),
    "HYPOVOLEMIA": (2, []), "LVEDVOLUME": (3, ["HYPOVOLEMIA", "LVFAILURE"]), "LVFAILURE": (2, []),
    "STROKEVOLUME": (3, ["HYPOVOLEMIA", "LVFAILURE"]), "ERRLOWOUTPUT": (2, []),
    "HRBP": (3, ["ERRLOWOUTPUT", "HR"]), "HREKG": (3, ["ERRCAUTER", "HR"]), "ERRCAUTER": (2, []),
    "HRSAT": (3, ["ERRCAUTER", "HR"]), "INSUFFANESTH": (2, []), "ANAPHYLAXIS": (2, []),
    "TPR": (3, ["ANAPHYLAXIS"]), "EXPCO2": (4, ["ARTCO2", "VENTLUNG"]), "KINKEDTUBE": (2, []),
    "MINVOL": (4, ["INTUBATION", "VENTLUNG"]), "FIO2": (2, []), "PVSAT": (3, ["FIO2", "VENTALV"]),
    "SAO2": (3, ["PVSAT", "SHUNT"]), "PAP": (3, ["PULMEMBOLUS"]), "PULMEMBOLUS": (2, []),
    "SHUNT": (2, ["INTUBATION", "PULMEMBOLUS"]), "INTUBATION": (3, []),
    "PRESS": (4, ["INTUBATION", "KINKEDTUBE", "VENTTUBE"]), "DISCONNECT": (2, []), "MINVOLSET": (3, []),
    "VENTMACH": (4, ["MINVOLSET"]), "VENTTUBE": (4, ["DISCONNECT", "VENTMACH"]),
    "VENTLUNG": (4, ["INTUBATION", "KINKEDTUBE", "VENTTUBE"]), "VENTALV": (4, ["INTUBATION", "VENTLUNG"]),
    "ARTCO2": (3, ["VENTALV"]), "CATECHOL": (2, ["ARTCO2", "INSUFFANESTH", "SAO2", "TPR"]),
    "HR": (3, ["CATECHOL"]), "CO": (3, ["HR", "STROKEVOLUME"]), "BP": (3, ["CO", "TPR"]),
}


def _toposort(nodes: List[str], parents: Dict[str, List[str]]) -> List[str]:
    done, out = set(), []

    def visit(n):
        if n in done:
            return
        for p in parents[n]:
            visit(p)
        done.add(n)
        out.append(n)

    for n in nodes:
        visit(n)
    return out


def alarm_softmax(seed: int = 0, hidden=(32, 32)) -> dict:
    """BASELINE config 3: ALARM with softmax_nn CPDs in discrete mode (n_classes = cardinality,
    class values 0..k-1 as floats, like benchmarking/models/vbn.py:169-176,219)."""
    gen = torch.Generator().manual_seed(seed)
    nodes = list(ALARM)
    parents = {n: list(ALARM[n][1]) for n in nodes}
    cpds = {}
    for n in nodes:
        card, ps = ALARM[n]
        dp = len(ps)
        vals = torch.arange(card, dtype=torch.float32).view(1, card)
        edges = torch.linspace(-0.5, card - 0.5, card + 1).view(1, card + 1)
        c = {"kind": "softmax_nn", "input_dim": dp, "output_dim": 1, "n_classes": card,
             "activation": "relu", "temperature": 1.0, "min_bin_width": 1e-12, "within_bin": "triangular",
             "within_bin_scale": 0.25, "within_bin_clip": False, "bin_edges": edges, "class_values": vals,
             "sample_values": vals.clone(), "is_discrete": torch.ones(1, dtype=torch.bool),
             "bins_ready": True, "root_ready": dp == 0,
             "root_log_probs": torch.log_softmax(torch.randn(1, card, generator=gen), dim=-1)}
        if dp == 0:
            c["logits"] = torch.zeros(1, card)
        else:
            layers = mlp_layers(gen, dp, tuple(hidden), card)
            w, b = layers[-1]
            layers[-1] = (w * 4.0, b)  # sharper conditionals than a fresh init
            c["layers"] = layers
        cpds[n] = c
    return {"nodes": nodes, "parents": parents, "topo": _toposort(nodes, parents), "cpds": cpds}


def kde_pair(n_points: int = 200_000, seed: int = 0) -> dict:
    """BASELINE config 4: p -> y with a conditional Gaussian KDE over ``n_points`` stored points
    (p~N(0,1), y=sin(p)+0.1 eps), bandwidth = parent_bandwidth = 0.5."""
    gen = torch.Generator().manual_seed(seed)
    p = torch.randn(n_points, 1, generator=gen)
    y = torch.sin(p) + 0.1 * torch.randn(n_points, 1, generator=gen)
    cpds = {
        "p": {"kind": "linear_gaussian", "input_dim": 0, "output_dim": 1, "min_scale": 1e-4,
              "weight": torch.zeros(0, 1), "bias": torch.zeros(1), "var": torch.ones(1)},
        "y": {"kind": "kde", "input_dim": 1, "output_dim": 1, "bandwidth": 0.5, "parent_bandwidth": 0.5,
              "min_scale": 1e-4, "parents": p, "targets": y},
    }
    return {"nodes": ["p", "y"], "parents": {"p": [], "y": ["p"]}, "topo": ["p", "y"], "cpds": cpds}
