"""Converts a RecordingNoise log (oracle/vbn_oracle.py) into the injection format the product
accepts: {scope: {node: {"eps": [Bn,S,D], "u": [Bn,S,D], "idx": [Bn,S] or [Bn,S,D]}}}."""
from __future__ import annotations

import torch


def to_injection(log, spec, n_samples: int):
    out = {}
    for key, lst in log.items():
        scope, node, kind = key
        if node == "__resample__":  # resampled_importance_sampling: one [B, S] index tensor per resampling event
            out.setdefault(scope, {})[node] = [x.to(torch.int32).reshape(-1, n_samples) for x in lst]
            continue
        c = spec["cpds"][node]
        d = int(c["output_dim"])
        if kind in ("eps", "u"):
            t = torch.cat([x.reshape(-1, d) for x in lst]).reshape(-1, n_samples, d).float()
        else:
            width = d if c["kind"] == "softmax_nn" else 1
            t = torch.cat([x.reshape(-1) for x in lst]).to(torch.int32)
            t = t.reshape(-1, n_samples, width) if width > 1 or c["kind"] == "softmax_nn" else t.reshape(-1, n_samples)
        out.setdefault(scope, {}).setdefault(node, {})[kind] = t.contiguous()
    return out


def log_to_strkeys(log):
    return {"|".join(k): [t.clone() for t in v] for k, v in log.items()}


def log_from_strkeys(d):
    return {tuple(k.split("|")): v for k, v in d.items()}


def gibbs_injection(log, spec):
    """RecordingNoise log of oracle.gibbs_sample -> {"init": {node: {...}}, "cand": {node: {...}}, "choice": {node: {"idx"}}}.
    "cand" arrays keep the reference's call shape with a leading sweep axis: [T, B, 8, D] (eps / u / softmax_nn idx) or
    [T, B, 8] (idx); "choice" is [T, B]."""
    out = {"init": {}, "cand": {}, "choice": {}}
    for key, lst in log.items():
        scope, node = key[0], key[1]
        kind = key[-1]
        dt = torch.float32 if kind in ("eps", "u") else torch.int32
        if scope == "gibbs_init":
            d = int(spec["cpds"][node]["output_dim"])
            t = lst[0].to(dt)
            t = t.reshape(-1, 1, d) if kind in ("eps", "u") or spec["cpds"][node]["kind"] == "softmax_nn" else t.reshape(-1, 1)
            out["init"].setdefault(node, {})[kind] = t.contiguous()
        elif len(key) == 4:  # ("gibbs", node, "choice", "idx")
            out["choice"].setdefault(node, {})["idx"] = torch.stack([x.reshape(-1) for x in lst]).to(torch.int32).contiguous()
        else:
            out["cand"].setdefault(node, {})[kind] = torch.stack([x.reshape(-1) for x in lst]).to(dt).contiguous()
    return out
