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
