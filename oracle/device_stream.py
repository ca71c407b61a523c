"""TEST INFRASTRUCTURE ONLY -- noise source that replays, inside the CPU oracle, the random draws one run of the
CUDA schedule kernel made.

The kernels' fast paths (plain linear-Gaussian / MDN / table ops, op-embedded root mixtures, the tcgen05 MLP tail)
only exist on the Philox route, so they cannot be driven with reference-recorded noise.  Instead the test asks the
library for the draws of the run it just made (``vbn_stream_draws``: same generator code, same counters), hands them
to this class, and runs the oracle's restatement of the reference algorithm (``oracle.vbn_oracle``) on them: both
sides then see identical normals and uniforms, and every sample / weight must agree to fp32 parity tolerance.

Stream layout (csrc/vbn_schedule.cuh:104-117, plan.py): every drawn node owns ``n_normals`` consecutive values of
the row's normal stream starting at ``n_off`` and ``n_uniforms`` of the uniform stream starting at ``u_off``;
value ``i`` is element ``i & 3`` of Philox block ``i >> 2``.  Roots of LW / MCM / ancestral passes use the
*shared* streams (keyed by the sample index only), everything else the per-row streams.

Categorical picks: the reference draws them with an exponential race (``torch.multinomial``); the device draws the
same law by inverse CDF on one uniform, ``k = #{j < K-1 : u >= cdf_j}``.  This class applies that rule to the
probabilities the oracle computed.  A pick whose uniform lies within ``tie_tol`` of a CDF edge can legitimately
differ from the device's (the two sides round the edge differently), which changes the row completely; such rows are
recorded in ``suspect[scope]`` so the caller can exclude them -- and bound how many there are.
"""
from __future__ import annotations

from typing import Dict

import torch


class DeviceStreamNoise:
    """``scopes[scope]`` = dict(streams={node: {"n_off", "u_off", "shared", "wb_u"}}, normals=[Nn, B, S],
    uniforms=[Nu, B, S], normals_shared=[Nn, 1, S] | None, uniforms_shared=[Nu, 1, S] | None, B=int, S=int).
    Implements the noise-source interface of oracle.vbn_oracle (TorchNoise)."""

    def __init__(self, scopes: Dict[str, dict], tie_tol: float = 2e-6) -> None:
        self.scopes = scopes
        self.tie_tol = float(tie_tol)
        self.calls: Dict[tuple, int] = {}
        self.suspect = {k: torch.zeros(v["B"], v["S"], dtype=torch.bool) for k, v in scopes.items()}

    # ---- bookkeeping -------------------------------------------------------------------------------------
    def _rows(self, key, kind: str, lead: int):
        """(scope dict, stream entry, slice of query rows this call covers)."""
        sc = self.scopes[key[0]]
        st = sc["streams"][key[1]]
        if st["shared"] or lead == sc["B"]:
            return sc, st, slice(0, None)
        # importance sampling draws row by row (importance_sampling.py:37-54): call number = query index
        ck = (key[0], key[1], kind)
        b = self.calls.get(ck, 0)
        self.calls[ck] = b + 1
        if lead != 1 or b >= sc["B"]:
            raise ValueError(f"unexpected draw shape for {key}: leading dim {lead}, call {b}")
        return sc, st, slice(b, b + 1)

    @staticmethod
    def _table(sc, st, which: str) -> torch.Tensor:
        t = sc[which + ("_shared" if st["shared"] else "")]
        if t is None:
            raise ValueError(f"no {which} table for a {'shared' if st['shared'] else 'per-row'} stream")
        return t

    # ---- continuous draws --------------------------------------------------------------------------------
    def _normal(self, key, shape):
        lead, s, d = int(shape[0]), int(shape[1]), int(shape[-1])
        sc, st, rows = self._rows(key, "eps", lead)
        tab = self._table(sc, st, "normals")
        out = torch.stack([tab[st["n_off"] + j][rows] for j in range(d)], dim=-1)  # [rows, S, D]
        return out.expand(lead, s, d).reshape(shape).clone()

    def normal(self, key, like):
        return self._normal(key, tuple(like.shape)).to(like.dtype)

    def normal_shape(self, key, shape, like):
        shape = tuple(shape)
        return self._normal(key, (shape[0], shape[1], int(torch.tensor(shape[2:]).prod()) if len(shape) > 2 else 1)
                            ).reshape(shape).to(like.dtype)

    def uniform(self, key, like):
        """softmax_nn within-bin variate (softmax_nn.py:666-669).  The reference draws it for every dim; a node whose
        dims are all discrete never uses it and the device does not spend stream values on it."""
        lead, s, d = int(like.shape[0]), int(like.shape[1]), int(like.shape[-1])
        sc, st, rows = self._rows(key, "u", lead)
        if st.get("wb_u") is None:
            return torch.zeros_like(like)
        tab = self._table(sc, st, "uniforms")
        out = torch.stack([tab[st["wb_u"] + j][rows] for j in range(d)], dim=-1)
        return out.expand(lead, s, d).reshape(like.shape).clone().to(like.dtype)

    # ---- categorical picks -------------------------------------------------------------------------------
    def _pick(self, key, probs: torch.Tensor, lead: int, n_dims: int) -> torch.Tensor:
        """probs [lead, S, n_dims, K] -> picks [lead, S, n_dims] by the device's inverse-CDF rule."""
        sc, st, rows = self._rows(key, "idx", lead)
        tab = self._table(sc, st, "uniforms")
        s = probs.shape[1]
        u = torch.stack([tab[st["u_off"] + j][rows] for j in range(n_dims)], dim=-1).expand(lead, s, n_dims)
        cdf = torch.cumsum(probs.float(), dim=-1)[..., :-1]
        cdf = cdf / probs.float().sum(dim=-1, keepdim=True)
        pick = (u.unsqueeze(-1) >= cdf).sum(dim=-1)
        near = ((u.unsqueeze(-1) - cdf).abs() < self.tie_tol).any(dim=-1).any(dim=-1)  # [lead, S]
        mask = self.suspect[key[0]]
        if st["shared"]:
            mask |= near.expand_as(mask) if near.shape[0] == 1 else near
        else:
            mask[rows] |= near
        return pick

    def categorical_probs(self, key, probs):
        # mdn.py:229 -- probs [b*S, K]
        sc = self.scopes[key[0]]
        s = sc["S"]
        lead = probs.shape[0] // s
        return self._pick(key, probs.reshape(lead, s, 1, -1), lead, 1).reshape(probs.shape[:-1])

    def categorical_logits(self, key, logits):
        # softmax_nn.py:651-652 / categorical_table.py -- logits [b, S, D, C]
        p = torch.softmax(logits.float(), dim=-1)
        if p.dim() == 3:
            p = p.unsqueeze(2)
        return self._pick(key, p, int(p.shape[0]), int(p.shape[2])).reshape(logits.shape[:-1])

    def multinomial(self, key, weights):
        # kde.py:178 -- weights [chunk, N] for chunk consecutive rows: not on any fast path
        raise NotImplementedError("KDE sampling is not replayed from the device stream")

    def randint(self, key, n, count, device):
        raise NotImplementedError("KDE sampling is not replayed from the device stream")

    def resample(self, key, weights, n):
        raise NotImplementedError("resampling is not replayed from the device stream")
