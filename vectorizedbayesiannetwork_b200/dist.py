"""Multi-GPU sharding of the posterior path: one process per GPU, torch.distributed (NCCL over
NVLink) only as plumbing.  Rows (query b, sample s) are independent through the whole schedule,
so the only cross-rank step is the per-query logsumexp merge when SAMPLES are sharded
(SURVEY.md section 8e); sharding QUERIES needs no data-path collective at all (only the
batch-global IS->LW fallback flag, 4 bytes).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch


def block_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Balanced contiguous partition of range(n): returns (count, offset) of ``rank``."""
    base, extra = divmod(int(n), int(world))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return count, offset


@dataclass
class Shard:
    kind: str            # "queries" | "samples"
    rank: int
    world: int
    group: Optional[object] = None

    def __post_init__(self):
        if self.kind not in ("queries", "samples"):
            raise ValueError("Shard.kind must be 'queries' or 'samples'")
        if not (0 <= self.rank < self.world):
            raise ValueError("bad rank/world")

    def local_queries(self, b: int) -> Tuple[int, int]:
        if self.kind != "queries":
            return int(b), 0
        if b < self.world:  # the same test on EVERY rank: nobody walks into a collective the others never reach
            raise ValueError(f"B={b} queries cannot be sharded over {self.world} ranks; shard samples instead")
        return block_bounds(b, self.rank, self.world)

    def local_samples(self, s: int) -> Tuple[int, int]:
        if self.kind != "samples":
            return int(s), 0
        if s < self.world:  # raised on every rank alike
            raise ValueError(f"S={s} samples cannot be sharded over {self.world} ranks")
        return block_bounds(s, self.rank, self.world)

    def slice_queries(self, v: torch.Tensor) -> torch.Tensor:
        if self.kind != "queries":
            return v
        cnt, off = block_bounds(v.shape[0], self.rank, self.world)
        return v[off : off + cnt]

    def any_flag(self, flag: torch.Tensor) -> torch.Tensor:
        """Logical OR of a per-rank int flag over all ranks."""
        if self.world == 1:
            return flag
        import torch.distributed as dist

        out = flag.clone()
        dist.all_reduce(out, op=dist.ReduceOp.MAX, group=self.group)
        return out


_SESSION_KEYS: dict = {}  # process group -> [rank 0's base key, calls so far]


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def reset_shared_seed(group=None) -> None:
    """Forget the session key of ``group`` (all groups when None): the next sharded call without ``seed=`` agrees on a
    fresh one -- call it on every rank after re-seeding torch if the sharded draws are to follow the new seed."""
    if group is None:
        _SESSION_KEYS.clear()
    else:
        _SESSION_KEYS.pop(id(group), None)


def shared_seed(seed: int, shard: Optional[Shard], device) -> int:
    """The Philox key of a sharded call whose caller did not pass ``seed=``.  The ranks' rows are pieces of ONE draw set
    (counters carry global (query, sample) indices; roots are shared across query shards), so every rank must use
    the same key.  Rank 0's freshly drawn key is broadcast ONCE per process group; call k then uses
    splitmix64(base + k) on every rank -- the ranks make the same sequence of sharded calls anyway (the passes hold
    collectives) --, so there is no collective and no device -> host read on the per-call path (a per-call broadcast
    cost two blocking round trips per importance-sampling step: 0.4 ms of a 2 ms step on 8 GPUs)."""
    if shard is None or shard.world == 1:
        return seed
    key = id(shard.group) if shard.group is not None else 0
    ent = _SESSION_KEYS.get(key)
    if ent is None:
        import torch.distributed as dist

        t = torch.tensor([seed], dtype=torch.int64, device=device)
        dist.broadcast(t, src=0, group=shard.group)
        ent = _SESSION_KEYS[key] = [int(t.item()), 0]
    ent[1] += 1
    return _splitmix64((ent[0] + ent[1]) & 0xFFFFFFFFFFFFFFFF) & 0x7FFFFFFFFFFFFFFF


def gather_stats(stats: torch.Tensor, shard: Shard) -> torch.Tensor:
    """All-gather per-rank per-query records ([B, k]: (m, l, q) triples or the 16-float merged records of
    vbn_segment_merge) -> [B, world, k].  Payload is 4 k B bytes per rank -- latency bound on NVLink: one collective
    into one preallocated tensor; the merge itself is a CUDA kernel (vbn_segment_merge / vbn_lse_merge)."""
    if shard.world == 1:
        return stats.unsqueeze(1)
    import torch.distributed as dist

    b = stats.shape[0]
    out = torch.empty((shard.world * b,) + tuple(stats.shape[1:]), device=stats.device, dtype=stats.dtype)
    dist.all_gather_into_tensor(out, stats.contiguous(), group=shard.group)  # rank-major concatenation along dim 0
    return out.view((shard.world, b) + tuple(stats.shape[1:])).transpose(0, 1).contiguous()


def auto_shard(n_queries: int, rank: int, world: int, group=None) -> Optional[Shard]:
    """Queries when there are enough of them (no collective), else samples."""
    if world <= 1:
        return None
    return Shard("queries" if n_queries >= world else "samples", rank, world, group)
