"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample sharding + the per-query
(m, l, q) all-gather merge reproduces the single-process softmax weights and ESS; query sharding
reproduces the batch-global fallback flag."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vectorizedbayesiannetwork_b200.dist import Shard, gather_stats, shared_seed


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stats(logw):
    m = logw.max(dim=1).values
    e = torch.exp(logw - m[:, None])
    return torch.stack([m, e.sum(1), (e * e).sum(1)], dim=1)


def _merge(g):  # [B, world, 3] -> [B, 3]  (same algebra as vbn_lse_merge)
    m = g[..., 0].max(dim=1).values
    sc = torch.exp(g[..., 0] - m[:, None])
    return torch.stack([m, (g[..., 1] * sc).sum(1), (g[..., 2] * sc * sc).sum(1)], dim=1)


def _worker(rank, world, port, logw, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, S = logw.shape
        sh = Shard("samples", rank, world)
        s_loc, s_off = sh.local_samples(S)
        local = logw[:, s_off: s_off + s_loc]
        merged = _merge(gather_stats(_stats(local), sh))
        w_local = torch.exp(local - merged[:, :1]) / merged[:, 1:2]
        ess = merged[:, 1] ** 2 / merged[:, 2]
        shq = Shard("queries", rank, world)
        b_loc, b_off = shq.local_queries(B)
        flag = torch.tensor([int(bool((ess[b_off: b_off + b_loc] < 0.1 * S).any()))], dtype=torch.int32)
        flag = shq.any_flag(flag)
        # calls without seed=: every rank must key Philox identically although each seeds torch on its own; the key
        # is agreed once per group, later calls derive theirs without a collective
        torch.manual_seed(1000 + rank)
        seeds = [shared_seed(int(torch.randint(0, 2**31 - 1, (1,)).item()), sh, "cpu") for _ in range(3)]
        out[rank] = (s_off, w_local, ess, int(flag.item()), shq.slice_queries(logw).shape[0], seeds)
    finally:
        dist.destroy_process_group()


def test_sample_sharded_merge_matches_single_process():
    torch.manual_seed(0)
    B, S, world = 5, 37, 2
    logw = torch.randn(B, S) * 3
    logw[3] = logw[3] * 10  # one degenerate query -> low ESS on rank 1's block only
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), logw, out), nprocs=world, join=True)
    w = torch.softmax(logw, dim=1)
    ess = 1.0 / (w**2).sum(1)
    want_flag = int(bool((ess < 0.1 * S).any()))
    got = torch.zeros_like(w)
    for r in range(world):
        s_off, w_local, ess_r, flag, b_loc, seeds = out[r]
        assert seeds == out[0][5] and len(set(seeds)) == 3
        got[:, s_off: s_off + w_local.shape[1]] = w_local
        torch.testing.assert_close(ess_r, ess, rtol=1e-5, atol=1e-6)
        assert flag == want_flag  # identical on every rank
    assert [out[r][4] for r in range(world)] == [3, 2]
    torch.testing.assert_close(got, w, rtol=1e-5, atol=1e-8)
