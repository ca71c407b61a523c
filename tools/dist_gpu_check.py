"""Multi-GPU check, run under torchrun on N >= 2 GPUs of one box:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_gpu_check.py
Sample-sharded and query-sharded importance sampling / likelihood weighting over NCCL must
reproduce the single-GPU result for the same seed (Philox counters carry GLOBAL (query, sample)
indices, so the union of the ranks' draws is the single-GPU draw set)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectorizedbayesiannetwork_b200 as V  # noqa: E402
from vectorizedbayesiannetwork_b200 import synthetic as S  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for name, spec, target, ev_nodes in (("lg_chain", S.lg_chain(12), "x5", ["x11"]),
                                         ("lg_mdn", S.random_dag_lg_mdn(40, seed=2), "n20", ["n38", "n39"])):
        g = torch.Generator().manual_seed(3)
        B, Sn = 2 * world + 2, 4096
        q = {"target": target, "evidence": {n: 0.5 * torch.randn(B, 1, generator=g) for n in ev_nodes}}
        model = V.VBN.from_spec(spec, device=dev)
        for method in ("likelihood_weighting", "importance_sampling"):
            model.set_inference_method(method, n_samples=Sn)
            w1, s1 = model.infer_posterior(q, seed=77)  # single GPU (every rank computes it)
            for kind in ("samples", "queries"):
                sh = V.Shard(kind, rank, world)
                w, s = model.infer_posterior(q, seed=77, shard=sh)
                if kind == "samples":
                    cnt, off = sh.local_samples(Sn)
                    rw, rs = w1[:, off:off + cnt], s1[:, off:off + cnt]
                else:
                    cnt, off = sh.local_queries(B)
                    rw, rs = w1[off:off + cnt], s1[off:off + cnt]
                fb = bool(getattr(model._inference, "_last_fallback", False))
                good = (torch.allclose(s, rs, rtol=1e-6, atol=1e-7) and torch.allclose(w, rw, rtol=2e-5, atol=1e-9))
                tot = w.sum(1) if kind == "queries" else None
                if kind == "samples":  # weights of one query sum to 1 across ranks
                    part = w.sum(1)
                    dist.all_reduce(part)
                    good = good and torch.allclose(part, torch.ones_like(part), rtol=1e-4, atol=1e-4)
                elif tot is not None:
                    good = good and torch.allclose(tot, torch.ones_like(tot), rtol=1e-4, atol=1e-4)
                flag = torch.tensor([int(good)], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print(f"{name:9s} {method:22s} shard={kind:8s} world={world} fallback={fb} "
                          f"{'OK' if flag.item() else 'MISMATCH'}", flush=True)
                ok = ok and bool(flag.item())
            # the fused summary (no [B,S] tensor): sample-sharded ranks merge their per-query records over NCCL
            st1 = model.infer_posterior(q, seed=77, summary=True)
            sh = V.Shard("samples", rank, world)
            st = model.infer_posterior(q, seed=77, shard=sh, summary=True)
            good = all(torch.allclose(st[k], st1[k], rtol=1e-4, atol=1e-5) for k in ("mean", "std", "ess"))
            flag = torch.tensor([int(good)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"{name:9s} {method:22s} summary shard=samples world={world} {'OK' if flag.item() else 'MISMATCH'}",
                      flush=True)
            ok = ok and bool(flag.item())
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if rank == 0:
        print("dist_gpu_check: all sharded runs match the single-GPU result")


if __name__ == "__main__":
    main()
