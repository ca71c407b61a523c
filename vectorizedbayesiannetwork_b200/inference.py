"""CUDA drop-ins for the reference's sampling-based inference methods, registered under the same
keys (vbn/inference/{likelihood_weighting,importance_sampling,monte_carlo_marginalization}.py)
and the ancestral sampler (vbn/sampling/ancestral.py).  Constructors, ``infer_posterior`` /
``sample`` signatures, return shapes and the attributes tests read (``n_samples``,
``ess_threshold``, ``_last_ess``, ``_last_fallback``, ``_cache``) follow the reference.

Extra keyword arguments understood by every method (all optional):
  noise : {node: {"eps"|"u"|"idx": tensor}} -- inject the draws instead of Philox (parity tests);
          importance_sampling takes {"is": {...}, "lw": {...}} for its two passes
  seed  : int Philox key (default: one draw from torch's global CPU generator per pass)
  shard : dist.Shard -- this rank's slice of the queries or samples (multi-GPU)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import engine as E
from .core import (Query, infer_batch_size, model_cpds, register_inference, register_sampling)
from .cpds import TABLE_KINDS
from .dist import Shard, gather_stats, shared_seed
from .plan import Role, compile_gibbs, compile_schedule


def _check_model(vbn) -> torch.device:
    dev = E.require_cuda(getattr(vbn, "device", None))
    return dev


def _topology(vbn):
    topo = list(vbn.dag.topological_order())
    parents = {n: list(vbn.dag.parents(n)) for n in topo}
    return topo, parents


def clamp_evidence(x: torch.Tensor) -> torch.Tensor:
    """vbn/inference/_core.py:112-114."""
    x = torch.nan_to_num(x, nan=0.0, posinf=1e6, neginf=-1e6)
    return x.clamp(min=-1e6, max=1e6)


class _ScheduleRunner:
    """Shared by all methods: builds (and caches) the schedule for a query signature and runs it."""

    def __init__(self) -> None:
        self._cache: Dict[tuple, E.DevicePlan] = {}

    @staticmethod
    def barren_nodes(topo, parents, query: Query, mode: str) -> set:
        """Nodes whose value cannot reach the result: not the target, not scored evidence, and no directed path to
        either (the classic barren-node rule).  ``do`` nodes -- and, in the unweighted modes, evidence nodes -- are
        clamped, so nothing is needed upstream of them."""
        scored = set(query.evidence) if mode in ("lw", "is") else set()
        need = {query.target} | scored
        cut = set(query.do) | (set(query.evidence) - scored)
        stack = [n for n in need if n not in cut or n in scored]
        while stack:
            n = stack.pop()
            if n in cut and n not in scored:
                continue
            for p in parents[n]:
                if p not in need:
                    need.add(p)
                    stack.append(p)
        return set(topo) - need

    def plan_for(self, vbn, query: Query, mode: str, *, inject=frozenset(), store_all=False,
                 only=None, summary: bool = False, prune: bool = False) -> E.DevicePlan:
        cpds = model_cpds(vbn)
        fp = tuple((c._uid, c._version) for c in cpds.values())
        key = (fp, mode, query.target, tuple(sorted(query.evidence)), tuple(sorted(query.do)),
               tuple(sorted(inject)), bool(store_all), None if only is None else tuple(only), str(vbn.device),
               bool(summary), bool(prune))
        plan = self._cache.get(key)
        if plan is not None:
            return plan
        topo, parents = _topology(vbn)
        roles: Dict[str, Role] = {}
        for n in topo:
            if only is not None and n not in only:
                continue
            if n in query.do:
                r = Role(src="fixed_q", density=False)
            elif n in query.evidence:
                if mode in ("lw", "is", "rb"):
                    r = Role(src="fixed_q", add_logw=True)
                else:
                    r = Role(src="fixed_q", density=False)
            else:
                shared = mode != "is" and len(parents[n]) == 0 and mode != "mcm_fast"
                r = Role(src="sample", shared=shared, inject=n in inject)
            if (n == query.target and not summary) or store_all:
                r.store = True
            if mode == "rb" and n == query.target:
                r = Role(src="sample", store=True, out_params=True)
            roles[n] = r
        if mode in ("mcm", "mcm_fast"):
            t = roles[query.target]
            t.out_logp = True
            t.density = True
        barren = (self.barren_nodes(topo, parents, query, mode)
                  if prune and mode in ("lw", "is", "mcm") and only is None and not store_all else ())
        prog = compile_schedule(topo, parents, cpds, roles, table_fn=E.discrete_table_fn,
                                keep_live=query.target if summary else None, barren=barren)
        plan = E.DevicePlan(prog, vbn.device)
        if len(self._cache) > 64:
            self._cache.clear()
        self._cache[key] = plan
        return plan

    def fixed_table(self, plan: E.DevicePlan, query: Query, b: int, *, clamp_obs: bool, shard: Optional[Shard]):
        """[n_fixed_cols][B] table of the per-query evidence / do values (prepare_fixed_values, _core.py:117-135).
        The last table is kept: a caller that repeats a query with the same (unmodified) tensors -- the steady state
        of a serving loop -- pays no clamp / concatenate launches again."""
        prog = plan.program
        if not prog.n_fixed_cols:
            return None
        src = [query.do[n] if n in query.do else query.evidence[n] for n in prog.fixed_cols]
        key = (id(plan), b, bool(clamp_obs), None if shard is None else (shard.kind, shard.rank, shard.world),
               tuple((id(v), v._version, v.data_ptr()) for v in src))
        hit = getattr(self, "_fixed_cache", None)
        if hit is not None and hit[0] == key:
            return hit[2]
        table = self._build_fixed_table(plan, query, b, clamp_obs=clamp_obs, shard=shard)
        self._fixed_cache = (key, src, table)  # `src` keeps the ids alive
        return table

    @staticmethod
    def _build_fixed_table(plan: E.DevicePlan, query: Query, b: int, *, clamp_obs: bool, shard: Optional[Shard]):
        prog = plan.program
        cols = []
        for n in prog.fixed_cols:  # insertion order == op order
            v = query.do[n] if n in query.do else query.evidence[n]
            if v.dtype != torch.float32 or v.device != plan.device:
                v = v.to(device=plan.device, dtype=torch.float32)
            if clamp_obs and n in query.evidence:
                v = clamp_evidence(v)
            if shard is not None:
                v = shard.slice_queries(v)
            if v.shape[0] != b:
                raise ValueError("Evidence and do batch sizes must match.")
            cols.append(v.t())
        return torch.cat(cols, dim=0).contiguous()

    def forward(self, vbn, query: Query, n_samples: int, mode: str, *, noise=None, seed=None,
                shard: Optional[Shard] = None, clamp_obs: bool = False, store_all: bool = False,
                only=None, n_queries: Optional[int] = None, summary: bool = False, classes: int = 0,
                prune: bool = False):
        """Runs one pass.  Returns dict(logw, logp, stores={node: [B,S,D]}, seg, plan, b, s).
        Weighted passes (``lw`` / ``is``) also emit the per-warp records of the fused weight reduction (``seg``);
        ``summary``: the target is not stored and no log-weight buffer exists -- its weighted moments (and, with
        ``classes`` in 1..8, its class histogram) are accumulated into the records instead."""
        dev = _check_model(vbn)
        noise = noise or {}
        plan = self.plan_for(vbn, query, mode, inject=frozenset(noise), store_all=store_all, only=only,
                             summary=summary, prune=prune)
        prog = plan.program
        b_full = infer_batch_size(query.evidence, query.do) if n_queries is None else n_queries
        b, s, q_off, s_off = b_full, int(n_samples), 0, 0
        if shard is not None:
            b, q_off = shard.local_queries(b_full)
            s, s_off = shard.local_samples(s)
        with torch.cuda.device(dev):
            fixed = self.fixed_table(plan, query, b, clamp_obs=clamp_obs, shard=shard)
            stores = {n: torch.empty(b, s, prog.store_widths[n], device=dev, dtype=torch.float32) for n in prog.stores}
            logw = torch.empty(b, s, device=dev, dtype=torch.float32) if prog.needs_logw and not summary else None
            logp = torch.empty(b, s, device=dev, dtype=torch.float32) if prog.needs_logp else None
            seg = E.segment_records(b, s, dev) if mode in ("lw", "is", "rb") else None
            # the device error flag is only ever raised by discrete ops (softmax_nn classes, table supports)
            needs_flag = getattr(prog, "_needs_flag", None)
            if needs_flag is None:
                needs_flag = prog._needs_flag = bool(np.isin(prog.ops["kind"], (L.OP_SNN, L.OP_TAB)).any())
            flag = torch.zeros(1, device=dev, dtype=torch.int32) if needs_flag else None
            plan.run(b, s, fixed=fixed, stores=[stores[n] for n in prog.stores],
                     noise=[noise[n] for n in prog.noise], logw=logw, logp=logp,
                     logp_as_pdf=mode in ("mcm", "mcm_fast"),
                     seed=shared_seed(E.draw_seed(), shard, dev) if seed is None else seed,
                     query_offset=q_off, sample_offset=s_off, error_flag=flag,
                     seg=seg, seg_slot=prog.keep_slot if summary else -1, seg_classes=classes if summary else 0)
        return {"logw": logw, "logp": logp, "stores": stores, "plan": plan, "b": b, "s": s, "flag": flag, "seg": seg,
                "dev": dev}


def _prune(kwargs) -> bool:
    """``prune=True`` (or VBN_PRUNE=1): barren nodes are left out of the schedule.  Off by default: the schedule then
    walks every node like the reference's loops do.  Either way the result is the same bit for bit."""
    import os

    v = kwargs.get("prune")
    return bool(v) if v is not None else os.environ.get("VBN_PRUNE", "0") == "1"


def _flag_message(cpds) -> str:
    """The reference's wording for an off-support value: categorical_table / categorical_embedded_softmax raise
    'Found values outside support.' (categorical_table.py:12-21), softmax_nn 'Found values outside discrete class
    set.' (softmax_nn.py:623-625)."""
    kinds = {c.kind for c in cpds}
    if kinds & set(TABLE_KINDS) and "softmax_nn" not in kinds:
        return "Found values outside support."
    return "Found values outside discrete class set."


def _raise_if_flagged(vbn, out) -> None:
    """The device error flag exists exactly when the schedule holds an op that can raise it (VBN_OP_SNN / VBN_OP_TAB:
    softmax_nn class sets, strict supports of the table kinds); a non-zero flag is the reference's ValueError."""
    if out["flag"] is not None and int(out["flag"].item()) != 0:
        raise ValueError(_flag_message(model_cpds(vbn).values()))


def _reduce(out, *, shard: Optional[Shard] = None, ess_threshold: float = 0.0):
    """Per-query reduction of a weighted pass from the records the kernel emitted: returns (merged [B,16], stats
    [B,3], flag or None) -- see engine.segment_merge.  With the samples sharded over ranks the per-rank merged records
    are all-gathered (64 B per query) and folded once more; the fallback flag is raised from the GLOBAL ESS."""
    sharded = shard is not None and shard.kind == "samples" and shard.world > 1
    want_flag = ess_threshold > 0.0
    merged, stats, flag = E.segment_merge(out["seg"], out["s"], ess_threshold=0.0 if sharded else ess_threshold,
                                          want_flag=want_flag and not sharded)
    if sharded:
        merged, stats, flag = E.segment_merge(gather_stats(merged, shard), 0, ess_threshold=ess_threshold,
                                              want_flag=want_flag)
    return merged, stats, flag


def _weights(out, *, normalize=True, eps=1e-12, shard: Optional[Shard] = None, ess_threshold: float = 0.0):
    """softmax over samples (+ESS) from a pass's log-weights; merges across ranks when the
    samples are sharded.  Returns (w, ess, stats, flag)."""
    b, s = out["b"], out["s"]
    logw = out["logw"]
    if logw is None:  # no evidence with a density: uniform weights, like softmax of zeros
        logw = torch.zeros(b, s, device=out["dev"], dtype=torch.float32)
    merged, stats, flag = _reduce(out, shard=shard, ess_threshold=ess_threshold)
    w, _ = E.normalize_weights(logw, stats, normalize=normalize, eps=eps, want_ess=False)
    return w, merged[:, 7], stats, flag


def _summary(merged: torch.Tensor, classes: int = 0) -> Dict[str, torch.Tensor]:
    """VBN._posterior_stats (vbn/vbn.py:495-504) of a one-dimensional target from a merged record: weighted mean,
    std = sqrt(sum w (x - mean)^2), ESS = 1 / sum w^2; ``probs`` [B, classes] when a class histogram was asked for
    (benchmarking/models/vbn.py:202-242)."""
    l = merged[:, 1].clamp_min(1e-30)
    out = {"mean": merged[:, 3:4], "std": (merged[:, 5:6] / l.unsqueeze(1)).clamp_min(0.0).sqrt(), "ess": merged[:, 7]}
    if classes:
        out["probs"] = merged[:, 8:8 + classes] / l.unsqueeze(1)
    return out


def _summary_args(vbn, query: Query, kwargs):
    """(fused?, classes) for ``summary=True``: the kernel-side accumulation covers one-dimensional targets and up to 8
    classes; anything else is summarised from the materialised tensors (still on the device)."""
    want = kwargs.get("summary")
    if not want:
        return False, 0
    classes = int(want.get("classes", 0)) if isinstance(want, dict) else 0
    cpd = model_cpds(vbn)[query.target]
    return int(cpd.output_dim) == 1 and 0 <= classes <= 8 and not kwargs.get("noise"), classes


def _summary_from_tensors(w: torch.Tensor, samples: torch.Tensor, classes: int) -> Dict[str, torch.Tensor]:
    st = E.posterior_stats(w, samples)
    if classes:
        st["probs"] = E.weighted_histogram(samples, w, classes)
    return st


@register_inference("likelihood_weighting")
class LikelihoodWeighting:
    """vbn/inference/likelihood_weighting.py:11-82."""

    def __init__(self, n_samples: int = 512, eps: float = 1e-12, normalize: bool = True, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.eps = float(eps)
        self.normalize = bool(normalize)
        self._runner = _ScheduleRunner()
        self._cache = self._runner._cache

    def infer_posterior(self, vbn, query: Query, **kwargs):
        n_samples = int(kwargs.get("n_samples", self.n_samples))
        normalize = bool(kwargs.get("normalize", self.normalize))
        eps = float(kwargs.get("eps", self.eps))
        shard = kwargs.get("shard")
        fused, classes = _summary_args(vbn, query, kwargs)
        out = self._runner.forward(vbn, query, n_samples, "lw", noise=kwargs.get("noise"),
                                   seed=kwargs.get("seed"), shard=shard, clamp_obs=True, summary=fused,
                                   classes=classes, prune=_prune(kwargs))
        _raise_if_flagged(vbn, out)
        if fused:  # summary=True: per-query posterior summary, nothing of size [B, S] is materialised
            return _summary(_reduce(out, shard=shard)[0], classes)
        w, _, _, _ = _weights(out, normalize=normalize, eps=eps, shard=shard)
        if kwargs.get("summary"):
            return _summary_from_tensors(w, out["stores"][query.target], classes)
        return w, out["stores"][query.target]


@register_inference("importance_sampling")
class ImportanceSampling:
    """vbn/inference/importance_sampling.py:14-93: per-query independent root draws, ESS test,
    whole-batch fallback to likelihood weighting."""

    def __init__(self, n_samples: int = 200, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.ess_threshold = 0.1
        self._runner = _ScheduleRunner()
        self._cache = self._runner._cache
        self._lw = LikelihoodWeighting(n_samples=self.n_samples)
        self._last_fallback = False
        self._last_ess: Optional[torch.Tensor] = None

    def infer_posterior(self, vbn, query: Query, **kwargs):
        n_samples = int(kwargs.get("n_samples", self.n_samples))
        shard = kwargs.get("shard")
        noise = kwargs.get("noise") or {}
        seed = kwargs.get("seed")
        fused, classes = _summary_args(vbn, query, kwargs)
        out = self._runner.forward(vbn, query, n_samples, "is", noise=noise.get("is"), seed=seed, shard=shard,
                                   summary=fused, classes=classes, prune=_prune(kwargs))
        _raise_if_flagged(vbn, out)
        threshold = max(1.0, self.ess_threshold * float(n_samples))
        # the softmax statistics, the ESS and the fallback test any(ESS < threshold) come out of ONE merge launch
        merged, stats, flag = _reduce(out, shard=shard, ess_threshold=threshold)
        self._last_ess = merged[:, 7]
        if shard is not None and shard.kind == "queries" and shard.world > 1:
            flag = shard.any_flag(flag)  # the fallback is batch-global (importance_sampling.py:85-88)
        if int(flag.item()) != 0:  # the one device -> host read of the pass
            self._last_fallback = True
            lw_kwargs = {"n_samples": n_samples, "shard": shard, "noise": noise.get("lw"),
                         "summary": kwargs.get("summary"), "prune": kwargs.get("prune")}
            if seed is not None:
                lw_kwargs["seed"] = seed + 1
            return self._lw.infer_posterior(vbn, query, **lw_kwargs)
        self._last_fallback = False
        if fused:
            return _summary(merged, classes)
        logw = out["logw"] if out["logw"] is not None else torch.zeros(out["b"], out["s"], device=out["dev"])
        w, _ = E.normalize_weights(logw, stats, want_ess=False)
        if kwargs.get("summary"):
            return _summary_from_tensors(w, out["stores"][query.target], classes)
        return w, out["stores"][query.target]


@register_inference("monte_carlo_marginalization")
class MonteCarloMarginalization:
    """vbn/inference/monte_carlo_marginalization.py:12-92 (three paths)."""

    def __init__(self, n_samples: int = 200, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self._runner = _ScheduleRunner()
        self._cache = self._runner._cache

    def infer_posterior(self, vbn, query: Query, **kwargs):
        n_samples = int(kwargs.get("n_samples", self.n_samples))
        shard = kwargs.get("shard")
        dev = _check_model(vbn)
        b = infer_batch_size(query.evidence, query.do)
        target = query.target
        fixed = set(query.evidence) | set(query.do)
        tparents = list(vbn.dag.parents(target))

        if target in query.do:  # :33-37
            v = query.do[target].to(device=dev, dtype=torch.float32)
            if shard is not None:
                v = shard.slice_queries(v)
            return (torch.ones(v.shape[0], n_samples, device=dev, dtype=torch.float32),
                    v.unsqueeze(1).expand(v.shape[0], n_samples, -1))

        if all(p in fixed for p in tparents):  # :39-58 -- only the target CPD is evaluated
            only = tparents + [target]
            # a parentless, unobserved target is drawn with parents=None -> batch of 1 (:49-56)
            nq = 1 if (not tparents and target not in fixed) else b
            sub = Query(target=target,
                        evidence={k: v for k, v in query.evidence.items() if k in only},
                        do={k: v for k, v in query.do.items() if k in only})
            run_shard = shard if nq == b else None
            out = self._runner.forward(vbn, sub, n_samples, "mcm_fast", noise=kwargs.get("noise"),
                                       seed=kwargs.get("seed"), shard=run_shard, only=only, n_queries=nq)
            _raise_if_flagged(vbn, out)
            return out["logp"], out["stores"][target]

        out = self._runner.forward(vbn, query, n_samples, "mcm", noise=kwargs.get("noise"),
                                   seed=kwargs.get("seed"), shard=shard, prune=_prune(kwargs))  # :60-92
        _raise_if_flagged(vbn, out)
        return out["logp"], out["stores"][target]


@register_inference("resampled_importance_sampling")
class ResampledImportanceSampling:
    """vbn/inference/resampled_importance_sampling.py:13-105.

    Resampling permutes rows inside a query, which the fused row-local schedule cannot do on chip, so
    the schedule is cut after every evidence node: each segment is one fused launch that reads the live
    node columns of the previous segments from HBM (SoA, per-row inputs) and writes the columns later
    segments still need.  Between segments: ESS test (4-byte flag read, like IS), and when any query is
    below the threshold, multinomial resampling of every live column (vbn_row_cdf /
    vbn_resample_indices / vbn_gather_rows) and a reset of the log-weights."""

    def __init__(self, n_samples: int = 512, ess_threshold: float = 0.5, resample: bool = True,
                 clamp_obs: bool = True, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.ess_threshold = float(ess_threshold)
        self.resample = bool(resample)
        self.clamp_obs = bool(clamp_obs)
        self._cache: Dict[tuple, list] = {}
        self._last_ess: Optional[torch.Tensor] = None
        self._last_resampled = False

    # ---- segment compilation ------------------------------------------------------------------
    def _segments(self, vbn, query: Query, inject: frozenset):
        cpds = model_cpds(vbn)
        fp = tuple((c._uid, c._version) for c in cpds.values())
        key = (fp, query.target, tuple(sorted(query.evidence)), tuple(sorted(query.do)), tuple(sorted(inject)),
               str(vbn.device))
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        topo, parents = _topology(vbn)
        pos = {n: i for i, n in enumerate(topo)}
        fixed = set(query.evidence) | set(query.do)
        evaluated = lambda n: n not in query.do  # drawn or scored: reads its parents
        bounds, start = [], 0
        for i, n in enumerate(topo):
            if n in query.evidence:
                bounds.append((start, i + 1, True))
                start = i + 1
        if start < len(topo):
            bounds.append((start, len(topo), False))
        segs = []
        for lo, hi, ends_ev in bounds:
            seg_nodes = topo[lo:hi]
            needs = [p for n in seg_nodes if evaluated(n) for p in parents[n] if pos[p] < lo]
            carried_in = sorted({p for p in needs if p not in fixed}, key=pos.get)
            fixed_in = sorted({p for p in needs if p in fixed}, key=pos.get)
            later = {p for n in topo[hi:] if evaluated(n) for p in parents[n]}
            keep = lambda n: n not in fixed and (n in later or n == query.target)
            roles: Dict[str, Role] = {}
            for n in carried_in:
                roles[n] = Role(src="fixed_row", density=False)
            for n in fixed_in:
                roles[n] = Role(src="fixed_q", density=False)
            for n in seg_nodes:
                if n in query.do:
                    roles[n] = Role(src="fixed_q", density=False)
                elif n in query.evidence:
                    roles[n] = Role(src="fixed_q", add_logw=True)
                else:
                    roles[n] = Role(src="sample", shared=len(parents[n]) == 0, inject=n in inject, store=keep(n))
            order = carried_in + fixed_in + seg_nodes
            prog = compile_schedule(order, parents, cpds, roles, table_fn=E.discrete_table_fn)
            segs.append({"plan": E.DevicePlan(prog, vbn.device), "ends_ev": ends_ev,
                         "live_after": [n for n in topo[:hi] if keep(n)]})
        if len(self._cache) > 64:
            self._cache.clear()
        self._cache[key] = segs
        return segs

    def infer_posterior(self, vbn, query: Query, **kwargs):
        n_samples = int(kwargs.get("n_samples", self.n_samples))
        ess_threshold = float(kwargs.get("ess_threshold", self.ess_threshold))
        resample = bool(kwargs.get("resample", self.resample))
        clamp_obs = bool(kwargs.get("clamp_obs", self.clamp_obs))
        shard = kwargs.get("shard")
        if shard is not None and shard.kind != "queries":
            raise ValueError("resampled_importance_sampling shards queries only (resampling is per query)")
        noise = kwargs.get("noise") or {}
        picks = list(noise.get("__resample__", []))
        seed = E.draw_seed() if kwargs.get("seed") is None else int(kwargs["seed"])
        dev = _check_model(vbn)
        b_full = infer_batch_size(query.evidence, query.do)
        b, q_off = (b_full, 0) if shard is None else shard.local_queries(b_full)
        s = n_samples
        rows = b * s
        threshold = max(1.0, ess_threshold * float(s)) if ess_threshold <= 1.0 else float(ess_threshold)
        segs = self._segments(vbn, query, frozenset(k for k in noise if k != "__resample__"))
        live: Dict[str, torch.Tensor] = {}  # node -> [D, B*S] column block
        self._last_resampled = False
        with torch.cuda.device(dev):
            logw = torch.zeros(b, s, device=dev, dtype=torch.float32)
            flag_dev = torch.zeros(1, device=dev, dtype=torch.int32)
            for k, seg in enumerate(segs):
                plan = seg["plan"]
                prog = plan.program
                fixed = _ScheduleRunner._build_fixed_table(plan, query, b, clamp_obs=clamp_obs, shard=shard) if plan.program.n_fixed_cols else None
                new = {n: torch.empty(prog.dims[n], rows, device=dev, dtype=torch.float32) for n in prog.stores}
                plan.run(b, s, fixed=fixed, inputs=[live[n].t() for n in prog.inputs],
                         stores=[new[n].t() for n in prog.stores], noise=[noise[n] for n in prog.noise],
                         logw=logw if prog.needs_logw else None, logw_accumulate=True, seed=seed, call_offset=k,
                         query_offset=q_off, error_flag=flag_dev)
                live.update(new)
                live = {n: live[n] for n in seg["live_after"]}
                if seg["ends_ev"] and resample:
                    stats = E.lse_stats(logw)
                    w, ess = E.normalize_weights(logw, stats)
                    self._last_ess = ess
                    flag = E.ess_below(stats, threshold)
                    if shard is not None and shard.world > 1:
                        flag = shard.any_flag(flag)
                    if int(flag.item()) != 0:
                        if picks:
                            idx = picks.pop(0).to(device=dev, dtype=torch.int32).reshape(b, s).contiguous()
                        else:
                            idx = E.resample_indices(w, seed=seed, call_offset=0x40000000 + k, query_offset=q_off)
                        live = {n: E.gather_rows(t, idx) for n, t in live.items()}
                        logw.zero_()
                        self._last_resampled = True
            cpds = model_cpds(vbn)
            if any(c.kind in ("softmax_nn",) + TABLE_KINDS for c in cpds.values()) and int(flag_dev.item()) != 0:
                raise ValueError(_flag_message(cpds.values()))
            stats = E.lse_stats(logw)
            w, _ = E.normalize_weights(logw, stats)
            t = query.target
            if t in live:
                d = live[t].shape[0]
                samples = live[t].view(d, b, s).permute(1, 2, 0).contiguous()
            else:  # the target is itself fixed
                v = query.do[t] if t in query.do else query.evidence[t]
                v = v.to(device=dev, dtype=torch.float32)
                if clamp_obs and t in query.evidence:
                    v = clamp_evidence(v)
                if shard is not None:
                    v = shard.slice_queries(v)
                samples = v.unsqueeze(1).expand(b, s, -1)
        return w, samples


@register_inference("rao_blackwellized_marginalization")
class RaoBlackwellizedMarginalization(object):
    """vbn/inference/rao_blackwellized_marginalization.py:15-324: draw the non-descendants
    of the target with likelihood weighting, then mix the target's conditional analytically over the
    particles.  One fused launch draws the particles AND reads out the target's parameters per particle
    (VBN_F_OUT_PARAMS); the mixture (categorical marginal / Gaussian mixture on a grid) is one more kernel."""

    def __init__(self, n_samples: int = 200, n_particles: Optional[int] = None, stddevs: float = 4.0,
                 min_scale: float = 1e-6, fallback: str = "likelihood_weighting", **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.n_particles = int(n_particles) if n_particles is not None else self.n_samples
        self.stddevs = float(stddevs)
        self.min_scale = float(min_scale)
        self._runner = _ScheduleRunner()
        self._cache = self._runner._cache
        self._last_fallback = False
        self._last_reason = None
        self.fallback = str(fallback).strip().lower() if fallback is not None else "none"
        self._fallback = None
        if self.fallback != "none":
            from .core import INFERENCE_REGISTRY

            if self.fallback not in INFERENCE_REGISTRY:
                raise ValueError(f"Unknown fallback inference '{fallback}'. Available: {list(INFERENCE_REGISTRY.keys())}")
            if self.fallback == "rao_blackwellized_marginalization":
                raise ValueError("fallback cannot be 'rao_blackwellized_marginalization'")
            fk = dict(kwargs)
            fk.setdefault("n_samples", self.n_samples)
            self._fallback = INFERENCE_REGISTRY[self.fallback](**fk)

    def _fallback_infer(self, vbn, query, *, reason: str, **kwargs):
        self._last_fallback = True
        self._last_reason = reason
        if self._fallback is None:
            raise RuntimeError("rao_blackwellized_marginalization cannot handle this query and has no fallback")
        if kwargs.get("noise") is not None:  # injected draws: {"rb": particle pass, "lw": fallback pass}
            kwargs = dict(kwargs, noise=kwargs["noise"].get("lw"))
        return self._fallback.infer_posterior(vbn, query, **kwargs)

    def infer_posterior(self, vbn, query: Query, **kwargs):
        from . import _lib as L

        self._last_fallback, self._last_reason = False, None
        n_samples = max(1, int(kwargs.get("n_samples", self.n_samples)))
        n_particles = max(1, int(kwargs.get("n_particles", self.n_particles)))
        dev = _check_model(vbn)
        b = infer_batch_size(query.evidence, query.do)
        topo, parents = _topology(vbn)
        target = query.target
        children: Dict[str, list] = {n: [] for n in topo}
        for n in topo:
            for p in parents[n]:
                children[p].append(n)
        desc, stack = set(), [target]
        while stack:
            for ch in children[stack.pop()]:
                if ch not in desc:
                    desc.add(ch)
                    stack.append(ch)
        fixed = set(query.evidence) | set(query.do)
        if desc & fixed:
            return self._fallback_infer(vbn, query, reason="target has observed/intervened descendants", **kwargs)
        if target in fixed:
            v = query.do[target] if target in query.do else clamp_evidence(query.evidence[target])
            v = v.to(device=dev, dtype=torch.float32)
            return torch.ones(b, 1, device=dev), v.unsqueeze(1).expand(b, 1, -1)
        cpds = model_cpds(vbn)
        tc = cpds[target]
        categorical = tc.kind in ("softmax_nn",) + TABLE_KINDS and tc.output_dim == 1
        gaussian = tc.kind in ("linear_gaussian", "gaussian_nn", "rff_gaussian") and tc.output_dim == 1
        if not (categorical or gaussian):
            return self._fallback_infer(vbn, query, reason="unsupported target CPD for RB marginalization", **kwargs)
        only = [n for n in topo if n not in desc]
        out = self._runner.forward(vbn, query, n_particles, "rb", noise=(kwargs.get("noise") or {}).get("rb"),
                                   seed=kwargs.get("seed"),
                                   clamp_obs=True, only=only)
        _raise_if_flagged(vbn, out)
        w, _, _, _ = _weights(out)
        params = out["stores"][target]  # [B, S, width] parameter read-out of the target per particle
        lib = L.load()
        with torch.cuda.device(dev):
            sp = E._stream_ptr(dev)
            if categorical:
                k = int(tc.n_classes)
                marginal = torch.empty(b, k, device=dev, dtype=torch.float32)
                L.check(lib.vbn_weighted_sum(w.data_ptr(), params.data_ptr(), b, n_particles, k, marginal.data_ptr(), sp))
                L.count_launch(1)
                support = tc._sample_values[0].to(device=dev, dtype=torch.float32)
                return marginal, support.view(1, -1, 1).expand(b, -1, 1)
            pdf = torch.empty(b, n_samples, device=dev, dtype=torch.float32)
            grid = torch.empty(b, n_samples, 1, device=dev, dtype=torch.float32)
            L.check(lib.vbn_gaussian_mixture_grid(w.data_ptr(), params.data_ptr(), b, n_particles, n_samples,
                                                  self.stddevs, self.min_scale, pdf.data_ptr(), grid.data_ptr(), sp))
            L.count_launch(1)
        return pdf, grid


def _exact_setup(vbn, query: Query):
    """Common prologue of the exact methods (gaussian_exact.py:134-164, categorical_exact.py:89-118):
    returns (b, target cpd, fixed target value | None, parents all fixed?, parent tensor [b,Dp] | None)."""
    dev = _check_model(vbn)
    b = infer_batch_size(query.evidence, query.do)
    cpds = model_cpds(vbn)
    target = query.target

    def fixed(node):
        if node in query.do:
            return query.do[node].to(device=dev, dtype=torch.float32)
        if node in query.evidence:
            return clamp_evidence(query.evidence[node].to(device=dev, dtype=torch.float32))  # clamp_obs=True
        return None

    tval = fixed(target)
    plist = list(vbn.dag.parents(target))
    pvals = [fixed(p) for p in plist]
    ok = all(v is not None for v in pvals)
    ptensor = torch.cat(pvals, dim=-1) if (ok and plist) else None
    return dev, b, cpds[target], tval, ok, ptensor


class _ExactBase:
    def _init_fallback(self, fallback, own_name: str, kwargs: dict):
        self.fallback = str(fallback).strip().lower() if fallback is not None else "none"
        self._fallback = None
        if self.fallback != "none":
            from .core import INFERENCE_REGISTRY

            if self.fallback not in INFERENCE_REGISTRY:
                raise ValueError(f"Unknown fallback inference '{fallback}'. Available: {list(INFERENCE_REGISTRY.keys())}")
            if self.fallback == own_name:
                raise ValueError(f"fallback cannot be '{own_name}'")
            self._fallback = INFERENCE_REGISTRY[self.fallback](**kwargs)

    def _fallback_infer(self, vbn, query, name: str, **kwargs):
        if self._fallback is None:
            raise RuntimeError(f"{name} cannot handle this query and has no fallback")
        self._last_exact = False
        return self._fallback.infer_posterior(vbn, query, **kwargs)


@register_inference("gaussian_exact")
class GaussianExact(_ExactBase):
    """vbn/inference/gaussian_exact.py:14-183: Normal(loc, scale) of the target on a +-stddevs grid when all
    its parents are fixed (loc/scale from the GPU parameter read-out, grid by vbn_gaussian_grid); the
    fallback method otherwise."""

    def __init__(self, n_samples: int = 200, stddevs: float = 4.0, min_scale: float = 1e-6,
                 fallback: str = "likelihood_weighting", **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.stddevs = float(stddevs)
        self.min_scale = float(min_scale)
        self._cache = {}
        self._last_exact = False
        fk = dict(kwargs)
        fk.setdefault("n_samples", self.n_samples)
        self._init_fallback(fallback, "gaussian_exact", fk)

    def infer_posterior(self, vbn, query: Query, **kwargs):
        from . import _lib as L

        n_samples = max(1, int(kwargs.get("n_samples", self.n_samples)))
        dev, b, cpd, tval, ok, ptensor = _exact_setup(vbn, query)
        if cpd.output_dim != 1:
            return self._fallback_infer(vbn, query, "gaussian_exact", **kwargs)
        if tval is not None:
            self._last_exact = True
            return torch.ones(b, 1, device=dev), tval.unsqueeze(1).expand(b, 1, -1)
        if not ok or cpd.kind not in ("linear_gaussian", "gaussian_nn", "rff_gaussian"):
            return self._fallback_infer(vbn, query, "gaussian_exact", **kwargs)
        ls = cpd.params(ptensor).reshape(-1, 2)  # [b or 1, {loc, scale}]
        if ls.shape[0] == 1 and b > 1:
            ls = ls.expand(b, 2)
        ls = ls.contiguous()
        lib = L.load()
        with torch.cuda.device(dev):
            pdf = torch.empty(b, n_samples, device=dev, dtype=torch.float32)
            samples = torch.empty(b, n_samples, 1, device=dev, dtype=torch.float32)
            L.check(lib.vbn_gaussian_grid(ls.data_ptr(), b, n_samples, self.stddevs, self.min_scale,
                                          pdf.data_ptr(), samples.data_ptr(), E._stream_ptr(dev)))
            L.count_launch(1)
        self._last_exact = True
        return pdf, samples


@register_inference("categorical_exact")
class CategoricalExact(_ExactBase):
    """vbn/inference/categorical_exact.py:14-128: class probabilities of a categorical target whose parents
    are all fixed (softmax_nn through the GPU parameter read-out, categorical_table from its table)."""

    def __init__(self, fallback: str = "likelihood_weighting", **kwargs) -> None:
        self._last_exact = False
        self._init_fallback(fallback, "categorical_exact", dict(kwargs))

    def infer_posterior(self, vbn, query: Query, **kwargs):
        dev, b, cpd, tval, ok, ptensor = _exact_setup(vbn, query)
        if tval is not None:
            self._last_exact = True
            return torch.ones(b, 1, device=dev), tval.unsqueeze(1).expand(b, 1, -1)
        if not ok or cpd.kind not in ("softmax_nn",) + TABLE_KINDS or cpd.output_dim != 1:
            return self._fallback_infer(vbn, query, "categorical_exact", **kwargs)
        if cpd.kind == "softmax_nn":
            probs = cpd.params(ptensor).reshape(-1, cpd.n_classes)
        else:
            probs = cpd.probs(ptensor).reshape(-1, cpd.n_classes).to(dev)
        if probs.shape[0] == 1 and b > 1:
            probs = probs.expand(b, -1)
        support = cpd._sample_values[0].to(device=dev, dtype=torch.float32)
        self._last_exact = True
        return probs, support.view(1, -1, 1).expand(b, -1, 1)


@register_sampling("ancestral")
class AncestralSampler:
    """vbn/sampling/ancestral.py:57-65."""

    def __init__(self, n_samples: int = 200, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self._runner = _ScheduleRunner()

    def sample(self, vbn, query: Query, n_samples: Optional[int] = None, **kwargs):
        n_samples = int(n_samples or self.n_samples)
        joint = not query.target
        q = query if not joint else Query(target=vbn.dag.topological_order()[0],
                                          evidence=query.evidence, do=query.do)
        out = self._runner.forward(vbn, q, n_samples, "anc", noise=kwargs.get("noise"),
                                   seed=kwargs.get("seed"), shard=kwargs.get("shard"), store_all=joint)
        return out["stores"] if joint else out["stores"][query.target]


@register_inference("lbp")
class LoopyBeliefPropagation:
    """vbn/inference/lbp.py:11-70.  In the reference this method is a thin wrapper: one importance_sampling (or
    monte_carlo_marginalization) pass, then a damped re-normalisation of the [B, S] weights iterated until the
    largest change drops below ``tol``, else a fresh importance_sampling pass.  The passes are the CUDA classes
    above; the damping recurrence is a few elementwise torch ops on the device-resident weights (same expressions
    and order as lbp.py:52-63, including the batch-global convergence test)."""

    def __init__(self, n_samples: int = 200, n_iters: int = 10, damping: float = 0.5,
                 fallback: str = "importance_sampling", **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.n_iters = int(n_iters)
        self.damping = float(damping)
        self.fallback = str(fallback)
        if not (0.0 <= self.damping <= 1.0):
            raise ValueError("damping must be in [0,1]")
        if self.fallback not in {"importance_sampling", "monte_carlo_marginalization"}:
            raise ValueError("fallback must be 'importance_sampling' or 'monte_carlo_marginalization'")
        self._is = ImportanceSampling(n_samples=self.n_samples)
        self._mcm = MonteCarloMarginalization(n_samples=self.n_samples)

    def infer_posterior(self, vbn, query: Query, **kwargs):
        n_samples = int(kwargs.get("n_samples", self.n_samples))
        n_iters = int(kwargs.get("n_iters", self.n_iters))
        damping = float(kwargs.get("damping", self.damping))
        tol = float(kwargs.get("tol", 1e-4))
        eps = 1e-12
        extra = {k: kwargs[k] for k in ("seed", "shard") if k in kwargs}
        noise = kwargs.get("noise") or {}
        if self.fallback == "monte_carlo_marginalization":
            pdf, target_samples = self._mcm.infer_posterior(vbn, query, n_samples=n_samples, noise=noise.get("first"), **extra)
            weights = pdf / (pdf.sum(dim=-1, keepdim=True) + eps)
        else:
            weights, target_samples = self._is.infer_posterior(vbn, query, n_samples=n_samples, noise=noise.get("first"), **extra)
        converged = False
        for _ in range(max(n_iters, 0)):
            w_new = torch.clamp(weights, min=eps)
            w_new = w_new / (w_new.sum(dim=-1, keepdim=True) + eps)
            msg = damping * w_new + (1.0 - damping) * weights
            msg = msg / (msg.sum(dim=-1, keepdim=True) + eps)
            delta = (msg - weights).abs().max().item()
            weights = msg
            if delta < tol:
                converged = True
                break
        if not converged:
            return self._is.infer_posterior(vbn, query, n_samples=n_samples, noise=noise.get("retry"), **extra)
        return weights.detach(), target_samples.detach()


@register_sampling("gibbs")
class GibbsSampler:
    """vbn/sampling/gibbs.py:12-92.  One chain per query; the whole chain -- ancestral initial
    state, burn_in + n_samples * max(n_steps, 1) sweeps of 8-candidate propose / score / select per latent node --
    is ONE kernel launch (plan.compile_gibbs): a row is a chain, node values stay in shared memory.

    Reference behaviour kept on purpose: the chain's states are collected as VIEWS of the live state tensor
    (gibbs.py:83-91), so the returned [B, n_samples, D] tensor holds the FINAL target state n_samples times.
    Deviation: with B > 1 and a latent root the reference raises (IndexError / a torch.cat size error: root candidates are [1, 8, D]);
    here every chain draws its own root candidates."""

    def __init__(self, n_samples: int = 200, burn_in: int = 10, n_steps: int = 1, **kwargs) -> None:
        self.n_samples = int(n_samples)
        self.burn_in = int(burn_in)
        self.n_steps = int(n_steps)
        self.n_candidates = 8
        self._cache: Dict[tuple, E.DevicePlan] = {}

    def _plan(self, vbn, query: Query, total_steps: int, inject: bool) -> E.DevicePlan:
        cpds = model_cpds(vbn)
        fp = tuple((c._uid, c._version) for c in cpds.values())
        key = (fp, query.target, tuple(sorted(query.evidence)), tuple(sorted(query.do)), int(total_steps),
               self.n_candidates, bool(inject), str(vbn.device))
        plan = self._cache.get(key)
        if plan is None:
            topo, parents = _topology(vbn)
            children: Dict[str, List[str]] = {n: [] for n in topo}
            for p, c in vbn.dag.edges():
                children[p].append(c)
            prog = compile_gibbs(topo, parents, children, cpds, list(query.evidence) + list(query.do), query.target,
                                 total_steps, self.n_candidates, inject=inject)
            plan = E.DevicePlan(prog, vbn.device)
            if len(self._cache) > 64:
                self._cache.clear()
            self._cache[key] = plan
        return plan

    def sample(self, vbn, query: Query, n_samples: Optional[int] = None, **kwargs):
        n_samples = int(n_samples or self.n_samples)
        dev = _check_model(vbn)
        b = infer_batch_size(query.evidence, query.do)
        total_steps = self.burn_in + n_samples * max(self.n_steps, 1)
        noise = kwargs.get("noise")
        plan = self._plan(vbn, query, total_steps, inject=noise is not None)
        prog = plan.program
        with torch.cuda.device(dev):
            fixed = _ScheduleRunner._build_fixed_table(plan, query, b, clamp_obs=False, shard=None) if plan.program.n_fixed_cols else None
            out = torch.empty(b, 1, prog.dims[query.target], device=dev, dtype=torch.float32)
            flag = torch.zeros(1, device=dev, dtype=torch.int32)
            arrays = [] if noise is None else [noise[kind][node] for kind, node in prog.noise_keys]
            seed = kwargs.get("seed")
            plan.run(b, 1, fixed=fixed, stores=[out], noise=arrays, seed=E.draw_seed() if seed is None else seed,
                     error_flag=flag)
            cpds = model_cpds(vbn)
            if any(c.kind in ("softmax_nn",) + TABLE_KINDS for c in cpds.values()) and int(flag.item()) != 0:
                raise ValueError(_flag_message(cpds.values()))
        return out.expand(b, n_samples, -1).contiguous()
