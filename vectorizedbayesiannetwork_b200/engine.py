"""Runs compiled schedules on the GPU through the C ABI.  PyTorch is plumbing here: it owns
device memory and the current stream; every kernel is ours (libvbn_cuda.so)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .cpds import BaseCPD, KDECPD
from .plan import Program, Role, compile_schedule


KERNEL_EVENTS = None  # set to a list by bench.py to time every schedule-kernel launch


def tc_shape() -> int:
    """Geometry of the tensor-core kernel as VbnProgramDesc.tc = warpgroups | tiles per warpgroup << 8
    (VBN_TC_SHAPE=4x1|2x2; default 4x1: four 128-row warpgroups with one tile each)."""
    nwg, rpt = {"4x1": (4, 1), "2x2": (2, 2)}.get(os.environ.get("VBN_TC_SHAPE", "4x1"), (4, 1))
    return nwg | (rpt << 8)


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise L.VbnCudaError("no CUDA device: vectorizedbayesiannetwork_b200 has no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda")
    if dev.type != "cuda":
        raise L.VbnCudaError(f"device '{dev}' is not a CUDA device (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def draw_seed() -> int:
    """One value from torch's global CPU generator keys the whole call, so
    ``torch.manual_seed(k); infer(...)`` is reproducible like the reference
    (tests/test_performance_upgrades.py:23-36 relies on that)."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


class DevicePlan:
    """A Program uploaded to one device plus its native plan handle."""

    def __init__(self, program: Program, device) -> None:
        self.lib = L.load()
        self.device = require_cuda(device)
        self.program = program
        with torch.cuda.device(self.device):
            ops_i32 = np.frombuffer(program.ops.tobytes(), dtype=np.int32).reshape(len(program.ops), 32)
            self.ops = torch.from_numpy(ops_i32.copy()).to(self.device)
            self.par_slots = torch.from_numpy(program.par_slots.copy()).to(self.device)
            self.params = torch.from_numpy(program.params.copy()).to(self.device)
            self.tc_list = (torch.from_numpy(program.tc_list.copy()).to(self.device)
                            if program.tc else None)
            desc = L.ProgramDesc(
                ops_dev=self.ops.data_ptr(), n_ops=len(program.ops),
                par_slots_dev=self.par_slots.data_ptr(), n_par_slots=int(self.par_slots.numel()),
                params_dev=self.params.data_ptr(), n_params=int(self.params.numel()),
                n_slots=program.n_slots, n_scratch=program.n_scratch,
                heavy=1 if program.heavy else 0, tc=tc_shape() if program.tc else 0,
                tc_list_dev=self.tc_list.data_ptr() if program.tc else None,
                n_tc=int(program.tc_list.shape[0]) if program.tc else 0,
                tc_image_bytes=int(program.tc_list[:, 1].max()) if program.tc else 0,
                has_tables=1 if any(int(k) == L.OP_TAB for k in program.ops["kind"]) else 0,
                rows_per_thread=int(os.environ.get("VBN_ROWS_PER_THREAD", "0")),  # 0: 4 rows per thread (measured best
                # for chains and, since the plain table op became one compact body, for table schedules too)
            )
            handle = C.c_void_p()
            L.check(self.lib.vbn_plan_create(C.byref(desc), C.byref(handle)))
            self.handle = handle
            self._tables: Dict[tuple, torch.Tensor] = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.vbn_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    @staticmethod
    def _view_row(t: torch.Tensor, n_rows: int):
        """(ptr, row_stride, dim_stride) of a [B,S,D] / [R,D] tensor addressed by r = b*S+s."""
        if t.dim() == 3:
            b, s, _ = t.shape
            if b * s != n_rows:
                raise ValueError(f"tensor rows {b * s} != run rows {n_rows}")
            if b > 1 and t.stride(0) != s * t.stride(1):
                raise ValueError("row view is not collapsible; pass a contiguous tensor")
            return t.data_ptr(), int(t.stride(1) if s > 1 or b == 1 else t.stride(0)), int(t.stride(2))
        if t.dim() == 2:
            if t.shape[0] != n_rows:
                raise ValueError(f"tensor rows {t.shape[0]} != run rows {n_rows}")
            return t.data_ptr(), int(t.stride(0)), int(t.stride(1))
        raise ValueError("expected a 2-D or 3-D tensor")

    def run(self, n_queries: int, n_samples: int, *, fixed: Optional[torch.Tensor] = None,
            inputs: Sequence[torch.Tensor] = (), stores: Sequence[torch.Tensor] = (),
            noise: Sequence[Dict[str, torch.Tensor]] = (), logw: Optional[torch.Tensor] = None,
            logp: Optional[torch.Tensor] = None, logp_as_pdf: bool = False, seed: int = 0,
            call_offset: int = 0, query_offset: int = 0, sample_offset: int = 0,
            error_flag: Optional[torch.Tensor] = None, logw_accumulate: bool = False,
            seg: Optional[torch.Tensor] = None, seg_slot: int = -1, seg_classes: int = 0) -> None:
        """``seg``: [B, P, 16] float32 record buffer (zero-filled, P >= ceil(S/32) + 1) -> the kernel also emits the
        per-warp partial records of the fused weight reduction (VbnRunDesc.seg_dev); with it the log-weight buffer is
        optional."""
        p = self.program
        n_rows = int(n_queries) * int(n_samples)
        if len(inputs) != len(p.inputs) or len(stores) != len(p.stores) or len(noise) != len(p.noise):
            raise ValueError("inputs/stores/noise do not match the compiled schedule")
        if p.n_fixed_cols and (fixed is None or tuple(fixed.shape) != (p.n_fixed_cols, n_queries)):
            raise ValueError(f"fixed table must be [{p.n_fixed_cols}, {n_queries}]")
        if p.needs_logw and logw is None and seg is None:
            raise ValueError("schedule accumulates log-weights: logw buffer required")
        if seg is not None and (seg.dtype != torch.float32 or seg.dim() != 3 or seg.shape[0] != n_queries
                                or seg.shape[1] < (int(n_samples) + 31) // 32 + 1 or seg.shape[2] != 16
                                or not seg.is_contiguous()):
            raise ValueError("seg must be a contiguous float32 [B, >= ceil(S/32)+1, 16] tensor")
        if p.needs_logp and logp is None:
            raise ValueError("schedule evaluates a density: logp buffer required")
        rows: List[List[int]] = []
        for t in list(inputs) + list(stores):
            if t.dtype != torch.float32 or t.device != self.device:
                raise ValueError("views must be float32 tensors on the plan's device")
            rows.append(list(self._view_row(t, n_rows)))
        keep = []
        for nz in noise:
            ent = []
            for key, dt in (("eps", torch.float32), ("u", torch.float32), ("idx", torch.int32)):
                t = nz.get(key)
                if t is None:
                    ent.append(0)
                else:
                    t = t.to(device=self.device, dtype=dt).contiguous()
                    keep.append(t)
                    ent.append(t.data_ptr())
            rows.append(ent)
        with torch.cuda.device(self.device):
            # The view table (pointers + strides of this call's tensors) lives on the device.  The caching allocator
            # hands a steady-state caller the same blocks call after call, so the table of a few recent calls is kept:
            # no pageable host->device copy on the per-call path (it is immutable once uploaded).
            key = tuple(tuple(r) for r in rows)
            table = self._tables.get(key)
            if table is None:
                table = torch.tensor(rows if rows else [[0, 0, 0]], dtype=torch.int64).to(self.device)
                if len(self._tables) >= 16:
                    self._tables.pop(next(iter(self._tables)))
                self._tables[key] = table
            base = table.data_ptr()
            n_in, n_st = len(inputs), len(stores)
            run = L.RunDesc(
                n_queries=int(n_queries), n_samples=int(n_samples),
                query_offset=int(query_offset), sample_offset=int(sample_offset),
                seed=int(seed) & (2**64 - 1), call_offset=int(call_offset),
                fixed_dev=fixed.data_ptr() if fixed is not None else None,
                inputs_dev=base if n_in else None,
                stores_dev=base + 24 * n_in if n_st else None,
                noise_dev=base + 24 * (n_in + n_st) if len(noise) else None,
                logw_dev=logw.data_ptr() if logw is not None else None,
                logp_dev=logp.data_ptr() if logp is not None else None,
                logp_as_pdf=1 if logp_as_pdf else 0, logw_accumulate=1 if logw_accumulate else 0,
                error_flag_dev=error_flag.data_ptr() if error_flag is not None else None,
                seg_dev=seg.data_ptr() if seg is not None else None,
                seg_per_query=int(seg.shape[1]) if seg is not None else 0,
                seg_slot=int(seg_slot), seg_classes=int(seg_classes),
            )
            events = KERNEL_EVENTS
            if events is not None:  # bench.py: CUDA events on the launching stream around the kernel
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            L.check(self.lib.vbn_run_forward(self.handle, C.byref(run), _stream_ptr(self.device)))
            if events is not None:
                e1.record()
                events.append((e0, e1))
            L.count_launch(1)
        # `table` / `keep` are stream-ordered torch allocations: safe to drop after the launch


def stream_draws(kind: str, n_values: int, n_queries: int, n_samples: int, *, seed: int, shared: bool = False,
                 call_offset: int = 0, query_offset: int = 0, sample_offset: int = 0, device=None) -> torch.Tensor:
    """Replay hook (vbn_stream_draws): the first ``n_values`` normals / uniforms / raw generator words of every row's
    stream for a run with this (seed, call_offset, offsets), through the same generator code as the schedule
    kernels.  kind: "normal" | "uniform" | "bits".  Returns [ceil4(n_values), Bn, S] (float32; int32 bit patterns
    for "bits"), Bn = 1 for the shared streams (roots of LW / MCM / ancestral passes)."""
    lib = L.load()
    dev = require_cuda(device)
    n_blocks = max(1, (int(n_values) + 3) // 4)
    bn = 1 if shared else int(n_queries)
    with torch.cuda.device(dev):
        out = torch.empty(4 * n_blocks, bn, int(n_samples), device=dev, dtype=torch.float32)
        L.check(lib.vbn_stream_draws(int(seed) & (2**64 - 1), int(call_offset),
                                     {"normal": 0, "uniform": 1, "bits": 2}[kind], 1 if shared else 0, 0, n_blocks,
                                     int(n_queries), int(n_samples), int(query_offset), int(sample_offset),
                                     out.data_ptr(), _stream_ptr(dev)))
        L.count_launch(1)
    return out.view(torch.int32) if kind == "bits" else out


# --------------------------------------------------------------------------------------------
# weight reduction
# --------------------------------------------------------------------------------------------


def pick_split(n_queries: int, n_samples: int) -> int:
    """CTAs per query for the logsumexp pass: enough CTAs to fill 148 SMs a few times over,
    but at least ~2048 samples each."""
    want = max(1, (148 * 4 + n_queries - 1) // n_queries)
    return int(max(1, min(want, (n_samples + 2047) // 2048)))


def lse_stats(logw: torch.Tensor) -> torch.Tensor:
    """[B,S] log-weights -> [B,3] (max, sum exp, sum exp^2)."""
    lib = L.load()
    b, s = logw.shape
    dev = logw.device
    split = pick_split(b, s)
    with torch.cuda.device(dev):
        partials = torch.empty(b, split, 3, device=dev, dtype=torch.float32)
        stats = torch.empty(b, 3, device=dev, dtype=torch.float32)
        sp = _stream_ptr(dev)
        L.check(lib.vbn_lse_partials(logw.data_ptr(), b, s, split, partials.data_ptr(), sp))
        L.check(lib.vbn_lse_merge(partials.data_ptr(), b, split, stats.data_ptr(), sp))
        L.count_launch(2)
    return stats


def segment_records(n_queries: int, n_samples: int, device) -> torch.Tensor:
    """Record buffer for DevicePlan.run(seg=...): [B, ceil(S/32) + 1, 16].  The CUDA kernels write every record the
    merge reads; the host-emulation test build accumulates row by row and needs it zero-filled."""
    alloc = torch.zeros if torch.device(device).type == "cpu" else torch.empty
    return alloc(int(n_queries), (int(n_samples) + 31) // 32 + 1, 16, device=device, dtype=torch.float32)


def segment_merge(records: torch.Tensor, n_samples: int, *, ess_threshold: float = 0.0, want_flag: bool = False):
    """Folds per-warp (n_samples > 0) or per-rank (n_samples == 0) records: returns (merged [B,16], stats [B,3],
    flag int32[1] or None).  merged = {m, l, q, mean, 0, M2, rows, ess, class sums[8]} (vbn_segment_merge)."""
    lib = L.load()
    dev = records.device
    b, p, _ = records.shape
    with torch.cuda.device(dev):
        merged = torch.empty(b, 16, device=dev, dtype=torch.float32)
        stats = torch.empty(b, 3, device=dev, dtype=torch.float32)
        flag = torch.zeros(1, device=dev, dtype=torch.int32) if want_flag else None
        L.check(lib.vbn_segment_merge(records.data_ptr(), b, int(n_samples), p, float(ess_threshold),
                                      merged.data_ptr(), stats.data_ptr(),
                                      flag.data_ptr() if flag is not None else None, _stream_ptr(dev)))
        L.count_launch(1)
    return merged, stats, flag


def merge_stats(gathered: torch.Tensor) -> torch.Tensor:
    """[B, n_ranks, 3] per-rank stats -> [B,3] merged (cross-GPU logsumexp merge)."""
    lib = L.load()
    b, n, _ = gathered.shape
    dev = gathered.device
    with torch.cuda.device(dev):
        out = torch.empty(b, 3, device=dev, dtype=torch.float32)
        L.check(lib.vbn_lse_merge(gathered.contiguous().data_ptr(), b, n, out.data_ptr(), _stream_ptr(dev)))
        L.count_launch(1)
    return out


def normalize_weights(logw: torch.Tensor, stats: torch.Tensor, *, normalize: bool = True,
                      eps: float = 1e-12, want_ess: bool = True):
    lib = L.load()
    b, s = logw.shape
    dev = logw.device
    with torch.cuda.device(dev):
        w = torch.empty_like(logw)
        ess = torch.empty(b, device=dev, dtype=torch.float32) if want_ess else None
        L.check(lib.vbn_weights_normalize(logw.data_ptr(), stats.data_ptr(), b, s, 1 if normalize else 0,
                                          float(eps), w.data_ptr(),
                                          ess.data_ptr() if ess is not None else None, _stream_ptr(dev)))
        L.count_launch(1)
    return w, ess


def resample_indices(weights: torch.Tensor, *, seed: int, call_offset: int, query_offset: int = 0) -> torch.Tensor:
    """idx[b, :] ~ multinomial(weights[b], S, replacement=True) (resampled_importance_sampling.py:37-38):
    per-query CDF + one Philox uniform and a binary search per output row.  int32 [B, S]."""
    lib = L.load()
    dev = weights.device
    b, s = weights.shape
    with torch.cuda.device(dev):
        cdf = torch.empty(b, s, device=dev, dtype=torch.float32)
        idx = torch.empty(b, s, device=dev, dtype=torch.int32)
        sp = _stream_ptr(dev)
        L.check(lib.vbn_row_cdf(weights.contiguous().data_ptr(), b, s, cdf.data_ptr(), sp))
        L.check(lib.vbn_resample_indices(cdf.data_ptr(), b, s, int(seed) & (2**64 - 1), int(call_offset),
                                         int(query_offset), 0, idx.data_ptr(), sp))
        L.count_launch(2)
    return idx


def gather_rows(cols: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """cols [n_cols, B*S] (node columns, SoA) -> same shape with rows re-drawn: out[c][b,s] = cols[c][b, idx[b,s]]."""
    lib = L.load()
    dev = cols.device
    b, s = idx.shape
    with torch.cuda.device(dev):
        out = torch.empty_like(cols)
        L.check(lib.vbn_gather_rows(cols.data_ptr(), out.data_ptr(), idx.data_ptr(), int(cols.shape[0]), b, s,
                                    _stream_ptr(dev)))
        L.count_launch(1)
    return out


def posterior_stats(pdf: torch.Tensor, samples: torch.Tensor, eps: float = 1e-12):
    """VBN._posterior_stats (vbn/vbn.py:483-504) on the GPU: dict(mean [B,D], std [B,D], ess [B])."""
    if pdf.dim() != 2:
        raise ValueError(f"Expected pdf with shape [B,S], got {tuple(pdf.shape)}")
    if samples.dim() != 3:
        raise ValueError(f"Expected samples with shape [B,S,D], got {tuple(samples.shape)}")
    if pdf.shape[0] != samples.shape[0] or pdf.shape[1] != samples.shape[1]:
        raise ValueError("pdf and samples shapes are incompatible.")
    lib = L.load()
    dev = require_cuda(pdf.device)
    b, s = pdf.shape
    d = int(samples.shape[2])
    pdf = pdf.to(device=dev, dtype=torch.float32).contiguous()
    samples = samples.to(device=dev, dtype=torch.float32).contiguous()
    split = pick_split(b, s)
    with torch.cuda.device(dev):
        partials = torch.empty(b, split, 10, device=dev, dtype=torch.float32)
        stats = torch.empty(b, 2 + 2 * d, device=dev, dtype=torch.float32)
        L.check(lib.vbn_posterior_stats(pdf.data_ptr(), samples.data_ptr(), b, s, d, split, float(eps),
                                        partials.data_ptr(), stats.data_ptr(), _stream_ptr(dev)))
        L.count_launch(4)
    return {"mean": stats[:, 2:2 + d], "std": stats[:, 2 + d:2 + 2 * d], "ess": stats[:, 1]}


def weighted_histogram(samples: torch.Tensor, weights: torch.Tensor, k: int) -> torch.Tensor:
    """probs [B,k] of a discrete target from (samples [B,S] or [B,S,D], weights [B,S]): the GPU form of
    _estimate_discrete_posterior_batch + _normalize_probs (benchmarking/models/vbn.py:116-121, 202-242)."""
    lib = L.load()
    dev = require_cuda(weights.device)
    b, s = weights.shape
    samples = samples.to(device=dev, dtype=torch.float32).contiguous()
    weights = weights.to(device=dev, dtype=torch.float32).contiguous()
    stride = int(samples.shape[2]) if samples.dim() == 3 else 1
    with torch.cuda.device(dev):
        probs = torch.empty(b, int(k), device=dev, dtype=torch.float32)
        L.check(lib.vbn_weighted_histogram(samples.data_ptr(), weights.data_ptr(), b, s, stride, int(k),
                                           probs.data_ptr(), _stream_ptr(dev)))
        L.count_launch(1)
    return probs


def ess_below(stats: torch.Tensor, threshold: float) -> torch.Tensor:
    lib = L.load()
    dev = stats.device
    with torch.cuda.device(dev):
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        L.check(lib.vbn_ess_below(stats.data_ptr(), stats.shape[0], float(threshold), flag.data_ptr(),
                                  _stream_ptr(dev)))
        L.count_launch(1)
    return flag


# --------------------------------------------------------------------------------------------
# CPD-level entry points (CPDHandle.sample / log_prob; vbn/core/cpd_handle.py:253-262)
# --------------------------------------------------------------------------------------------

_CPD_PLANS: Dict[tuple, DevicePlan] = {}


def _cpd_plan(cpd: BaseCPD, device, *, mode: str, parents_kind: str, x_kind: str, inject: bool) -> DevicePlan:
    key = (cpd._uid, cpd._version, str(device), mode, parents_kind, x_kind, inject)
    plan = _CPD_PLANS.get(key)
    if plan is not None:
        return plan
    topo, parents, cpds, roles = [], {}, {}, {}
    if cpd.input_dim > 0:
        class _Stub:
            output_dim = cpd.input_dim
        topo.append("__parents__")
        cpds["__parents__"] = _Stub()
        roles["__parents__"] = Role(src=parents_kind, density=False)
        parents["__x__"] = ["__parents__"]
    topo.append("__x__")
    cpds["__x__"] = cpd
    if mode == "params":
        roles["__x__"] = Role(src="sample", store=True, out_params=True)
    elif mode == "sample":
        roles["__x__"] = Role(src="sample", store=True, inject=inject)
    else:
        roles["__x__"] = Role(src=x_kind, out_logp=True)
    plan = DevicePlan(compile_schedule(topo, parents, cpds, roles), device)
    if len(_CPD_PLANS) > 256:
        _CPD_PLANS.clear()
    _CPD_PLANS[key] = plan
    return plan


def discrete_table_fn(cpd, x, parents):
    """log p(x | parents) rows for plan.compile_schedule's table builder, through the GPU CPD path."""
    return cpd_log_prob(cpd, x, parents).reshape(-1).detach().cpu()


def _as_dev(t, device) -> torch.Tensor:
    return torch.as_tensor(t).detach().to(device=device, dtype=torch.float32)


def cpd_sample(cpd: BaseCPD, parents, n_samples: int, *, noise=None, seed=None) -> torch.Tensor:
    dev = require_cuda(cpd.device)
    if cpd.input_dim == 0:
        b = 1 if parents is None else int(parents.shape[0])  # root: [1,S,D] (linear_gaussian.py:187)
        pk = "fixed_q"
        parents = None
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        parents = _as_dev(parents, dev)
        if parents.dim() not in (2, 3):
            raise ValueError(f"Expected 2D or 3D tensor, got shape {tuple(parents.shape)}")
        if parents.shape[-1] != cpd.input_dim:
            raise ValueError(f"Expected parents_dim {cpd.input_dim}, got {parents.shape[-1]}")
        b = int(parents.shape[0])
        pk = "fixed_q" if parents.dim() == 2 else "fixed_row"
        if parents.dim() == 3 and parents.shape[1] != n_samples:
            raise ValueError("3-D parents must have n_samples rows per query")
    plan = _cpd_plan(cpd, dev, mode="sample", parents_kind=pk, x_kind="fixed_row", inject=noise is not None)
    with torch.cuda.device(dev):
        out = torch.empty(b, n_samples, cpd.output_dim, device=dev, dtype=torch.float32)
        fixed, inputs = None, []
        if parents is not None:
            if pk == "fixed_q":
                fixed = parents.t().contiguous()
            else:
                inputs = [parents.contiguous()]
        plan.run(b, n_samples, fixed=fixed, inputs=inputs, stores=[out],
                 noise=[noise] if noise is not None else [],
                 seed=draw_seed() if seed is None else seed)
    return out


def cpd_params(cpd: BaseCPD, parents) -> torch.Tensor:
    """Conditional-distribution parameters per row, [B, S, width] (S = 1 for 2-D parents):
    LG / GNN loc[D], scale[D]; MDN weights[K], loc[K][D], scale[K][D]; softmax_nn probs[D][C]
    (vbn/core/cpd_handle.py:40-118).  One launch of the schedule kernel with VBN_F_OUT_PARAMS."""
    dev = require_cuda(cpd.device)
    width = cpd.param_width()
    if cpd.input_dim == 0:
        b, s, pk, parents = 1, 1, "fixed_q", None
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        parents = _as_dev(parents, dev)
        if parents.dim() not in (2, 3) or parents.shape[-1] != cpd.input_dim:
            raise ValueError(f"Expected parents [B,{cpd.input_dim}] or [B,S,{cpd.input_dim}], got {tuple(parents.shape)}")
        b = int(parents.shape[0])
        s = 1 if parents.dim() == 2 else int(parents.shape[1])
        pk = "fixed_q" if parents.dim() == 2 else "fixed_row"
    plan = _cpd_plan(cpd, dev, mode="params", parents_kind=pk, x_kind="fixed_row", inject=False)
    with torch.cuda.device(dev):
        out = torch.empty(b, s, width, device=dev, dtype=torch.float32)
        fixed, inputs = None, []
        if parents is not None:
            if pk == "fixed_q":
                fixed = parents.t().contiguous()
            else:
                inputs = [parents.contiguous()]
        plan.run(b, s, fixed=fixed, inputs=inputs, stores=[out])
    return out


def cpd_log_prob(cpd: BaseCPD, x, parents) -> torch.Tensor:
    dev = require_cuda(cpd.device)
    x = _as_dev(x, dev)
    if x.dim() == 1:
        x = x.unsqueeze(-1)
    if x.dim() not in (2, 3):
        raise ValueError(f"Expected x with 2D or 3D shape, got {tuple(x.shape)}")
    if x.shape[-1] != cpd.output_dim:
        raise ValueError(f"Expected x dim {cpd.output_dim}, got {x.shape[-1]}")
    b = int(x.shape[0])
    sx = 1 if x.dim() == 2 else int(x.shape[1])
    if cpd.input_dim == 0:
        parents = None
        sp = 1
    else:
        if parents is None:
            raise ValueError("parents cannot be None when input_dim > 0")
        parents = _as_dev(parents, dev)
        if parents.dim() == 1:
            parents = parents.unsqueeze(-1)
        if parents.shape[-1] != cpd.input_dim:
            raise ValueError(f"Expected parents_dim {cpd.input_dim}, got {parents.shape[-1]}")
        if parents.shape[0] != b:
            if parents.shape[0] == 1:
                parents = parents.expand(b, *parents.shape[1:])
            else:
                raise ValueError("x and parents batch sizes differ")
        sp = 1 if parents.dim() == 2 else int(parents.shape[1])
    s = max(sx, sp)
    if sx not in (1, s) or sp not in (1, s):
        raise ValueError("x and parents sample dims are incompatible")

    if isinstance(cpd, KDECPD) and b * s >= 4096:
        return _kde_log_prob_bulk(cpd, x, parents, b, s, dev)

    xk = "fixed_q" if x.dim() == 2 or (sx == 1 and s > 1) else "fixed_row"
    pk = "fixed_q" if parents is None or parents.dim() == 2 or (sp == 1 and s > 1) else "fixed_row"
    plan = _cpd_plan(cpd, dev, mode="log_prob", parents_kind=pk, x_kind=xk, inject=False)
    with torch.cuda.device(dev):
        out = torch.empty(b, s, device=dev, dtype=torch.float32)
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        cols, inputs = [], []
        # fixed table rows follow op order: parents first, then x
        if parents is not None:
            if pk == "fixed_q":
                cols.append(parents.reshape(b, -1).t())
            else:
                inputs.append(parents.contiguous())
        if xk == "fixed_q":
            cols.append(x.reshape(b, -1).t())
        else:
            inputs.append(x.contiguous())
        fixed = torch.cat(cols, dim=0).contiguous() if cols else None
        plan.run(b, s, fixed=fixed, inputs=inputs, logp=out, error_flag=flag)
        # only VBN_OP_SNN / VBN_OP_TAB ops ever raise the flag: softmax_nn class sets (softmax_nn.py:623-625) and the
        # strict supports of the table kinds (categorical_table.py:12-21)
        if bool(np.isin(plan.program.ops["kind"], (L.OP_SNN, L.OP_TAB)).any()) and int(flag.item()) != 0:
            raise ValueError("Found values outside discrete class set." if cpd.kind == "softmax_nn"
                             else "Found values outside support.")
    return out


def _kde_points(cpd: KDECPD, dev):
    cached = cpd._dev_points
    if cached is None or cached[0].device != dev:
        cached = (cpd._parents.to(dev).contiguous(), cpd._targets.to(dev).contiguous())
        cpd._dev_points = cached
    return cached


def _kde_log_prob_bulk(cpd: KDECPD, x, parents, b: int, s: int, dev) -> torch.Tensor:
    """Large KDE evaluations go to the tiled stand-alone kernel (vbn_kde_log_prob)."""
    cpd._check_fitted()
    lib = L.load()
    tp, ty = _kde_points(cpd, dev)
    if x.dim() == 2:
        x = x.unsqueeze(1)
    qx = x.expand(b, s, cpd.output_dim).reshape(b * s, cpd.output_dim).contiguous()
    qp = None
    if parents is not None:
        if parents.dim() == 2:
            parents = parents.unsqueeze(1)
        qp = parents.expand(b, s, cpd.input_dim).reshape(b * s, cpd.input_dim).contiguous()
    # Dp + Dx >= 8: the pairwise distance term goes to the tensor cores (VBN_KDE_TC=0/1 forces either kernel)
    want_tc = os.environ.get("VBN_KDE_TC", "1" if cpd.input_dim + cpd.output_dim >= 8 else "0") == "1"
    if want_tc:
        nbytes = C.c_int64(0)
        L.check(lib.vbn_kde_tc_workspace_bytes(int(ty.shape[0]), cpd.input_dim, cpd.output_dim, C.byref(nbytes)))
        if nbytes.value > 0:
            with torch.cuda.device(dev):
                ws = getattr(cpd, "_tc_workspace", None)
                if ws is None or ws.device != dev or ws.numel() * 4 < nbytes.value:
                    ws = cpd._tc_workspace = torch.empty((nbytes.value + 3) // 4, device=dev, dtype=torch.float32)
                    # both clouds are centred on the stored points' mean (smaller numbers in the fp32 accumulators)
                    parts = ([tp.mean(dim=0)] if cpd.input_dim else []) + [ty.mean(dim=0)]
                    cpd._tc_center = torch.cat(parts).contiguous()
                out = torch.empty(b * s, device=dev, dtype=torch.float32)
                L.check(lib.vbn_kde_log_prob_tc(
                    tp.data_ptr() if cpd.input_dim else None, ty.data_ptr(), int(ty.shape[0]), cpd.input_dim,
                    cpd.output_dim, qp.data_ptr() if qp is not None else None, qx.data_ptr(), b * s,
                    cpd.bandwidth, cpd.parent_bandwidth, cpd.min_scale, cpd._tc_center.data_ptr(), ws.data_ptr(),
                    out.data_ptr(), _stream_ptr(dev)))
                L.count_launch(3)
            return out.reshape(b, s)
    with torch.cuda.device(dev):
        out = torch.empty(b * s, device=dev, dtype=torch.float32)
        L.check(lib.vbn_kde_log_prob(
            tp.data_ptr() if cpd.input_dim else None, ty.data_ptr(), int(ty.shape[0]), cpd.input_dim,
            cpd.output_dim, qp.data_ptr() if qp is not None else None, qx.data_ptr(), b * s,
            cpd.bandwidth, cpd.parent_bandwidth, cpd.min_scale, out.data_ptr(), _stream_ptr(dev)))
        L.count_launch(1)
    return out.reshape(b, s)
