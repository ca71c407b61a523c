#!/usr/bin/env python
"""Benchmark of the posterior-inference hot path (BASELINE.json metric: posterior samples/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg2|cfg3|cfg4] [--impl reference]

A *step* is one ``infer_posterior`` call (importance sampling, with its likelihood-weighting
fallback when ESS is low -- exactly the reference semantics) over one batch of synthetic queries.

Default workload = BASELINE config 5 (1000-node random DAG, linear_gaussian + mdn CPDs, 10 000 queries x 4096
samples).  At N = 1 the whole configuration runs on the one GPU; at N > 1 every rank owns 1250 queries x 4096 samples
("weak" scaling over queries, no data-path collective; N = 8 is the configuration as stated).

One JSON line is printed by rank 0.  ``value`` = device-timed throughput with the evidence already resident in HBM;
``e2e`` = the same metric through the public API with HOST evidence buffers and the (weights, samples) result copied
back to pinned host memory inside the timed region; ``e2e_summary`` = the same with the posterior summarised on the
device (``summary=True``: weighted mean / std / ESS per query, nothing of size [B, S] crosses PCIe).
At N = 1 the line also carries ``others``: short runs of cfg2 / cfg3 / cfg4 with their own roofline, e2e and clocks.
Under torchrun (N > 1) it also carries ``strong_scaling``: BASELINE cfg2 (64 queries x 1M samples, samples sharded,
the one path with a data-path collective) on 1 GPU and on N GPUs.

``--impl reference`` times the UNMODIFIED reference package (``baseline/_ref``; oracle/reference_arm.py loads the same
parameters into the reference's own CPD classes) through its public API on this box's host cores; if the package is
absent it falls back to the oracle port (bit-identical arithmetic, tests/test_oracle_pin.py) and says so.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (description, queries at N = 1, queries per rank at N > 1, samples, method)
    "cfg5": ("cfg5: 1000-node random DAG, even=linear_gaussian odd=mdn(K=3,[32,32]), importance_sampling, "
             "evidence on last 5 nodes, target n500", 10_000, 1250, 4096, "importance_sampling"),
    "cfg2": ("cfg2: 50-node linear_gaussian chain, importance_sampling, evidence x49, target x25",
             64, 64, 1_000_000, "importance_sampling"),
    "cfg4": ("cfg4: kde CPD p->y with 200k stored points, CPDHandle.log_prob over 1M query rows",
             1_000_000, 1_000_000, 1, "kde_log_prob"),
    "cfg3": ("cfg3: ALARM (37 nodes) softmax_nn discrete CPDs, likelihood_weighting, 4 evidence nodes, "
             "target LVFAILURE", 4096, 4096, 16384, "likelihood_weighting"),
}


def build_workload(name: str, n_queries_total: int):
    from vectorizedbayesiannetwork_b200 import synthetic as S

    g = torch.Generator().manual_seed(1)
    if name == "cfg5":
        spec = S.random_dag_lg_mdn(1000, seed=0)
        ev_nodes = spec["nodes"][-5:]
        evidence = {n: 0.3 * torch.randn(n_queries_total, 1, generator=g) for n in ev_nodes}
        target = "n500"
    elif name == "cfg2":
        spec = S.lg_chain(50)
        # x49 marginal: mean 4.9, var 1 + 49*0.25; evidence within +-1.5 sigma
        sd = (1 + 49 * 0.25) ** 0.5
        evidence = {"x49": 4.9 + sd * (torch.rand(n_queries_total, 1, generator=g) * 3 - 1.5)}
        target = "x25"
    elif name == "cfg3":
        spec = S.alarm_softmax(seed=0)
        evidence = {}
        for n in ("HRBP", "BP", "EXPCO2", "PRESS"):
            card = S.ALARM[n][0]
            evidence[n] = torch.randint(0, card, (n_queries_total, 1), generator=g).float()
        target = "LVFAILURE"
    else:
        raise SystemExit(f"unknown workload {name}")
    return spec, target, evidence


def algorithmic_work(program) -> dict:
    """Per-row algorithmic flops (2*MAC of every affine map the schedule evaluates) and the
    SURVEY 8(d) per-level byte count (what a one-launch-per-level SoA design would move)."""
    from vectorizedbayesiannetwork_b200 import _lib as L

    flops = 0
    level_bytes = 8  # final reduction: read logw + write weight
    for op in program.ops:
        kind, dim, dp = int(op["kind"]), int(op["dim"]), int(op["n_par"])
        if kind == L.OP_NONE:
            continue
        sampled = (int(op["flags"]) & 3) == L.SRC_SAMPLE
        if kind == L.OP_LG:
            flops += 2 * dp * dim
        elif kind in (L.OP_GNN, L.OP_MDN, L.OP_SNN) and int(op["n_layers"]) > 0:
            dims = [dp] + [int(x) for x in op["layer_dim"][: int(op["n_layers"])]]
            flops += sum(2 * a * b for a, b in zip(dims[:-1], dims[1:]))
        level_bytes += 4 * (dp + dim) if sampled else 4 * dp + 8
    return {"flops_per_row": flops, "level_bytes_per_row": level_bytes}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md) through NVML in a
    background thread (an `nvidia-smi -lms` child process perturbs short steps: its polling stalled
    launches by tens of ms); falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml as N

            N.nvmlInit()
            self._nvml = N
            # honour CUDA_VISIBLE_DEVICES-less boxes: device index == NVML index here (one box, all GPUs visible)
            self._h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(self._h, N.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
        return self

    def _loop(self):
        N = self._nvml
        names = {
            getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self._h, N.NVML_CLOCK_SM)))
                try:
                    mask = N.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    mask = N.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def mark(self):
        """Keep only the samples taken from now on (start of a timed region)."""
        self.sm.clear()
        self.reasons.clear()

    def read(self) -> dict:
        if self._nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=20).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML, 20 ms period"}

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)


_PEAKS = {}


def measure_fma_peak(dev) -> dict:
    """Measured FP32 FMA peak (TFLOP/s) of this GPU with our probe kernel: scalar FFMA and FFMA2."""
    if "fma" in _PEAKS:
        return _PEAKS["fma"]
    from vectorizedbayesiannetwork_b200 import _lib as L

    lib = L.load()
    out = {}
    with torch.cuda.device(dev):
        scratch = torch.zeros(4, device=dev)
        blocks = 148 * 8
        for mode, name in ((0, "ffma"), (1, "ffma2")):
            iters = 20000
            lib.vbn_fma_peak(mode, 100, blocks, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
            best = 0.0
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.vbn_fma_peak(mode, iters, blocks, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = max(best, 2.0 * 16 * iters * 256 * blocks / (ms * 1e-3) / 1e12)
            out[name] = round(best, 2)
    _PEAKS["fma"] = out
    return out


def measure_tf32_peak(dev) -> float:
    """Measured dense tcgen05 kind::tf32 peak (TFLOP/s) with our probe kernel (one CTA per SM)."""
    if "tf32" in _PEAKS:
        return _PEAKS["tf32"]
    from vectorizedbayesiannetwork_b200 import _lib as L

    lib = L.load()
    with torch.cuda.device(dev):
        scratch = torch.zeros(4, device=dev)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        st = torch.cuda.current_stream().cuda_stream
        L.check(lib.vbn_tf32_peak(256, sms, scratch.data_ptr(), st))
        best = 0.0
        iters = 20000
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.vbn_tf32_peak(iters, sms, scratch.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * 128 * 256 * 8 * iters * sms / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    _PEAKS["tf32"] = round(best, 1)
    return _PEAKS["tf32"]


def _captures() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}


def _traffic(workload: str, rows: int):
    cap = _captures().get(workload)
    if not cap:
        return None, None
    t = {"bytes_per_launch": round(cap["dram_bytes"] * rows / cap["rows"]),
         "source": f"ncu --set full capture at {cap['rows']} rows ({cap['file']}), scaled by rows"}
    issue = None
    if "issue_active_pct" in cap:
        issue = {"issue_active_pct": cap["issue_active_pct"], "inst_per_row": cap.get("inst_per_row"),
                 "source": f"same capture ({cap['file']}): smsp__issue_active.avg.pct_of_peak_sustained_active"}
    return t, issue


class Dist:
    def __init__(self, rank, local_rank, world):
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()

    def max_ms(self, ms: float) -> float:
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


# ------------------------------------------------------------------------------------------------------------
# one inference workload (cfg5 / cfg2 / cfg3), device-timed + end to end
# ------------------------------------------------------------------------------------------------------------
def run_inference(name: str, d: Dist, *, steps: int, warmup: int, shard_kind: str = "queries", queries=None,
                  samples=None, clocks: ClockSampler = None, active_world=None) -> dict:
    """Times ``steps`` infer_posterior calls of workload ``name``.  ``active_world``: 1 = only rank 0 works (the
    1-GPU leg of the strong-scaling pair under torchrun); None = all ranks."""
    import vectorizedbayesiannetwork_b200 as V
    from vectorizedbayesiannetwork_b200 import _lib as L
    from vectorizedbayesiannetwork_b200 import engine as E

    desc, b_one, b_rank, s, method = WORKLOADS[name]
    world = d.world if active_world is None else active_world
    working = d.rank < world
    if queries:
        b_one = b_rank = queries
    if samples:
        s = samples
    by_samples = shard_kind == "samples" and world > 1
    b_total = b_one if (world == 1 or by_samples) else b_rank * world
    b_local = b_total if (world == 1 or by_samples) else b_rank
    dev = d.dev
    warmup = max(warmup, 3)
    out = None
    if working:
        spec, target, evidence_host = build_workload(name, b_total)
        shard = V.Shard(shard_kind, d.rank, world) if world > 1 else None
        model = V.VBN.from_spec(spec, device=dev)
        model.set_inference_method(method, n_samples=s)
        evidence_dev = {k: v.to(dev) for k, v in evidence_host.items()}
        evidence_pinned = {k: v.pin_memory() for k, v in evidence_host.items()}
        q_dev = {"target": target, "evidence": evidence_dev}
        kw = {"shard": shard} if shard is not None else {}
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2
        for _ in range(warmup):
            # keep the previous result alive like the timed loop does, so the caching allocator reaches its steady
            # state before timing starts -- a cudaMalloc of a 256 MB block inside a timed step costs tens of ms
            w, smp = model.infer_posterior(q_dev, **kw)
        torch.cuda.synchronize()
    gc.collect()
    gc.disable()  # no collector pauses inside the timed steps
    step_ms, kernel_ms, fallbacks, launches = [], [], 0, 0
    if working:
        E.KERNEL_EVENTS = []
        launches0 = L.launch_count()
    d.barrier()
    if clocks is not None:
        clocks.mark()
    if working:
        events = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            w, smp = model.infer_posterior(q_dev, **kw)
            e1.record()
            events.append((e0, e1))
            fallbacks += int(bool(getattr(model._inference, "_last_fallback", False)))
    d.barrier()
    clock_info = clocks.read() if clocks is not None else None
    if working:
        launches = L.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in events]
        kernel_ms = [a.elapsed_time(b) for a, b in E.KERNEL_EVENTS]
        E.KERNEL_EVENTS = None
    total_ms = d.max_ms(sum(step_ms))

    # ---- end to end: host evidence in, host result out -------------------------------------------------
    e2e_ms = {"full": 0.0, "summary": 0.0}
    h2d = d2h = d2h_summary = 0
    if working:
        s_local = shard.local_samples(s)[0] if by_samples else s
        out_w = torch.empty(b_local, s_local, dtype=torch.float32).pin_memory()
        out_s = torch.empty(b_local, s_local, 1, dtype=torch.float32).pin_memory()
        out_sum = torch.empty(b_local, 3, dtype=torch.float32).pin_memory()
        h2d = sum(v.numel() * 4 for v in evidence_pinned.values())
        d2h = out_w.numel() * 4 + out_s.numel() * 4
        d2h_summary = out_sum.numel() * 4

        def e2e_full():
            ev = {k: v.to(dev, non_blocking=True) for k, v in evidence_pinned.items()}
            w_, s_ = model.infer_posterior({"target": target, "evidence": ev}, **kw)
            out_w.copy_(w_, non_blocking=True)
            out_s.copy_(s_, non_blocking=True)

        def e2e_summary():
            ev = {k: v.to(dev, non_blocking=True) for k, v in evidence_pinned.items()}
            st = model.infer_posterior({"target": target, "evidence": ev}, summary=True, **kw)
            out_sum[:, 0:1].copy_(st["mean"], non_blocking=True)
            out_sum[:, 1:2].copy_(st["std"], non_blocking=True)
            out_sum[:, 2].copy_(st["ess"], non_blocking=True)

        legs = {"full": e2e_full, "summary": e2e_summary}
    for leg in ("full", "summary"):
        if working:
            legs[leg]()
        d.barrier()
        ms = 0.0
        if working:
            ev_ = []
            for _ in range(steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                legs[leg]()
                e1.record()
                ev_.append((e0, e1))
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev_)
        d.barrier()
        e2e_ms[leg] = d.max_ms(ms)
    # ---- the same call with barren-node pruning (exact plan-level optimisation; NOT the headline) -----------
    pruned = None
    if working:
        for _ in range(3):
            model.infer_posterior(q_dev, prune=True, **kw)
    d.barrier()
    ms = 0.0
    if working:
        ev_ = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.infer_posterior(q_dev, prune=True, **kw)
            e1.record()
            ev_.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev_)
    d.barrier()
    pruned_ms = d.max_ms(ms)
    gc.enable()
    if d.rank != 0:
        return None
    plans = list(model._inference._runner._cache.values())
    n_full = max(len(p.program.ops) for p in plans)
    n_pruned = min(len(p.program.ops) for p in plans)
    pruned = {"value": b_total * s * steps / (pruned_ms * 1e-3), "unit": "samples/s", "ms_per_step": pruned_ms / steps,
              "ops_evaluated": n_pruned, "ops_total": n_full,
              "what": "infer_posterior(..., prune=True): unobserved nodes without an observed or queried descendant "
                      "(barren nodes) are left out of the schedule; weights and samples are bit-identical to the full "
                      "walk (tests/test_reference_semantics.py::test_barren_node_pruning_is_exact). Reported beside the "
                      "headline, which walks every node like the reference does"}

    total_s = total_ms * 1e-3
    value = b_total * s * steps / total_s
    plan = max(plans, key=lambda p: len(p.program.ops))  # the full (unpruned) schedule of the timed steps
    work = algorithmic_work(plan.program)
    rows = b_local * s_local
    k_avg_ms = sum(kernel_ms) / max(len(kernel_ms), 1)
    peaks = measure_fma_peak(dev)
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = measured.get("hbm_gbs", 6650.0)
    traffic, issue = _traffic(name, rows)
    if plan.program.tc:
        achieved = work["flops_per_row"] * rows / (k_avg_ms * 1e-3) / 1e12
        tf32_peak = measure_tf32_peak(dev)
        roofline = {"bound": "tensor", "achieved": round(achieved, 3), "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": round(achieved / tf32_peak, 4), "traffic": traffic,
                    "peak_source": "tcgen05 kind::tf32 probe kernel measured in this run (MEASURED_PEAKS.json has "
                                   "only the bf16 figure; the MLP layers run as 3xTF32 for fp32 parity)",
                    "note": "achieved = ALGORITHMIC flops (2*MAC of every affine map); the tensor pipe executes 3x "
                            "that (3xTF32 split) on padded tiles (K>=8, N>=16). MMA tiles are 128x32x8, so the kernel "
                            "is bound by the per-row epilogue instruction issue, not by the tensor pipe",
                    "fp32_fma_peak_tflops": max(peaks.values()),
                    "frac_of_fp32_fma_peak": round(achieved / max(peaks.values()), 4)}
    elif plan.program.heavy:
        achieved = work["flops_per_row"] * rows / (k_avg_ms * 1e-3) / 1e12
        peak = max(peaks.values())
        roofline = {"bound": "fp32_fma", "achieved": round(achieved, 3), "peak": peak, "unit": "TFLOP/s",
                    "frac": round(achieved / peak, 4), "traffic": traffic,
                    "peak_source": "FFMA/FFMA2 probe kernel measured in this run (MEASURED_PEAKS.json has no fp32 "
                                   "figure; SURVEY 8d: FFMA path -> measured FP32 FMA peak)",
                    "fma_peaks_tflops": peaks}
    else:
        achieved = work["level_bytes_per_row"] * rows / (k_avg_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(achieved / hbm_peak, 4), "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if measured else "fallback 6650",
                    "note": "algorithmic bytes = SURVEY 8(d) per-level SoA formula, an EQUIVALENCE: the fused kernel "
                            "keeps node columns in shared memory, so real DRAM traffic is only the stored columns and "
                            "the kernel is instruction-issue bound (see `issue`)"}
    if issue:
        roofline["issue"] = issue
    roofline.update({"kernel": "vbn::tc::schedule_tc_kernel" if plan.program.tc else "vbn::schedule_kernel",
                     "kernel_ms_avg": round(k_avg_ms, 4), "kernel_launches_timed": len(kernel_ms),
                     "rows_per_launch": rows, "flops_per_row": work["flops_per_row"],
                     "level_bytes_per_row": work["level_bytes_per_row"],
                     "kernel_share_of_step": round(sum(kernel_ms) / max(sum(step_ms), 1e-9), 4)})
    jobs = b_total * s * steps
    return {
        "metric": "posterior_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": total_s / steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if by_samples else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "method": method, "queries_per_gpu": b_local, "queries_total": b_total,
                   "samples_per_query": s, "sharding": shard_kind if world > 1 else "none",
                   "l2": "256 MB buffer rewritten between timed steps (L2 flush)",
                   "is_fallback_steps": fallbacks, "weights": "random-init (nn.Linear default), seeded; the "
                   "reference arm loads the same tensors into the reference's own CPD classes"},
        "queries_per_sec": b_total * steps / total_s,
        "step_ms": [round(x, 3) for x in step_ms], "kernel_ms": [round(x, 3) for x in kernel_ms],
        "e2e": {"value": jobs / (e2e_ms["full"] * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "e2e_summary": {"value": jobs / (e2e_ms["summary"] * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h_summary,
                        "what": "infer_posterior(..., summary=True): weighted mean / std / ESS per query reduced on "
                                "the device (VBN._posterior_stats semantics); nothing of size [B,S] leaves the GPU"},
        "gpu_launches": launches, "clocks": clock_info, "roofline": roofline, "pruned": pruned,
    }


# ------------------------------------------------------------------------------------------------------------
# cfg4: KDE conditional density
# ------------------------------------------------------------------------------------------------------------
def kde_inputs(n_rows: int):
    g = torch.Generator().manual_seed(1)
    return torch.randn(n_rows, 1, generator=g), torch.randn(n_rows, 1, generator=g)


def run_kde(d: Dist, *, steps: int, warmup: int, rows=None, n_points=None, clocks: ClockSampler = None) -> dict:
    """cfg4: conditional KDE density over query rows (rows sharded across ranks, no collective)."""
    import vectorizedbayesiannetwork_b200 as V
    from vectorizedbayesiannetwork_b200 import _lib as L
    from vectorizedbayesiannetwork_b200 import synthetic as S

    desc, rows_one, rows_rank, _, method = WORKLOADS["cfg4"]
    rows_rank = rows or (rows_one if d.world == 1 else rows_rank)
    dev, rank, world = d.dev, d.rank, d.world
    n_points = n_points or 200_000
    spec = S.kde_pair(n_points)
    model = V.VBN.from_spec(spec, device=dev)
    handle = model.get_cpd("y")
    x_h, p_h = kde_inputs(rows_rank * world)
    x_h, p_h = x_h[rank * rows_rank:(rank + 1) * rows_rank], p_h[rank * rows_rank:(rank + 1) * rows_rank]
    x_d, p_d = x_h.to(dev), p_h.to(dev)
    x_p, p_p = x_h.pin_memory(), p_h.pin_memory()
    out_h = torch.empty(rows_rank, 1).pin_memory()
    warmup = max(warmup, 3)
    for _ in range(warmup):
        handle.log_prob(x_d, p_d)
    l0 = L.launch_count()
    d.barrier()
    if clocks is not None:
        clocks.mark()
    ev = []
    for _ in range(steps):  # inputs (8 MB) + stored points (1.6 MB) are L2-sized; the kernel is MUFU bound
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        handle.log_prob(x_d, p_d)
        e1.record()
        ev.append((e0, e1))
    d.barrier()
    launches = L.launch_count() - l0
    clock_info = clocks.read() if clocks is not None else None
    tot = d.max_ms(sum(a.elapsed_time(b) for a, b in ev))
    ev2 = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lp = handle.log_prob(x_p.to(dev, non_blocking=True), p_p.to(dev, non_blocking=True))
        out_h.copy_(lp, non_blocking=True)
        e1.record()
        ev2.append((e0, e1))
    d.barrier()
    tot2 = d.max_ms(sum(a.elapsed_time(b) for a, b in ev2))
    if rank != 0:
        return None
    sec = tot * 1e-3
    rows_total = rows_rank * world
    value = rows_total * steps / sec
    pairs_per_s = rows_rank * n_points * steps / sec  # per GPU
    sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.0
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mufu_peak = 16 * sms * sm_mhz * 1e6  # ex2/s
    traffic, _ = _traffic("cfg4", rows_rank)
    return {
        "metric": "density_rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": sec / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "method": method, "rows_per_gpu": rows_rank, "rows_total": rows_total,
                   "stored_points": n_points, "sharding": "rows" if world > 1 else "none",
                   "l2": "working set (9.6 MB) is L2 resident by nature; kernel is MUFU bound, no flush needed"},
        "pair_evals_per_sec_per_gpu": pairs_per_s,
        "e2e": {"value": rows_total * steps / (tot2 * 1e-3), "unit": "rows/s",
                "h2d_bytes_per_step": 8 * rows_rank, "d2h_bytes_per_step": 4 * rows_rank},
        "gpu_launches": launches, "clocks": clock_info,
        "roofline": {"bound": "mufu", "achieved": round(2 * pairs_per_s / 1e12, 4), "peak": round(mufu_peak / 1e12, 4),
                     "unit": "Tex2/s", "frac": round(2 * pairs_per_s / mufu_peak, 4), "traffic": traffic,
                     "peak_source": "16 MUFU/clk/SM x SMs x median SM clock during the run (SURVEY 8d: KDE is "
                                    "exp-limited at small dims; 2 exp2 per (row, point) pair)",
                     "note": "achieved counts every exponential the algorithm needs; the kernel evaluates 1 in 4 of "
                             "them with a degree-5 polynomial on the FMA pipe (FlashAttention-4 style), so the "
                             "fraction of the MUFU-only ceiling can exceed 1",
                     "kernel": "vbn::kde_log_prob_kernel<1,1>"},
    }


# ------------------------------------------------------------------------------------------------------------
# the reference arm (CPU)
# ------------------------------------------------------------------------------------------------------------
def _cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank: int) -> None:
    """CPU arm on the box's host cores: the unmodified reference package (baseline/_ref) through its own public API,
    parameters = the same tensors the B200 arm uses; the oracle port only if the package is missing.  A step is a
    bounded sample of the workload (cost is linear in queries: two sample sizes are timed first to show it), sized
    so that the whole --steps K --warmup W run ends within a couple of minutes."""
    if rank != 0:
        return
    from oracle import reference_arm as R
    from oracle import vbn_oracle as O
    from vectorizedbayesiannetwork_b200 import synthetic as S

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    have_ref = R.reference_path() is not None
    kind = "reference" if have_ref else "port"
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    budget_s = 100.0

    if args.workload == "cfg4":
        n_points = args.samples or 200_000
        spec = S.kde_pair(n_points)
        if have_ref:
            handle = R.reference_from_spec(spec).get_cpd("y")
            fn = lambda x, p: handle.log_prob(x, {"p": p})
        else:
            fn = lambda x, p: O.kde_log_prob(spec["cpds"]["y"], x, p)
        unit, metric, per = "rows/s", "density_rows_per_sec", 1

        def timed(n):
            x, p = kde_inputs(n)
            t0 = time.perf_counter()
            fn(x, p)
            return time.perf_counter() - t0

        sizes = (512, 2048)
        sample_of = lambda n: f"{n} query rows x {n_points} stored points"
        method, desc = "kde_log_prob", WORKLOADS["cfg4"][0]
        cap = 8192
    else:
        desc, _, _, s_full, method = WORKLOADS[args.workload]
        sample_s = args.samples or {"cfg5": 4096, "cfg2": 65536, "cfg3": 4096}[args.workload]
        sizes = {"cfg5": (4, 16), "cfg2": (2, 8), "cfg3": (32, 128)}[args.workload]
        cap = sizes[1]
        spec, target, _ = build_workload(args.workload, 1)
        if have_ref:
            model = R.reference_from_spec(spec)
            model.set_inference_method(method, n_samples=sample_s)
            fn = lambda q: model.infer_posterior(q)
        else:
            ofn = O.importance_sampling if method == "importance_sampling" else O.likelihood_weighting
            fn = lambda q: ofn(spec, q, sample_s)
        unit, metric, per = "samples/s", "posterior_samples_per_sec", sample_s

        def timed(n):
            _, tgt, ev = build_workload(args.workload, n)
            q = {"target": tgt, "evidence": ev}
            t0 = time.perf_counter()
            fn(q)
            return time.perf_counter() - t0

        sample_of = lambda n: f"{n} queries x {sample_s} samples"

    torch.manual_seed(0)
    timed(sizes[0])  # first call: allocator / thread-pool warm-up
    t_small, t_big = timed(sizes[0]), timed(sizes[1])
    per_unit = t_big / sizes[1]
    n_step = int(max(sizes[0] // 2 or 1, min(cap, budget_s / (steps + warmup) / per_unit)))
    for _ in range(warmup):
        timed(n_step)
    t0 = time.perf_counter()
    for _ in range(steps):
        timed(n_step)
    dt = time.perf_counter() - t0
    value = n_step * per * steps / dt
    sample = (f"{sample_of(n_step)} per step (bounded sample of the workload; cost is linear in queries: "
              f"{sample_of(sizes[0])} took {t_small:.2f} s, {sample_of(sizes[1])} took {t_big:.2f} s = "
              f"{t_big / t_small:.2f}x for {sizes[1] / sizes[0]:.0f}x the queries)")
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "method": method, "sample": sample,
                   "reference": ("unmodified vbn 0.3.0 from baseline/_ref, VBN(device='cpu'), parameters loaded into "
                                 "the reference's own CPD classes (oracle/reference_arm.py)") if have_ref else
                                "oracle port (reference package not found on this box)",
                   "cpu": _cpu_model(), "threads": cores},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample,
                         "linearity": {"sizes": list(sizes), "seconds": [round(t_small, 3), round(t_big, 3)]}},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if per != 1:
        line["queries_per_sec"] = n_step * steps / dt
    print(json.dumps(line), flush=True)


def cpu_baseline(workload: str) -> dict:
    """The reference arm in a child process (its torch thread pool must not fight this process's)."""
    try:
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload",
                               workload, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                              timeout=900, env={k: v for k, v in os.environ.items()
                                                if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        return json.loads(proc.stdout.strip().splitlines()[-1])["cpu_baseline"]
    except Exception as exc:  # keep the bench line even if the CPU leg fails
        return {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
                "sample": f"failed: {exc}"}


def secondary_cuda_reference(dev, n_queries: int = 4) -> dict:
    """Optional, labelled extra: the unmodified reference with device='cuda' (torch eager) on
    this B200, cfg5, a few queries (its cost is linear in queries: a Python loop over nodes and query rows)."""
    try:
        from oracle import reference_arm as R

        if R.reference_path() is None:
            return {"unavailable": "reference package not on this box"}
        spec, target, ev = build_workload("cfg5", n_queries)
        model = R.reference_from_spec(spec, device=dev)
        model.set_inference_method("importance_sampling", n_samples=4096)
        q = {"target": target, "evidence": {k: v.to(dev) for k, v in ev.items()}}
        model.infer_posterior(q)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.infer_posterior(q)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"what": "unmodified reference, VBN(device='cuda'), torch eager on one B200 (secondary, not the target)",
                "value": n_queries * 4096 / dt, "unit": "samples/s",
                "sample": f"{n_queries} queries x 4096 samples, wall clock around one infer_posterior call"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the cfg2 / cfg3 / cfg4 side runs of the default line")
    ap.add_argument("--queries-per-gpu", type=int, default=None)
    ap.add_argument("--samples", type=int, default=None)
    ap.add_argument("--shard", default="queries", choices=["queries", "samples"],
                    help="multi-GPU split: queries (weak scaling, no data-path collective) or samples "
                         "(strong scaling of a fixed B x S job; per-query (m,l,q) all-gather over NCCL)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    from vectorizedbayesiannetwork_b200 import _lib as L

    d = Dist(rank, local_rank, world)
    torch.cuda.set_device(d.dev)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=d.dev)
    L.load()
    clocks = ClockSampler(local_rank).start() if rank == 0 else None  # started before any timed region

    default_line = args.workload == "cfg5" and not args.queries_per_gpu and not args.samples and args.shard == "queries"
    if args.workload == "cfg4":
        line = run_kde(d, steps=args.steps, warmup=args.warmup, rows=args.queries_per_gpu, n_points=args.samples,
                       clocks=clocks)
    else:
        line = run_inference(args.workload, d, steps=args.steps, warmup=args.warmup, shard_kind=args.shard,
                             queries=args.queries_per_gpu, samples=args.samples, clocks=clocks)

    others, strong = None, None
    if default_line and not args.no_others:
        if world == 1:
            # the other BASELINE workloads, short runs in the same process: driver-visible numbers for all four
            others = {}
            for name in ("cfg2", "cfg3"):
                others[name] = run_inference(name, d, steps=8, warmup=4, clocks=clocks)  # (~13 / ~7 ms steps)
            others["cfg4"] = run_kde(d, steps=3, warmup=3, clocks=clocks)
        else:
            # strong scaling of the one path with a data-path collective: cfg2 (64 x 1M), samples sharded
            one = run_inference("cfg2", d, steps=20, warmup=3, clocks=clocks, active_world=1)
            many = run_inference("cfg2", d, steps=20, warmup=3, shard_kind="samples", clocks=clocks)
            if rank == 0:
                strong = {"workload": WORKLOADS["cfg2"][0], "sharding": "samples", "n_gpus": world,
                          "collective": "one all_gather of [B, 3 + moments] floats per pass (NCCL)",
                          "ms_per_step_1gpu": one["ms_per_step"], "ms_per_step": many["ms_per_step"],
                          "value_1gpu": one["value"], "value": many["value"],
                          "speedup": one["ms_per_step"] / many["ms_per_step"],
                          "efficiency": one["ms_per_step"] / many["ms_per_step"] / world,
                          "kernel_share_of_step": many["roofline"]["kernel_share_of_step"],
                          "gpu_launches": many["gpu_launches"], "is_fallback_steps": many["config"]["is_fallback_steps"],
                          "clocks": many["clocks"]}
    if rank == 0:
        if others is not None:
            for name, o in others.items():
                if not args.no_cpu_baseline:
                    o["cpu_baseline"] = cpu_baseline(name)
            line["others"] = others
        if strong is not None:
            line["strong_scaling"] = strong
        line["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline(args.workload)
        if default_line and world == 1 and not args.no_cpu_baseline:
            line["secondary_baseline"] = secondary_cuda_reference(d.dev)
        print(json.dumps(line), flush=True)
    if clocks is not None:
        clocks.stop()
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
