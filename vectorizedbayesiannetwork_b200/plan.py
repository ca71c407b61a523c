"""Plan compiler: (DAG, CPDs, query roles) -> flat schedule for the CUDA kernel.

Replaces the reference's per-call bookkeeping -- InferenceState / get_inference_state /
prepare_fixed_values (vbn/inference/_core.py:13-135) -- and decides, per node, what the fused
kernel does with it: draw it, clamp it to a per-query value, score it into the log-weight, or
evaluate its density.  Node values live in shared-memory *slots*; a liveness pass reuses a slot as
soon as the last reader of a node has run, so the on-chip footprint is the widest live set of the
DAG, not the node count (a 1000-node DAG with parents drawn from the previous 20 nodes needs ~25).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
from .cpds import BaseCPD, Packed, pack_table, table_eligible


@dataclass
class Role:
    """What the schedule does with one node."""

    src: str = "sample"      # "sample" | "fixed_q" (per-query value) | "fixed_row" (per-row input)
    add_logw: bool = False   # evidence: logw += log p(value | parents)
    out_logp: bool = False   # write log p(value | parents) to the logp output
    shared: bool = False     # one draw shared by all queries (root drawn with parents=None)
    store: bool = False      # write the value to an output column
    inject: bool = False     # draws come from caller-supplied noise instead of Philox
    density: bool = True     # False: the node's CPD is never evaluated (do / plain clamp)
    out_params: bool = False  # write the conditional-distribution parameters instead of drawing


@dataclass
class Program:
    ops: np.ndarray
    par_slots: np.ndarray
    params: np.ndarray
    n_slots: int
    n_scratch: int
    heavy: bool
    nodes: List[str]                       # emitted nodes, op order
    fixed_cols: Dict[str, int]             # node -> first row of the fixed[][B] table
    n_fixed_cols: int
    inputs: List[str]                      # node order of inputs[]
    stores: List[str]                      # node order of stores[]
    noise: List[str]                       # node order of noise[]
    needs_logw: bool
    needs_logp: bool
    dims: Dict[str, int] = field(default_factory=dict)
    tc: bool = False                       # ops carry tensor-core MLP images -> tcgen05 kernel
    tc_list: Optional[np.ndarray] = None   # [n_tc, 2] int32 {image float offset, image bytes}, op order
    store_widths: Dict[str, int] = field(default_factory=dict)  # floats per row of each store (dims, or read-out width)
    keep_slot: int = -1                    # slot of the node kept live to the end of the walk (fused summaries), -1 none


def tensor_cores_enabled() -> bool:
    """VBN_TC=0 keeps every MLP on the FFMA path (A/B measurements); default on."""
    return os.environ.get("VBN_TC", "1") != "0"


def tables_enabled() -> bool:
    """VBN_TABLE=0 keeps discrete nodes on the per-row MLP path (A/B measurements); default on."""
    return os.environ.get("VBN_TABLE", "1") != "0"


def _roots_enabled() -> bool:
    """VBN_MDNROOT=0 keeps parent-less MDN nodes on the generic per-row path (A/B + equivalence test)."""
    return os.environ.get("VBN_MDNROOT", "1") != "0"


def _tails_enabled() -> bool:
    """VBN_TC_TAILS=0: the tensor-core kernel fetches every op descriptor from global memory (A/B measurements)."""
    return os.environ.get("VBN_TC_TAILS", "1") != "0"


def _tabplain_enabled() -> bool:
    """VBN_TABPLAIN=0 keeps drawn table nodes on the generic lookup op (A/B + equivalence test)."""
    return os.environ.get("VBN_TABPLAIN", "1") != "0"


def _tab_plain_fields(pk: Packed, slots: Sequence[int], param_off: int, out_slot: int, last_word: int, scored: bool,
                      trusted: Sequence[bool] = ()):
    """VBN_F_TABPLAIN descriptor words (include/vbn_cuda.h) of a table op, or None when it has more than 4 classes or
    a parent's classes / the node's own values are not coded 0..k-1.  Parses the VBN_OP_TAB block (cpds._tab_params).
    ``last_word``: u_off of a drawn op, fixed_col of a scored (evidence) one."""
    P = pk.params
    c, n_cfg, cpad, strict = int(P[0]), int(P[1]), int(P[2]), int(P[3] != 0)
    if c > 4:
        return None
    words = [0] * 12
    for p in range(pk.n_par):
        pi = 4 + (4 + cpad) * p
        card, stride = int(P[pi]), int(P[pi + 1])
        if stride >= 65536 or card >= 32768 or not np.array_equal(P[pi + 4: pi + 4 + card], np.arange(card, dtype=np.float32)):
            return None
        words[p] = stride | (card << 16)
        if p < len(trusted) and trusted[p]:
            words[p] -= 1 << 31  # bit 31 (as a signed word): value is always a valid class index, no range check
    sv = 4 + (4 + cpad) * pk.n_par
    if not np.array_equal(P[sv: sv + c], np.arange(c, dtype=np.float32)):
        return None
    if scored and not np.array_equal(P[sv + cpad: sv + cpad + c], np.arange(c, dtype=np.float32)):
        return None  # class values (what an evidence value is matched against) must be 0..k-1 too
    if max(list(slots) + [out_slot]) >= 65536:
        return None
    tab4 = (sv + 2 * cpad + 2 * n_cfg * c + 3) & ~3          # cdf4[n_cfg][4], then logp4[n_cfg][4]
    words[4] = int(param_off) + tab4 + (4 * n_cfg if scored else 0)
    words[5] = c | (strict << 16)
    words[6] = int(out_slot)
    words[7] = int(last_word)
    sl = list(slots) + [0] * (4 - len(slots))
    words[8:12] = [int(x) for x in sl]  # one word per parent slot (a slot address is then one multiply-add)
    return words


def compile_schedule(topo: Sequence[str], parents: Dict[str, Sequence[str]], cpds: Dict[str, BaseCPD],
                     roles: Dict[str, Role], use_tc: Optional[bool] = None, table_fn=None,
                     keep_live: Optional[str] = None, barren: Sequence[str] = ()) -> Program:
    """``topo``: nodes to emit, in topological order (nodes absent from ``roles`` are skipped).
    ``table_fn(cpd, x, parents) -> log_prob``: lets discrete nodes with all-discrete parents be
    compiled into lookup tables (VBN_OP_TAB); None keeps them on the MLP path.
    ``keep_live``: a node whose value slot must survive to the end of the walk (the kernel's fused summary reads
    it there: VbnRunDesc.seg_slot).
    ``barren``: nodes (with roles, like every other) whose ops are NOT emitted: unobserved nodes without an observed
    or queried descendant cannot influence the weights or the target.  They keep their place in the random streams, so
    the emitted nodes draw exactly what they draw in the full schedule (pruned and full runs agree bit for bit)."""
    if use_tc is None:
        use_tc = tensor_cores_enabled()
    any_tc = False
    lg_fast_ops: List[int] = []
    tc_list: List[tuple] = []
    tc_ops: List[tuple] = []  # (op index, weight image) of every op with a tensor-core image
    plain_drawn: set = set()  # nodes drawn by VBN_F_TABPLAIN ops
    order = [n for n in topo if n in roles]
    # Roots whose draw is shared by all queries (LW / MCM / ancestral passes) go first -- still a topological
    # order -- so that consecutive shared draws are served from one generator block (csrc cached_uniform /
    # cached_normal).  Bounded: every hoisted root stays live until its first reader, which costs a slot each.
    hoist = [n for n in order if roles[n].src == "sample" and roles[n].shared and not parents.get(n)]
    if 1 < len(hoist) <= 32:
        hs = set(hoist)
        order = hoist + [n for n in order if n not in hs]
    # stream positions are those of the FULL schedule (barren nodes included)
    stream_base: Dict[str, tuple] = {}
    n_acc = u_acc = 0
    for n in order:
        if roles[n].density and roles[n].src == "sample":
            pk0 = cpds[n].pack()
            stream_base[n] = (n_acc, u_acc)
            n_acc += pk0.n_normals
            u_acc += pk0.n_uniforms
    if barren:
        drop = set(barren)
        order = [n for n in order if n not in drop]
    index = {n: i for i, n in enumerate(order)}
    packed: Dict[str, Optional[Packed]] = {}
    dims: Dict[str, int] = {}
    for n in order:
        r = roles[n]
        evaluates = r.density and (r.src == "sample" or r.add_logw or r.out_logp)
        packed[n] = cpds[n].pack() if evaluates else None
        if evaluates and table_fn is not None and tables_enabled():
            plist = list(parents.get(n, ()))
            # parents must carry class values: drawn discrete nodes, or evidence that is itself scored
            # (an off-class evidence value then raises through that parent's own density)
            ok = all(p in roles and roles[p].density and (roles[p].src == "sample" or roles[p].add_logw)
                     for p in plist)
            if ok and table_eligible(cpds[n], [cpds[p] for p in plist]):
                packed[n] = pack_table(cpds[n], [cpds[p] for p in plist], table_fn)
        dims[n] = int(cpds[n].output_dim)

    # ---- liveness: last op that reads each node's value -------------------------------------
    last_use = {n: index[n] for n in order}
    for n in order:
        if packed[n] is None:
            continue
        for p in parents.get(n, ()):
            if p not in index:
                raise ValueError(f"node '{n}' is evaluated but its parent '{p}' is not in the schedule")
            last_use[p] = max(last_use[p], index[n])

    if keep_live is not None:
        last_use[keep_live] = len(order)  # never released

    # ---- slot allocation (first fit, D consecutive slots per node) --------------------------
    occupied: List[bool] = []
    slot_of: Dict[str, int] = {}
    expiring: Dict[int, List[str]] = {}
    for n, lu in last_use.items():
        expiring.setdefault(lu, []).append(n)

    def alloc(d: int) -> int:
        run = 0
        for i, busy in enumerate(occupied):
            run = 0 if busy else run + 1
            if run == d:
                start = i - d + 1
                for j in range(start, start + d):
                    occupied[j] = True
                return start
        start = len(occupied) - run  # extend the free tail
        occupied.extend([True] * (d - run))
        for j in range(start, start + d):
            occupied[j] = True
        return start

    # ---- parameter blob ------------------------------------------------------------------------
    blob: List[np.ndarray] = []
    blob_len = 0
    param_off: Dict[object, int] = {}

    ops = np.zeros(len(order), dtype=L.OP_DTYPE)
    par_slots: List[int] = []
    fixed_cols: Dict[str, int] = {}
    n_fixed = 0
    inputs: List[str] = []
    stores: List[str] = []
    noise: List[str] = []
    n_scratch = 0
    heavy = False
    needs_logw = False
    needs_logp = False

    for i, n in enumerate(order):
        r = roles[n]
        pk = packed[n]
        d = dims[n]
        slot_of[n] = alloc(d)
        op = ops[i]
        op["dim"] = d
        op["out_slot"] = slot_of[n]
        flags = {"sample": L.SRC_SAMPLE, "fixed_q": L.SRC_FIXED_Q, "fixed_row": L.SRC_FIXED_ROW}[r.src]
        op["fixed_col"] = -1
        if r.src == "fixed_q":
            fixed_cols[n] = n_fixed
            op["fixed_col"] = n_fixed
            n_fixed += d
        elif r.src == "fixed_row":
            op["fixed_col"] = len(inputs)
            inputs.append(n)
        op["store_idx"] = -1
        if r.store:
            op["store_idx"] = len(stores)
            stores.append(n)
        op["noise_idx"] = -1
        if pk is None:
            op["kind"] = L.OP_NONE
            op["flags"] = flags
        else:
            plist = list(parents.get(n, ()))
            pdim = sum(dims[p] for p in plist)
            if pdim != pk.n_par:
                raise ValueError(f"node '{n}': CPD expects {pk.n_par} parent dims, DAG provides {pdim}")
            op["kind"] = pk.kind
            op["n_par"] = pk.n_par
            op["par_off"] = len(par_slots)
            for p in plist:
                par_slots.extend(range(slot_of[p], slot_of[p] + dims[p]))
            key = id(pk)
            if key not in param_off:
                pad = (-blob_len) % 4
                if pad:
                    blob.append(np.zeros(pad, np.float32))
                    blob_len += pad
                param_off[key] = blob_len
                blob.append(pk.params.astype(np.float32, copy=False))
                blob_len += pk.params.size
            op["param_off"] = param_off[key]
            if use_tc and pk.tc_blob is not None:
                # the image itself (weights + the descriptors of the ops that follow, see below) is placed
                # once the whole op table is known
                op["tc"][:] = [1, 0, pk.tc_k1, pk.tc_n3]
                tc_ops.append((i, pk.tc_blob.astype(np.float32, copy=False)))
                any_tc = True
            op["n_layers"] = pk.n_layers
            op["act"] = pk.act
            op["n_out"] = pk.n_out
            op["k"] = pk.k
            ld = list(pk.layer_dim)[: L.MAX_LAYERS]
            op["layer_dim"][: len(ld)] = ld
            op["aux"][:] = pk.aux
            if pk.n_layers == 3 and list(pk.layer_dim[:2]) == [32, 32]:
                flags |= L.F_FAST32
            if pk.kind == L.OP_LG and d == 1 and pk.n_par <= 4:
                # params = W[Dp], bias, scale, 2 ln scale, var  (cpds.LinearGaussianCPD._pack)
                flags |= L.F_LGFAST
                pv = pk.params.astype(np.float32)
                emb = np.zeros(8, np.float32)
                emb[0:4] = pv[pk.n_par : pk.n_par + 4]
                emb[4 : 4 + pk.n_par] = pv[: pk.n_par]
                op["layer_dim"][:] = emb.view(np.int32)
                sl = list(par_slots[len(par_slots) - pk.n_par :] if pk.n_par else []) + [0] * (4 - pk.n_par)
                if max(sl + [slot_of[n]]) < 65536:
                    op["aux"][:] = [sl[0] | (sl[1] << 16), sl[2] | (sl[3] << 16), slot_of[n], 0]
                    lg_fast_ops.append(i)
                else:
                    flags &= ~L.F_LGFAST
            if r.add_logw:
                flags |= L.F_ADD_LOGW
                needs_logw = True
            if r.out_logp:
                flags |= L.F_OUT_LOGP
                needs_logp = True
            if r.src == "sample":
                if r.shared:
                    flags |= L.F_SHARED
                op["n_off"], op["u_off"] = stream_base[n]
                if r.inject:
                    op["noise_idx"] = len(noise)
                    noise.append(n)
            if pk.kind in (L.OP_GNN, L.OP_MDN) and 0 < pk.n_par <= 4 and pk.n_layers > 0:
                sl = list(par_slots[len(par_slots) - pk.n_par :]) + [0] * (4 - pk.n_par)
                if max(sl) < 65536:
                    flags |= L.F_PAR4
                    op["aux"][:] = [int(np.float32(cpds[n].min_scale).view(np.int32)),
                                    sl[0] | (sl[1] << 16), sl[2] | (sl[3] << 16), 0]
                    if (pk.kind == L.OP_MDN and int(op["tc"][0]) and d == 1 and 2 <= pk.k <= 5
                            and r.src == "sample" and not r.shared and not r.inject
                            and not r.add_logw and not r.out_logp and not r.out_params):
                        flags |= L.F_MDNPLAIN
                        if (pk.k == 3 and int(op["act"]) == L.ACT["relu"] and int(op["tc"][2]) == 0
                                and slot_of[n] < 65536):
                            flags |= L.F_MDNFAST
            if int(op["tc"][0]) and int(op["tc"][2]) == 0 and not (flags & L.F_PAR4):
                # the FP32-pipe first layer reads its parent slots from the descriptor (cpds.pack_mlp_tc l1_fma)
                raise ValueError(f"node '{n}': tensor-core image without first-layer MMA needs packed parent slots")
            if (_roots_enabled() and pk.kind == L.OP_MDN and pk.n_par == 0 and d == 1 and 2 <= pk.k <= 4
                    and r.src == "sample" and not r.shared and not r.inject and not r.store
                    and not r.add_logw and not r.out_logp and not r.out_params):
                # parent-less MDN that is only drawn: the mixture is row-independent, evaluate it here
                # (mdn.py:190-196 root parameters, 227-235 clamp / renormalise / softplus)
                cpd = cpds[n]
                pi = torch.softmax(cpd._logits.detach().float().cpu(), dim=0).clamp_min(1e-5)
                pi = pi / pi.sum().clamp_min(1e-12)
                cum = torch.cumsum(pi, dim=0)[: pk.k - 1]
                loc = cpd._loc.detach().float().cpu().reshape(pk.k)
                sc = F.softplus(cpd._log_scale.detach().float().cpu().reshape(pk.k)) + float(cpd.min_scale)
                emb = np.zeros(12, np.float32)
                emb[: pk.k - 1] = cum.numpy()
                emb[pk.k - 1 : pk.k - 1 + 2 * pk.k : 2] = loc.numpy()
                emb[pk.k : pk.k + 2 * pk.k : 2] = sc.numpy()
                op["layer_dim"][:] = emb[:8].view(np.int32)
                op["aux"][:] = emb[8:].view(np.int32)
                op["tc"][:] = [slot_of[n], int(op["n_off"]), int(op["u_off"]), pk.k]
                flags |= L.F_MDNROOT
            drawn_plain = (r.src == "sample" and not r.inject and not r.store and not r.add_logw and not r.out_logp)
            scored_plain = (r.src == "fixed_q" and r.add_logw and not r.store and not r.out_logp)
            if (_tabplain_enabled() and pk.kind == L.OP_TAB and d == 1 and pk.n_par <= 4 and not r.out_params
                    and (drawn_plain or scored_plain)):
                # (the tensor-core kernel has no TABPLAIN branch: there the op runs the generic lookup, which
                # reads none of the words rewritten here)
                emb = _tab_plain_fields(pk, par_slots[len(par_slots) - pk.n_par:] if pk.n_par else [],
                                        param_off[id(pk)], slot_of[n],
                                        int(op["u_off"]) if drawn_plain else int(op["fixed_col"]), scored_plain,
                                        trusted=[p in plain_drawn and dims[p] == 1 for p in plist])
                if emb is not None:
                    op["layer_dim"][:] = emb[:8]
                    op["aux"][:] = emb[8:]
                    flags |= L.F_TABPLAIN
                    if drawn_plain:
                        plain_drawn.add(n)  # its slot always holds a class index 0..k-1 produced by the kernel itself
            if r.out_params:
                flags |= L.F_OUT_PARAMS
                heavy = True  # the read-out lives in the HEAVY kernels only
            if flags & L.F_LGFAST:
                op["aux"][3] = op["n_off"]
                if (r.src == "sample" and not r.shared and not r.inject and not r.store
                        and not r.add_logw and not r.out_logp):
                    flags |= L.F_LGPLAIN
                    # op_lg_plain reads one word per parent slot: 0, 1 in aux[0..1], 2, 3 in layer_dim[2..3]
                    # (2 ln scale / var: only a scored op needs them)
                    a0, a1 = int(op["aux"][0]), int(op["aux"][1])
                    op["aux"][0], op["aux"][1] = a0 & 0xFFFF, (a0 >> 16) & 0xFFFF
                    op["layer_dim"][2], op["layer_dim"][3] = a1 & 0xFFFF, (a1 >> 16) & 0xFFFF
            op["flags"] = flags
            n_scratch = max(n_scratch, pk.scratch)
            heavy = heavy or pk.heavy
        for gone in expiring.get(i, ()):  # release slots whose last reader was this op
            s = slot_of[gone]
            for j in range(s, s + dims[gone]):
                occupied[j] = False

    if tc_ops:
        # Weight-ring images: [weights][tail], tail = verbatim copies of the descriptors of the ops that follow the
        # MLP op, up to and including the next MLP op.  The image is in shared memory long before the warps get
        # there, so the kernel reads those descriptors from the ring buffer instead of L2 (the 128-byte records do not
        # stay in the small L1 left beside ~180 KB of shared memory).  layer_dim[7] of the MLP op = records in its tail.
        tail_of = []
        for k, (i, w) in enumerate(tc_ops):
            nxt = tc_ops[k + 1][0] if k + 1 < len(tc_ops) else len(order) - 1
            room = (L.TC_WBUF_BYTES - 4 * int(w.size)) // 128
            # all or nothing: the kernel walks one running descriptor pointer, re-pointed once per MLP op
            n_tail = (nxt - i) if (_tails_enabled() and 0 < nxt - i <= room) else 0
            tail_of.append(n_tail)
            ops[i]["layer_dim"][7] = n_tail
            if int(ops[i]["flags"]) & L.F_MDNFAST:  # short descriptor fetch: out_slot and the tail count in tc[1]
                ops[i]["tc"][1] = int(ops[i]["out_slot"]) | (n_tail << 16)
        offs = []
        for (i, w), n_tail in zip(tc_ops, tail_of):  # offsets first: a tail may hold the NEXT MLP op's descriptor
            pad = (-blob_len) % 4  # the bulk copy needs a 16-byte aligned source
            blob_len += pad
            offs.append((blob_len, pad))
            if not int(ops[i]["flags"]) & L.F_MDNFAST:  # (informative: the kernel takes image offsets from tc_list)
                ops[i]["tc"][1] = blob_len
            blob_len += int(w.size) + 32 * n_tail
        for (i, w), n_tail, (off, pad) in zip(tc_ops, tail_of, offs):
            if pad:
                blob.append(np.zeros(pad, np.float32))
            blob.append(w)
            if n_tail:
                blob.append(np.frombuffer(ops[i + 1: i + 1 + n_tail].tobytes(), dtype=np.float32))
            tc_list.append((off, 4 * int(w.size) + 128 * n_tail))
    params = np.concatenate(blob) if blob else np.zeros(4, np.float32)
    if params.size == 0:
        params = np.zeros(4, np.float32)
    return Program(
        ops=ops,
        par_slots=np.asarray(par_slots if par_slots else [0], dtype=np.int32),
        params=params,
        n_slots=max(len(occupied), 1),
        n_scratch=int(n_scratch),
        heavy=bool(heavy),
        nodes=order,
        fixed_cols=fixed_cols,
        n_fixed_cols=n_fixed,
        inputs=inputs,
        stores=stores,
        noise=noise,
        needs_logw=needs_logw,
        needs_logp=needs_logp,
        dims=dims,
        store_widths={n: (int(cpds[n].param_width()) if roles[n].out_params else dims[n]) for n in stores},
        keep_slot=slot_of[keep_live] if keep_live is not None else -1,
        tc=any_tc,
        tc_list=np.asarray(tc_list, dtype=np.int32).reshape(-1, 2) if any_tc else None,
    )


# ------------------------------------------------------------------------------------------------
# Gibbs sampler (vbn/sampling/gibbs.py:23-92)
# ------------------------------------------------------------------------------------------------
@dataclass
class GibbsProgram(Program):
    """One launch runs the whole chain: ancestral initial state, then a VBN_OP_JUMP loop of sweeps.
    noise_keys[i] names the injected array set of noise[i]: ("init", node) | ("cand", node) | ("choice", node)."""

    noise_keys: List[tuple] = field(default_factory=list)
    n_candidates: int = 8


def compile_gibbs(topo: Sequence[str], parents: Dict[str, Sequence[str]], children: Dict[str, Sequence[str]],
                  cpds: Dict[str, BaseCPD], fixed: Sequence[str], target: str, total_steps: int,
                  n_candidates: int = 8, inject: bool = False) -> GibbsProgram:
    """Schedule of the reference's Gibbs sampler, one ROW per chain (= per query):

        init   : every node gets a permanent slot; fixed nodes are loaded from the per-query table, the others
                 drawn ancestrally (gibbs.py:32, _ancestral_sample_tensor with one sample: roots shared)
        sweep  : per latent node, per candidate c: the node's own CPD draws candidate c from the current parents
                 and adds its log-density; every child's CPD scores the child's CURRENT value (VBN_SRC_SLOT)
                 with the candidate substituted; VBN_OP_TAKEW moves the score to a slot.  VBN_OP_SELECT
                 then draws one candidate by softmax and overwrites the node's slot (gibbs.py:41-82)
        loop   : VBN_OP_JUMP repeats the sweep total_steps times with fresh Philox stream blocks
        output : the target's final state (the reference returns it n_samples times, see oracle gibbs_sample)
    """
    k = int(n_candidates)
    fixed = set(fixed)
    order = list(topo)
    dims = {n: int(cpds[n].output_dim) for n in order}
    slot_of, nxt = {}, 0
    for n in order:
        slot_of[n] = nxt
        nxt += dims[n]
    dmax = max(dims.values())
    s_score, s_cand, s_tmp = nxt, nxt + k, nxt + k + k * dmax
    n_slots = s_tmp + dmax

    blob: List[np.ndarray] = []
    blob_len = 0
    param_off: Dict[int, int] = {}
    par_slots: List[int] = []
    rows: List[dict] = []
    noise_keys: List[tuple] = []
    state = {"n_off": 0, "u_off": 0, "scratch": 0, "heavy": False, "n_fixed": 0}
    fixed_cols: Dict[str, int] = {}

    def params_of(pk: Packed) -> int:
        nonlocal blob_len
        key = id(pk)
        if key not in param_off:
            pad = (-blob_len) % 4
            if pad:
                blob.append(np.zeros(pad, np.float32))
                blob_len += pad
            param_off[key] = blob_len
            blob.append(pk.params.astype(np.float32, copy=False))
            blob_len += pk.params.size
        return param_off[key]

    def cpd_op(node: str, out_slot: int, par_of: Dict[str, int], flags: int, *, fixed_col: int = -1,
               noise_key=None, member: int = 0, group: int = 0, draws: bool = False) -> dict:
        pk = cpds[node].pack()
        plist = list(parents.get(node, ()))
        if sum(dims[p] for p in plist) != pk.n_par:
            raise ValueError(f"node '{node}': CPD expects {pk.n_par} parent dims")
        op = dict(kind=pk.kind, flags=flags, dim=dims[node], n_par=pk.n_par, out_slot=out_slot,
                  par_off=len(par_slots), param_off=params_of(pk), fixed_col=fixed_col, store_idx=-1, noise_idx=-1,
                  n_off=0, u_off=0, n_layers=pk.n_layers, act=pk.act, n_out=pk.n_out, k=pk.k,
                  layer_dim=list(pk.layer_dim)[: L.MAX_LAYERS], aux=list(pk.aux), tc=[0, group, member, 0])
        for p in plist:
            par_slots.extend(range(par_of[p], par_of[p] + dims[p]))
        if pk.n_layers == 3 and list(pk.layer_dim[:2]) == [32, 32]:
            op["flags"] |= L.F_FAST32
        if draws:
            op["n_off"], op["u_off"] = state["n_off"], state["u_off"]
            state["n_off"] += pk.n_normals
            state["u_off"] += pk.n_uniforms
            if inject and noise_key is not None:
                if noise_key not in noise_keys:
                    noise_keys.append(noise_key)
                op["noise_idx"] = noise_keys.index(noise_key)
        state["scratch"] = max(state["scratch"], pk.scratch)
        state["heavy"] = state["heavy"] or pk.heavy
        return op

    def glue(kind: int, **kw) -> dict:
        op = dict(kind=kind, flags=L.SRC_SAMPLE, dim=1, n_par=0, out_slot=0, par_off=0, param_off=0, fixed_col=-1,
                  store_idx=-1, noise_idx=-1, n_off=0, u_off=0, n_layers=0, act=0, n_out=0, k=0, layer_dim=[],
                  aux=[0, 0, 0, 0], tc=[0, 0, 0, 0])
        op.update(kw)
        return op

    # ---- initial state ----------------------------------------------------------------------------
    for n in order:
        if n in fixed:
            fixed_cols[n] = state["n_fixed"]
            rows.append(glue(L.OP_NONE, flags=L.SRC_FIXED_Q, dim=dims[n], out_slot=slot_of[n], fixed_col=state["n_fixed"]))
            state["n_fixed"] += dims[n]
        else:
            flags = L.SRC_SAMPLE | (L.F_SHARED if not parents.get(n) else 0)
            rows.append(cpd_op(n, slot_of[n], slot_of, flags, noise_key=("init", n), draws=True))
    # the loop's stream blocks start on a block boundary so that iteration t adds t * (blocks per sweep)
    state["n_off"] = (state["n_off"] + 3) & ~3
    state["u_off"] = (state["u_off"] + 3) & ~3
    n0, u0 = state["n_off"], state["u_off"]
    loop_start = len(rows)
    latent = [n for n in order if n not in fixed]
    for n in latent:
        d = dims[n]
        for c in range(k):
            cand = s_cand + c * d
            rows.append(cpd_op(n, cand, slot_of, L.SRC_SAMPLE | L.F_ADD_LOGW, noise_key=("cand", n), member=c,
                               group=k, draws=True))
            for ch in children.get(n, ()):
                par_of = dict(slot_of)
                par_of[n] = cand
                rows.append(cpd_op(ch, s_tmp, par_of, L.SRC_SLOT | L.F_ADD_LOGW, fixed_col=slot_of[ch]))
            rows.append(glue(L.OP_TAKEW, out_slot=s_score + c))
        sel = glue(L.OP_SELECT, dim=d, out_slot=slot_of[n], k=k, aux=[s_score, s_cand, 0, 0], u_off=state["u_off"])
        state["u_off"] += 1
        if inject:
            noise_keys.append(("choice", n))
            sel["noise_idx"] = len(noise_keys) - 1
        rows.append(sel)
    nq = (state["n_off"] - n0 + 3) // 4
    uq = (state["u_off"] - u0 + 3) // 4
    if latent and total_steps > 1:
        rows.append(glue(L.OP_JUMP, k=int(total_steps), aux=[loop_start, 0, 0, 0], layer_dim=[nq, uq]))
    rows.append(glue(L.OP_NONE, dim=dims[target], out_slot=slot_of[target], store_idx=0))

    ops = np.zeros(len(rows), dtype=L.OP_DTYPE)
    for i, r in enumerate(rows):
        for key, v in r.items():
            if key in ("layer_dim", "aux", "tc"):
                ops[i][key][: len(v)] = v
            else:
                ops[i][key] = v
    params = np.concatenate(blob) if blob else np.zeros(4, np.float32)
    return GibbsProgram(
        ops=ops, par_slots=np.asarray(par_slots if par_slots else [0], dtype=np.int32), params=params,
        n_slots=n_slots, n_scratch=int(state["scratch"]), heavy=True,  # the glue ops live in the HEAVY FFMA kernels only
        nodes=order,
        fixed_cols=fixed_cols, n_fixed_cols=state["n_fixed"], inputs=[], stores=[target],
        noise=[str(key) for key in noise_keys], needs_logw=False, needs_logp=False, dims=dims, tc=False, tc_list=None,
        store_widths={target: dims[target]}, noise_keys=noise_keys, n_candidates=k)
