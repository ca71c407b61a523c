"""CPU-only checks of the host side: the C-ABI library exports every symbol the header declares,
struct layouts match, the plan compiler's liveness/slot allocation, query validation, registry."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import vectorizedbayesiannetwork_b200 as V
from vectorizedbayesiannetwork_b200 import _lib as L
from vectorizedbayesiannetwork_b200 import synthetic as S
from vectorizedbayesiannetwork_b200.plan import Role, compile_schedule
from oracle.philox import KAT, philox4x32_10

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "vbn_cuda.h")).read()
    declared = set(re.findall(r"\b(vbn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = L.load()  # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vbn_cuda.h but not exported"
    assert declared == set(L.EXPORTS), (declared ^ set(L.EXPORTS))
    assert lib.vbn_cuda_abi_version() == L.ABI_VERSION


def test_stale_library_is_refused(monkeypatch):
    """A library built from another revision of the descriptor layout must not be used: load() compares
    vbn_cuda_abi_version() with the Python side's ABI_VERSION and raises (no silent mis-read of new descriptors)."""
    header = open(os.path.join(ROOT, "include", "vbn_cuda.h")).read()
    assert int(re.search(r"#define VBN_CUDA_ABI_VERSION (\d+)", header).group(1)) == L.ABI_VERSION
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "ABI_VERSION", L.ABI_VERSION + 1)
    with pytest.raises(L.VbnCudaError, match="ABI version mismatch"):
        L.load()


def test_descriptor_tails_are_all_or_nothing():
    """plan.py: an MLP op carries the descriptors of ALL ops up to and including the next MLP op in its weight image,
    or none (the tcgen05 kernel walks one running descriptor pointer, re-pointed once per MLP op)."""
    import random
    from vectorizedbayesiannetwork_b200 import synthetic as S
    from vectorizedbayesiannetwork_b200.plan import Role, compile_schedule
    import vectorizedbayesiannetwork_b200 as V
    rng, gen = random.Random(7), torch.Generator().manual_seed(7)
    names = [f"n{i}" for i in range(300)]
    parents, cpds = {}, {}
    mlp_at = (5, 9, 150, 290)
    for i, name in enumerate(names):
        ps = sorted(rng.sample(range(max(0, i - 20), i), min(rng.randint(1, 3), i))) if i else []
        parents[name] = [names[p] for p in ps]
        cpds[name] = S.mdn_cpd(gen, len(ps)) if i in mlp_at else S.lg_cpd(gen, len(ps))
    built = {n: V.cpd_from_spec(c, device="cpu") for n, c in cpds.items()}
    prog = compile_schedule(names, parents, built, {n: Role() for n in names}, use_tc=True)
    tc_idx = [i for i, op in enumerate(prog.ops) if op["tc"][0]]
    assert [prog.nodes[i] for i in tc_idx] == [f"n{i}" for i in mlp_at]
    tails = [int(prog.ops[i]["layer_dim"][7]) for i in tc_idx]
    assert tails == [4, 0, 0, 9], tails  # 9 -> 150 and 150 -> 290 do not fit a ring slot: no tail at all
    for (off, nbytes), i, t in zip(prog.tc_list, tc_idx, tails):
        assert nbytes % 128 == 0 or t == 0
        if t:  # the tail is a verbatim copy of the following descriptors
            tail = prog.params[off + nbytes // 4 - 32 * t: off + nbytes // 4]
            np.testing.assert_array_equal(tail.view(np.int32), prog.ops[i + 1: i + 1 + t].view(np.int32).ravel())


def test_struct_layouts_match_header():
    assert L.OP_DTYPE.itemsize == 128
    assert C.sizeof(L.ProgramDesc) == 88  # (has_tables fills the former padding)
    assert C.sizeof(L.RunDesc) == 136
    assert L.RunDesc.logw_dev.offset == 80 and L.RunDesc.error_flag_dev.offset == 104


def test_product_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(L.VbnCudaError):
        V.VBN.from_spec(S.lg_chain(3))
    cpd = V.cpd_from_spec(S.lg_chain(3)["cpds"]["x1"], device="cpu")
    with pytest.raises(L.VbnCudaError):
        cpd.sample(torch.zeros(2, 1), 4)


def _compile(spec, roles):
    cpds = {n: V.cpd_from_spec(c, device="cpu") for n, c in spec["cpds"].items()}
    return compile_schedule(spec["topo"], spec["parents"], cpds, roles)


def test_chain_needs_two_slots_regardless_of_length():
    spec = S.lg_chain(50)
    roles = {n: Role() for n in spec["topo"]}
    roles["x25"].store = True
    prog = _compile(spec, roles)
    assert len(prog.ops) == 50 and prog.n_slots == 2 and not prog.heavy
    assert prog.stores == ["x25"] and prog.n_fixed_cols == 0


def test_cfg5_live_set_is_small_and_slots_never_alias_live_values():
    spec = S.random_dag_lg_mdn(300, seed=0)
    roles = {n: Role() for n in spec["topo"]}
    for n in spec["topo"][-5:]:
        roles[n] = Role(src="fixed_q", add_logw=True)
    prog = _compile(spec, roles)
    assert prog.heavy and prog.needs_logw and prog.n_fixed_cols == 5
    assert prog.n_slots <= 40, prog.n_slots
    # replay the schedule symbolically: every parent slot must still hold that parent's value
    owner = {}
    slot_of = {n: int(op["out_slot"]) for n, op in zip(prog.nodes, prog.ops)}
    for n, op in zip(prog.nodes, prog.ops):
        if int(op["kind"]) != L.OP_NONE:
            got = list(prog.par_slots[int(op["par_off"]): int(op["par_off"]) + int(op["n_par"])])
            want = [slot_of[p] for p in spec["parents"][n]]
            assert got == want
            for p in spec["parents"][n]:
                assert owner[slot_of[p]] == p, f"{n}: parent {p} was overwritten"
        owner[int(op["out_slot"])] = n


def test_param_blocks_are_16_byte_aligned_and_shared():
    spec = S.random_dag_lg_mdn(40, seed=1)
    prog = _compile(spec, {n: Role() for n in spec["topo"]})
    assert all(int(op["param_off"]) % 4 == 0 for op in prog.ops)
    assert prog.params.dtype == np.float32


def test_mlp_packing_layout():
    g = torch.Generator().manual_seed(0)
    layers = S.mlp_layers(g, 3, (32, 32), 9)
    from vectorizedbayesiannetwork_b200.cpds import pack_mlp

    blob, dims = pack_mlp(layers, 3)
    assert dims == [32, 32, 9]
    w1, b1 = layers[0]
    np.testing.assert_array_equal(blob[: 3 * 32].reshape(3, 32), w1.numpy().T)
    np.testing.assert_array_equal(blob[96:128], b1.numpy())
    off = 128 + 32 * 32 + 32
    np.testing.assert_array_equal(blob[off: off + 9 * 32].reshape(9, 32), layers[2][0].numpy())
    assert blob.size == off + 9 * 32 + 12
    with pytest.raises(ValueError):
        pack_mlp(S.mlp_layers(g, 3, (200, 8), 2), 3)  # wider than the generic device path


def test_registry_and_resolution_mirror_reference():
    assert set(V.INFERENCE_REGISTRY) == {"likelihood_weighting", "importance_sampling", "monte_carlo_marginalization",
                                         "gaussian_exact", "categorical_exact", "resampled_importance_sampling", "rao_blackwellized_marginalization",
                                         "lbp"}
    ge = V.INFERENCE_REGISTRY["gaussian_exact"](n_samples=31, fallback="likelihood_weighting")
    assert ge.n_samples == 31 and ge._fallback.n_samples == 31 and ge.stddevs == 4.0
    with pytest.raises(ValueError):
        V.INFERENCE_REGISTRY["gaussian_exact"](fallback="gaussian_exact")
    with pytest.raises(ValueError):
        V.INFERENCE_REGISTRY["categorical_exact"](fallback="no_such_method")
    assert set(V.SAMPLING_REGISTRY) == {"ancestral", "gibbs"}
    from vectorizedbayesiannetwork_b200.core import register_inference

    with pytest.raises(ValueError):
        register_inference("importance_sampling")(object)  # duplicate key, registry.py:18-20
    inf = V.INFERENCE_REGISTRY["importance_sampling"](n_samples=12, unknown_kwarg=1)
    assert inf.n_samples == 12 and inf.ess_threshold == 0.1 and inf._last_fallback is False
    lw = V.INFERENCE_REGISTRY["likelihood_weighting"]()
    assert (lw.n_samples, lw.eps, lw.normalize) == (512, 1e-12, True)
    assert V.INFERENCE_REGISTRY["monte_carlo_marginalization"]().n_samples == 200
    assert V.SAMPLING_REGISTRY["ancestral"]().n_samples == 200


def test_philox_oracle_known_answers():
    for ctr, key, want in KAT:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), key[0], key[1])[0]
        assert [int(x) for x in got] == list(want)


def test_block_bounds_partition():
    from vectorizedbayesiannetwork_b200.dist import Shard, block_bounds

    for n, w in ((10, 3), (8, 8), (4096, 7), (5, 2)):
        parts = [block_bounds(n, r, w) for r in range(w)]
        assert sum(c for c, _ in parts) == n
        off = 0
        for c, o in parts:
            assert o == off
            off += c
    with pytest.raises(ValueError):
        Shard("queries", 3, 4).local_queries(2)
    assert Shard("samples", 1, 4).local_queries(2) == (2, 0)
    assert Shard("samples", 1, 4).local_samples(10) == (3, 3)
