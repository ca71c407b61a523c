"""Generates tests/golden/*.pt from the UNMODIFIED reference (needs /root/reference; run in the
build container):  python tests/golden/make_golden.py

Each case stores: the fitted model as a spec dict (parameters read from the reference objects),
the query, the draws the reference consumed (recorded through the oracle, which is bit-identical
to the reference under the same seed -- asserted here per case) and the REFERENCE outputs.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import refmodels  # noqa: E402
from noise_util import log_to_strkeys  # noqa: E402
from oracle import vbn_oracle as O  # noqa: E402

METHODS = {
    "lw": ("likelihood_weighting", O.likelihood_weighting),
    "is": ("importance_sampling", O.importance_sampling),
    "mcm": ("monte_carlo_marginalization", O.monte_carlo_marginalization),
    "ris": ("resampled_importance_sampling", O.resampled_importance_sampling),
    "rb": ("rao_blackwellized_marginalization", O.rao_blackwellized_marginalization),
    "gexact": ("gaussian_exact", O.gaussian_exact),
    "cexact": ("categorical_exact", O.categorical_exact),
}


def _same(a, b):
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=7.0), torch.nan_to_num(b, nan=7.0))


def run_case(model, spec, query, S, method, seed):
    q = {"target": query["target"], "evidence": query.get("evidence", {}), "do": query.get("do", {})}
    rec = O.RecordingNoise()
    info = {}
    if method == "anc":
        model.set_sampling_method("ancestral")
        torch.manual_seed(seed)
        ref_s = model.sample(q, n_samples=S)
        torch.manual_seed(seed)
        ora_s = O.ancestral_sample(spec, q, S, noise=rec)
        assert _same(ref_s, ora_s)
        expect = {"samples": ref_s}
    else:
        name, fn = METHODS[method]
        model.set_inference_method(name, n_samples=S)
        torch.manual_seed(seed)
        ref_w, ref_s = model.infer_posterior(q)
        torch.manual_seed(seed)
        if method == "ris":
            ora_w, ora_s, inf = fn(spec, q, S, noise=rec, return_info=True)
            info = {"resampled": bool(inf["resampled"])}
            assert info["resampled"] == model._inference._last_resampled
        elif method == "is":
            ora_w, ora_s, inf = fn(spec, q, S, noise=rec, return_info=True)
            info = {"fallback": bool(inf["fallback"]), "ess": inf["ess"]}
            assert info["fallback"] == model._inference._last_fallback
        else:
            ora_w, ora_s = fn(spec, q, S, noise=rec)
        assert _same(ref_w, ora_w) and _same(ref_s, ora_s), (method, query["target"])
        expect = {"pdf": ref_w, "samples": ref_s}
    return {"query": q, "S": S, "method": method, "noise": log_to_strkeys(rec.log), "expect": expect, "info": info}


def cpd_cases(model, spec, S=9, seed=3):
    out = []
    for node, cpd in model.nodes.items():
        c = spec["cpds"][node]
        dp = c["input_dim"]
        if c["kind"] in ("categorical_table", "categorical_embedded_softmax") and dp:
            pick = lambda *shape: torch.stack([v[torch.randint(0, v.numel(), shape)] for v in c["parent_values"]], dim=-1)
            variants = [pick(4), pick(4, S)]
        else:
            variants = [None] if dp == 0 else [torch.randn(4, dp), torch.randn(4, S, dp)]
        for parents in variants:
            rec = O.RecordingNoise()
            torch.manual_seed(seed)
            ref = cpd.sample(parents, S).detach()
            torch.manual_seed(seed)
            ora = O.cpd_sample(c, parents, S, noise=rec, key=("cpd", node))
            assert _same(ref, ora)
            lp = cpd.log_prob(ref, parents).detach()
            lp2 = cpd.log_prob(ref[:, 0], parents if parents is None or parents.dim() == 2 else parents[:, :1]).detach()
            out.append({"node": node, "parents": parents, "S": S, "noise": log_to_strkeys(rec.log),
                        "samples": ref, "log_prob": lp, "log_prob_2d": lp2})
    return out


def table_files():
    """categorical_table + discrete softmax_nn model; generated on its own so the
    earlier fixtures stay byte-identical:  python tests/golden/make_golden.py table"""
    torch.manual_seed(4321)
    m = refmodels.table_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "rain", "evidence": {"slip": torch.tensor([[1.0], [0.0]]), "wet": torch.tensor([[2.0], [1.0]])}},
          {"target": "slip", "evidence": {"season": torch.tensor([[3.0], [0.0], [1.0]])}},
          {"target": "season", "evidence": {"slip": torch.tensor([[1.0]]), "sprinkler": torch.tensor([[0.0]])}},
          {"target": "wet", "evidence": {}, "do": {"rain": torch.tensor([[1.0], [0.0]])}}]
    cases = [run_case(m, spec, q, 64, meth, 17) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    return {"table": {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}}


def rff_files():
    """rff_gaussian nodes:  python tests/golden/make_golden.py rff"""
    torch.manual_seed(8642)
    m = refmodels.rff_model()
    spec = O.spec_from_reference(m)
    ev = torch.tensor([[0.4], [-1.1], [2.0]])
    qs = [{"target": "b", "evidence": {"d": ev}},
          {"target": "c", "evidence": {"a": ev, "e": -ev}},
          {"target": "a", "evidence": {"c": torch.randn(2, 2)}},
          {"target": "d", "evidence": {}, "do": {"b": ev}}]
    cases = [run_case(m, spec, q, 40, meth, 23) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    cases += [run_case(m, spec, {"target": "b", "evidence": {"a": ev, "e": 0.5 * ev}}, 21, meth, 5) for meth in ("gexact", "rb")]
    cases += [run_case(m, spec, {"target": "d", "evidence": {"a": ev}}, 21, meth, 5) for meth in ("gexact", "rb")]
    return {"rff": {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}}


def embedded_files():
    """categorical_embedded_softmax nodes:  python tests/golden/make_golden.py embedded"""
    torch.manual_seed(7531)
    m = refmodels.embedded_model()
    spec = O.spec_from_reference(m)
    t = lambda *v: torch.tensor([[float(x)] for x in v])
    qs = [{"target": "v", "evidence": {"z": t(1, 0)}},
          {"target": "u", "evidence": {"y": t(2, 0, 1)}},
          {"target": "y", "evidence": {"t": t(1)}, "do": {"u": t(2)}},
          {"target": "z", "evidence": {"u": t(0, 1, 2), "t": t(1, 1, 0)}}]
    cases = [run_case(m, spec, q, 64, meth, 29) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    for meth in ("cexact", "rb"):
        cases.append(run_case(m, spec, {"target": "v", "evidence": {"u": t(0, 2), "t": t(1, 0)}}, 16, meth, 5))
        cases.append(run_case(m, spec, {"target": "u", "evidence": {}}, 16, meth, 5))
        cases.append(run_case(m, spec, {"target": "y", "evidence": {"t": t(1, 0)}}, 16, meth, 5))
    return {"embedded": {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}}


def gibbs_case(model, spec, query, n, seed, **kw):
    q = {"target": query["target"], "evidence": query.get("evidence", {}), "do": query.get("do", {})}
    model.set_sampling_method("gibbs", n_samples=n, **kw)
    torch.manual_seed(seed)
    ref = model.sample(q, n_samples=n).detach()
    rec = O.RecordingNoise()
    torch.manual_seed(seed)
    ora = O.gibbs_sample(spec, q, n, noise=rec, **kw)
    assert _same(ref, ora)
    return {"query": q, "n": n, "kw": kw, "noise": log_to_strkeys(rec.log), "expect": ref}


def gibbs_file():
    """Gibbs sampler: chains recorded from the reference, per model kind:
    python tests/golden/make_golden.py gibbs"""
    torch.manual_seed(1593)
    out = []
    m = refmodels.readme_model(n=300, epochs=2)
    spec = O.spec_from_reference(m)
    cases = [gibbs_case(m, spec, {"target": "feature_2", "evidence": {"feature_0": torch.tensor([[0.3]]), "feature_1": torch.tensor([[-0.2]])}}, 12, 3),
             gibbs_case(m, spec, {"target": "feature_0", "evidence": {"feature_2": torch.tensor([[0.1]])}}, 9, 4, burn_in=3, n_steps=2)]
    out.append({"name": "readme", "spec": spec, "cases": cases})
    m = refmodels.lg_chain_model(n_nodes=5)
    spec = O.spec_from_reference(m)
    cases = [gibbs_case(m, spec, {"target": "x2", "evidence": {"x4": torch.tensor([[0.7]])}}, 10, 5, burn_in=4),
             gibbs_case(m, spec, {"target": "x3", "evidence": {"x0": torch.tensor([[0.2], [1.0], [-0.7]])}}, 6, 6, burn_in=2)]
    out.append({"name": "lg_chain", "spec": spec, "cases": cases})
    m = refmodels.mixed_model(rows=256, epochs=1)
    spec = O.spec_from_reference(m)
    out.append({"name": "mixed", "spec": spec,
                "cases": [gibbs_case(m, spec, {"target": "e", "evidence": {"g": torch.tensor([[0.3, -0.2]])}}, 5, 7, burn_in=2)]})
    m = refmodels.table_model(rows=300, epochs=1)
    spec = O.spec_from_reference(m)
    out.append({"name": "table", "spec": spec,
                "cases": [gibbs_case(m, spec, {"target": "rain", "evidence": {"slip": torch.tensor([[1.0]])}}, 7, 8, burn_in=2)]})
    path = os.path.join(HERE, "gibbs.pt")
    torch.save({"models": out}, path)
    print(f"gibbs: {sum(len(m['cases']) for m in out)} cases, {os.path.getsize(path)/1024:.0f} KiB")


def exact_files():
    """gaussian_exact / categorical_exact, incl. their likelihood-weighting fallbacks:
    python tests/golden/make_golden.py exact"""
    torch.manual_seed(9876)
    out = {}
    m = refmodels.lg_chain_model(n_nodes=5)
    spec = O.spec_from_reference(m)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    qs = [{"target": "x3", "evidence": {"x2": ev}}, {"target": "x0", "evidence": {"x4": ev}},
          {"target": "x1", "evidence": {"x4": ev}}, {"target": "x2", "evidence": {"x2": ev}},
          {"target": "x4", "evidence": {}, "do": {"x3": ev}}]
    out["exact_lg"] = {"spec": spec, "cases": [run_case(m, spec, q, 33, "gexact", 21) for q in qs], "cpd_cases": []}
    m = refmodels.readme_model(n=300, epochs=2)
    spec = O.spec_from_reference(m)
    qs = [{"target": "feature_0", "evidence": {"feature_2": ev}},
          {"target": "feature_2", "evidence": {"feature_0": ev, "feature_1": ev}}]
    out["exact_gnn"] = {"spec": spec, "cases": [run_case(m, spec, q, 24, "gexact", 22) for q in qs], "cpd_cases": []}
    m = refmodels.table_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "wet", "evidence": {"rain": torch.tensor([[1.0], [0.0]]), "sprinkler": torch.tensor([[0.0], [1.0]])}},
          {"target": "slip", "evidence": {"wet": torch.tensor([[2.0], [0.0], [1.0]])}},
          {"target": "season", "evidence": {"slip": torch.tensor([[1.0]])}},
          {"target": "sprinkler", "evidence": {"slip": torch.tensor([[1.0], [0.0]])}},
          {"target": "rain", "evidence": {"slip": torch.tensor([[1.0]])}},
          {"target": "wet", "evidence": {"wet": torch.tensor([[2.0]])}}]
    out["exact_cat"] = {"spec": spec, "cases": [run_case(m, spec, q, 40, "cexact", 23) for q in qs], "cpd_cases": []}
    return out


def rb_files():
    """rao_blackwellized_marginalization:  python tests/golden/make_golden.py rb"""
    torch.manual_seed(1357)
    out = {}
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    m = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(m)
    qs = [{"target": "x5", "evidence": {"x1": ev}}, {"target": "x0", "evidence": {}},
          {"target": "x2", "evidence": {"x4": ev}}, {"target": "x3", "evidence": {"x3": ev}},
          {"target": "x4", "evidence": {"x1": ev}, "do": {"x2": ev}}]
    out["rb_lg"] = {"spec": spec, "cases": [run_case(m, spec, q, 40, "rb", 41) for q in qs], "cpd_cases": []}
    m = refmodels.table_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "slip", "evidence": {"season": torch.tensor([[3.0], [0.0], [1.0]])}},
          {"target": "wet", "evidence": {"season": torch.tensor([[1.0], [2.0]])}},
          {"target": "season", "evidence": {}}, {"target": "sprinkler", "evidence": {"season": torch.tensor([[1.0]])}}]
    out["rb_table"] = {"spec": spec, "cases": [run_case(m, spec, q, 48, "rb", 42) for q in qs], "cpd_cases": []}
    m = refmodels.mixed_model(epochs=1)
    spec = O.spec_from_reference(m)
    qs = [{"target": "e", "evidence": {"a": torch.randn(3, 1)}},      # gaussian_nn target, mdn/lg parents sampled
          {"target": "c", "evidence": {"a": torch.randn(3, 1)}},      # mdn target -> fallback
          {"target": "g", "evidence": {"a": torch.randn(3, 1)}}]      # 2-D target -> fallback
    out["rb_mixed"] = {"spec": spec, "cases": [run_case(m, spec, q, 24, "rb", 43) for q in qs], "cpd_cases": []}
    return out


def ris_files():
    """resampled_importance_sampling:  python tests/golden/make_golden.py ris"""
    torch.manual_seed(2468)
    out = {}
    m = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(m)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    qs = [{"target": "x2", "evidence": {"x3": ev, "x5": ev * 2}}, {"target": "x5", "evidence": {"x1": ev}},
          {"target": "x0", "evidence": {"x4": ev}, "do": {"x2": ev}}, {"target": "x3", "evidence": {}},
          {"target": "x4", "evidence": {"x4": ev, "x1": ev}}]
    out["ris_lg"] = {"spec": spec, "cases": [run_case(m, spec, q, 48, "ris", 31) for q in qs], "cpd_cases": []}
    m = refmodels.mixed_model(epochs=1)
    spec = O.spec_from_reference(m)
    qs = [{"target": "e", "evidence": {"g": torch.randn(3, 2), "h": torch.randn(3, 1)}},
          {"target": "a", "evidence": {"f": torch.randn(3, 1) ** 2, "c": torch.randn(3, 1)}, "do": {"d": torch.randn(3, 1)}}]
    out["ris_mixed"] = {"spec": spec, "cases": [run_case(m, spec, q, 32, "ris", 32) for q in qs], "cpd_cases": []}
    m = refmodels.table_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "season", "evidence": {"slip": torch.tensor([[1.0], [0.0]]), "wet": torch.tensor([[2.0], [1.0]])}}]
    out["ris_table"] = {"spec": spec, "cases": [run_case(m, spec, q, 64, "ris", 33) for q in qs], "cpd_cases": []}
    return out


def summaries_file():
    """Posterior summaries of the benchmark adapter: inputs and the outputs of the
    reference's own functions (benchmarking/models/vbn.py:202-242, 381-423):
    python tests/golden/make_golden.py summaries"""
    refmodels.import_reference()
    from benchmarking.models import vbn as RB

    g = torch.Generator().manual_seed(97531)
    cases = []
    for (b, s, k) in ((1, 7, 2), (3, 200, 4), (5, 1000, 3), (2, 513, 20), (4, 64, 37)):
        x = torch.randint(-1, k + 1, (b, s), generator=g).float()  # some values outside the class set
        x[:, ::11] += 0.5  # exact halves: Python round() goes to even
        x[:, 1::13] += 0.25 * torch.randn(x[:, 1::13].shape, generator=g)
        w = torch.rand(b, s, generator=g)
        w = w / w.sum(1, keepdim=True)
        if s > 20:
            w[0, 3] = float("nan")
            w[0, 5] = float("inf")
        if b > 1:
            w[1] = 0.0  # empty histogram -> uniform
        probs = RB._estimate_discrete_posterior_batch(x.unsqueeze(-1), w, k)
        assert probs == O.estimate_discrete_posterior_batch(x.unsqueeze(-1), w, k)
        cases.append({"kind": "discrete", "samples": x.unsqueeze(-1), "weights": w, "k": k, "probs": torch.tensor(probs, dtype=torch.float64)})
    model = RB.VBNBenchmarkModel.__new__(RB.VBNBenchmarkModel)
    for (s, mode) in ((9, "w"), (400, "w"), (4096, "w"), (300, "none"), (300, "zero"), (50, "neg")):
        x = 2.0 * torch.randn(1, s, 1, generator=g) + 0.7
        w = torch.rand(1, s, generator=g)
        if mode == "zero":
            w = torch.zeros(1, s)
        if mode == "neg":
            w = w - 0.3  # negative weights are clipped at 0 (:396)
        arg = None if mode == "none" else w
        out = model._continuous_from_samples(x, weights=arg)
        mine = O.continuous_from_samples(x, weights=arg)
        assert out == mine, (out["mean"], mine["mean"])
        cases.append({"kind": "continuous", "samples": x, "weights": arg, "out": out})
    path = os.path.join(HERE, "summaries.pt")
    torch.save({"cases": cases}, path)
    print(f"summaries: {len(cases)} cases, {os.path.getsize(path)/1024:.0f} KiB")


def save(files):
    total = 0
    for name, blob in files.items():
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(blob, path)
        total += os.path.getsize(path)
        print(f"{name}: {len(blob['cases'])} cases, {len(blob['cpd_cases'])} cpd cases, {os.path.getsize(path)/1024:.0f} KiB")
    print(f"total {total/1024:.0f} KiB")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "table":
        save(table_files())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rb":
        save(rb_files())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ris":
        save(ris_files())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gibbs":
        gibbs_file()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "embedded":
        save(embedded_files())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rff":
        save(rff_files())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "summaries":
        summaries_file()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "exact":
        save(exact_files())
        return
    torch.manual_seed(1234)
    files = {}

    m = refmodels.readme_model(epochs=20)
    spec = O.spec_from_reference(m)
    q_readme = {"target": "feature_2", "evidence": {"feature_0": torch.tensor([[0.3]]), "feature_1": torch.tensor([[-0.2]])}}
    cases = [run_case(m, spec, q_readme, 200, "mcm", 0)]
    q2 = {"target": "feature_2", "evidence": {"feature_0": torch.randn(5, 1)}}
    q3 = {"target": "feature_0", "evidence": {"feature_2": torch.randn(3, 1)}}
    for q in (q_readme, q2, q3):
        for meth in ("lw", "is", "mcm", "anc"):
            cases.append(run_case(m, spec, q, 48, meth, 5))
    files["readme"] = {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}

    m = refmodels.lg_chain_model(n_nodes=6)
    spec = O.spec_from_reference(m)
    ev = torch.tensor([[0.2], [1.0], [-0.7]])
    qs = [{"target": "x2", "evidence": {"x5": ev}},
          {"target": "x0", "evidence": {"x5": ev}},
          {"target": "x4", "evidence": {"x5": ev}, "do": {"x1": torch.tensor([[0.5], [0.5], [-1.0]])}},
          {"target": "x3", "evidence": {}},
          {"target": "x3", "evidence": {"x2": ev, "x5": torch.tensor([[float("nan")], [float("inf")], [1.0]])}}]
    cases = [run_case(m, spec, q, 64, meth, 7) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    files["lg_chain"] = {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}

    for act in ("relu", "tanh", "gelu", "elu"):
        m = refmodels.mixed_model(activation=act, epochs=2)
        spec = O.spec_from_reference(m)
        B = 3
        qs = [{"target": "e", "evidence": {"g": torch.randn(B, 2), "h": torch.randn(B, 1)}},
              {"target": "a", "evidence": {"f": torch.randn(B, 1) ** 2, "c": torch.randn(B, 1)}, "do": {"d": torch.randn(B, 1)}},
              {"target": "h", "evidence": {}},
              {"target": "g", "evidence": {"e": torch.randn(B, 1), "f": torch.rand(B, 1)}}]
        meths = ("lw", "is", "mcm", "anc") if act == "relu" else ("lw",)
        cases = [run_case(m, spec, q, 32, meth, 9) for q in qs for meth in meths]
        files[f"mixed_{act}"] = {"spec": spec, "cases": cases,
                                 "cpd_cases": cpd_cases(m, spec) if act in ("relu", "gelu") else []}

    m = refmodels.discrete_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "rain", "evidence": {"slip": torch.tensor([[1.0], [0.0]]), "wet": torch.tensor([[2.0], [1.0]])}},
          {"target": "slip", "evidence": {"rain": torch.tensor([[1.0], [0.0], [1.0]])}},
          {"target": "wet", "evidence": {"slip": torch.tensor([[1.0]])}}]
    cases = [run_case(m, spec, q, 64, meth, 11) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    files["discrete"] = {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}

    for wb, clip in (("uniform", False), ("triangular", False), ("gaussian", False), ("uniform", True), ("triangular", True)):
        m = refmodels.binned_model(within_bin=wb, clip=clip)
        spec = O.spec_from_reference(m)
        qs = [{"target": "p", "evidence": {"q": torch.randn(4, 2) * 0.5}},
              {"target": "q", "evidence": {"p": torch.randn(4, 2) * 3.0}}]
        cases = [run_case(m, spec, q, 32, meth, 13) for q in qs for meth in ("lw", "is", "mcm")]
        files[f"binned_{wb}{'_clip' if clip else ''}"] = {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}

    m = refmodels.kde_model()
    spec = O.spec_from_reference(m)
    qs = [{"target": "p", "evidence": {"y": torch.randn(3, 1)}},
          {"target": "y", "evidence": {"p2": torch.randn(3, 1)}},
          {"target": "y", "evidence": {"p": torch.randn(2, 1), "p2": torch.randn(2, 1)}}]
    cases = [run_case(m, spec, q, 40, meth, 15) for q in qs for meth in ("lw", "is", "mcm", "anc")]
    files["kde"] = {"spec": spec, "cases": cases, "cpd_cases": cpd_cases(m, spec)}

    files.update(table_files())
    files.update(exact_files())
    files.update(ris_files())
    files.update(rb_files())
    save(files)


if __name__ == "__main__":
    main()
