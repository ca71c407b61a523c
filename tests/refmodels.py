"""Builders of small fitted *reference* models (only usable where /root/reference exists,
i.e. the build container).  Used by the oracle pin tests and by tests/golden/make_golden.py.
Nothing here runs on the GPU box."""
from __future__ import annotations

import os
import sys

import torch

REF_ROOT = os.environ.get("VBN_REFERENCE_ROOT", "/root/reference")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vbn"))


def import_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import vbn  # noqa: F401

    return vbn


def _fit(g, nodes_cpds, data, seed=0):
    vbn = import_reference()
    model = vbn.VBN(g, seed=seed, device="cpu")
    model.set_learning_method("node_wise", nodes_cpds=nodes_cpds)
    model.fit(data, verbosity=0)
    return model


def readme_model(n=1000, epochs=20):
    """README minimal example (README.md:88-127): gaussian_nn x2 -> mdn(K=3)."""
    import networkx as nx

    gen = torch.Generator().manual_seed(0)
    x0 = torch.randn(n, generator=gen)
    x1 = torch.randn(n, generator=gen)
    x2 = 0.5 * x0 - 0.2 * x1 + 0.1 * torch.randn(n, generator=gen)
    g = nx.DiGraph()
    g.add_edges_from([("feature_0", "feature_2"), ("feature_1", "feature_2")])
    fit = {"epochs": epochs, "batch_size": 256}
    cpds = {
        "feature_0": {"cpd": "gaussian_nn", "fit": fit},
        "feature_1": {"cpd": "gaussian_nn", "fit": fit},
        "feature_2": {"cpd": "mdn", "n_components": 3, "fit": fit},
    }
    data = {"feature_0": x0[:, None], "feature_1": x1[:, None], "feature_2": x2[:, None]}
    return _fit(g, cpds, data)


def lg_chain_model(n_nodes=6, rows=2048, slope=1.0, seed=0):
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    names = [f"x{i}" for i in range(n_nodes)]
    g.add_nodes_from(names)
    for a, b in zip(names[:-1], names[1:]):
        g.add_edge(a, b)
    data = {}
    x = torch.randn(rows, generator=gen)
    data[names[0]] = x[:, None]
    for nm in names[1:]:
        x = slope * x + 0.1 + 0.5 * torch.randn(rows, generator=gen)
        data[nm] = x[:, None]
    cpds = {nm: {"cpd": "linear_gaussian"} for nm in names}
    return _fit(g, cpds, data)


def mixed_model(rows=512, epochs=3, seed=0, activation="relu", dims=None):
    """Diamond-ish DAG mixing every CPD kind on the path, incl. a 2-D node.

        a(lg root) -> c(mdn) -> e(gaussian_nn) -> g(lg, D=2)
        b(gnn root) -> c ; b -> d(lg) ; d -> e ; c -> f(softmax_nn binned) ; f -> g
        r(mdn root) -> h(kde) ; a -> h
    """
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from(
        [("a", "c"), ("b", "c"), ("b", "d"), ("c", "e"), ("d", "e"), ("c", "f"), ("e", "g"),
         ("f", "g"), ("r", "h"), ("a", "h")]
    )
    rn = lambda *s: torch.randn(*s, generator=gen)
    a = rn(rows, 1)
    b = 0.5 + 1.5 * rn(rows, 1)
    r = torch.where(torch.rand(rows, 1, generator=gen) < 0.4, -2 + 0.3 * rn(rows, 1), 1 + 0.5 * rn(rows, 1))
    c = torch.tanh(a - 0.5 * b) + 0.3 * rn(rows, 1)
    d = 0.8 * b - 0.2 + 0.4 * rn(rows, 1)
    e = torch.sin(c) + 0.5 * d + 0.2 * rn(rows, 1)
    f = c**2 + 0.3 * rn(rows, 1)
    gg = torch.cat([e + 0.5 * f, e - f], dim=1) + 0.3 * rn(rows, 2)
    h = torch.sin(r) + 0.3 * a + 0.2 * rn(rows, 1)
    fit = {"epochs": epochs, "batch_size": 128}
    cpds = {
        "a": {"cpd": "linear_gaussian"},
        "b": {"cpd": "gaussian_nn", "fit": fit},
        "r": {"cpd": "mdn", "n_components": 2, "fit": fit},
        "c": {"cpd": "mdn", "n_components": 3, "activation": activation, "fit": fit},
        "d": {"cpd": "linear_gaussian"},
        "e": {"cpd": "gaussian_nn", "activation": activation, "fit": fit},
        "f": {"cpd": "softmax_nn", "n_classes": 5, "binning": "quantile",
              "within_bin": "triangular", "activation": activation, "fit": fit},
        "g": {"cpd": "linear_gaussian"},
        "h": {"cpd": "kde", "bandwidth": 0.4, "parent_bandwidth": 0.6, "max_points": 64},
    }
    data = {"a": a, "b": b, "r": r, "c": c, "d": d, "e": e, "f": f, "g": gg, "h": h}
    return _fit(g, cpds, data)


def discrete_model(rows=600, epochs=5, seed=0):
    """Small discrete BN with softmax_nn in discrete mode (like examples/06 and cfg3):
    rain(2) -> wet(3) <- sprinkler(2);  wet -> slip(2)."""
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from([("rain", "wet"), ("sprinkler", "wet"), ("wet", "slip")])
    u = lambda: torch.rand(rows, generator=gen)
    rain = (u() < 0.3).float()
    spr = (u() < 0.5).float()
    wet = torch.clamp(rain + spr + (u() < 0.2).float() - (u() < 0.2).float(), 0, 2)
    slip = ((wet > 0) & (u() < 0.6)).float()
    fit = {"epochs": epochs, "batch_size": 128}
    cpds = {
        "rain": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
        "sprinkler": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
        "wet": {"cpd": "softmax_nn", "n_classes": 3, "fit": fit},
        "slip": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
    }
    data = {"rain": rain[:, None], "sprinkler": spr[:, None], "wet": wet[:, None], "slip": slip[:, None]}
    return _fit(g, cpds, data)


def table_model(rows=800, epochs=4, seed=0):
    """Discrete BN mixing categorical_table and softmax_nn (discrete mode) nodes, incl. a 4-state node:
    rain(ct,2) -> wet(ct,3) <- sprinkler(snn,2);  wet -> slip(snn,2);  season(ct,4) -> rain."""
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from([("season", "rain"), ("rain", "wet"), ("sprinkler", "wet"), ("wet", "slip")])
    u = lambda: torch.rand(rows, generator=gen)
    season = torch.randint(0, 4, (rows,), generator=gen).float()
    rain = (u() < 0.15 + 0.15 * season).float()
    spr = (u() < 0.5).float()
    wet = torch.clamp(rain + spr + (u() < 0.2).float() - (u() < 0.2).float(), 0, 2)
    slip = ((wet > 0) & (u() < 0.6)).float()
    fit = {"epochs": epochs, "batch_size": 128}
    cpds = {
        "season": {"cpd": "categorical_table"},
        "rain": {"cpd": "categorical_table", "alpha": 0.5},
        "sprinkler": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
        "wet": {"cpd": "categorical_table", "alpha": 1.0, "alpha_mode": "total_mass", "prior": "global"},
        "slip": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
    }
    data = {"season": season[:, None], "rain": rain[:, None], "sprinkler": spr[:, None], "wet": wet[:, None],
            "slip": slip[:, None]}
    return _fit(g, cpds, data)


def rff_model(rows=600, seed=0):
    """rff_gaussian nodes (vbn/cpds/rff_gaussian.py) next to linear_gaussian ones, incl. a root and a 2-D node:
    a(rff root) -> b(rff, 24 features) ; a, b -> c(rff, D=2, default 256 features) ; c -> d(lg) ; e(lg root) -> b."""
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from([("a", "b"), ("e", "b"), ("a", "c"), ("b", "c"), ("c", "d")])
    rn = lambda *s: torch.randn(*s, generator=gen)
    a = 0.3 + 1.2 * rn(rows, 1)
    e = rn(rows, 1)
    b = torch.sin(1.5 * a) - 0.4 * e + 0.2 * rn(rows, 1)
    c = torch.cat([torch.cos(a) * b, a - b**2], dim=1) + 0.25 * rn(rows, 2)
    d = c.sum(1, keepdim=True) * 0.5 + 0.3 * rn(rows, 1)
    cpds = {
        "a": {"cpd": "rff_gaussian"},
        "e": {"cpd": "linear_gaussian"},
        # ridge keeps sum|coef| ~ |loc|: with the default 1e-6 the fit is ill-conditioned (sum|coef| ~ 400 for an
        # O(1) loc) and fp32 summation ORDER alone moves loc by 1e-5 relative -- in the reference as much as here
        "b": {"cpd": "rff_gaussian", "n_features": 24, "lengthscale": 0.8, "ridge": 3e-2},
        "c": {"cpd": "rff_gaussian", "ridge": 1e-1},
        "d": {"cpd": "linear_gaussian"},
    }
    return _fit(g, cpds, {"a": a, "e": e, "b": b, "c": c, "d": d})


def embedded_model(rows=700, epochs=4, seed=0):
    """categorical_embedded_softmax nodes (vbn/cpds/categorical_embedded_softmax.py) in a discrete BN with a
    categorical_table root and a discrete softmax_nn node:
    u(ces root,3) -> v(ces,4) <- t(ct,2) ;  v -> y(ces, n_classes given) ; u -> y ; y -> z(snn,2)."""
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from([("u", "v"), ("t", "v"), ("v", "y"), ("u", "y"), ("y", "z")])
    ri = lambda k: torch.randint(0, k, (rows,), generator=gen)
    u = ri(3)
    t = ri(2)
    v = (u + 2 * t + (torch.rand(rows, generator=gen) < 0.25).long()) % 4
    y = ((v + u) % 3 + (torch.rand(rows, generator=gen) < 0.2).long()) % 3
    z = ((y > 0) & (torch.rand(rows, generator=gen) < 0.7)).long()
    fit = {"epochs": epochs, "batch_size": 128}
    cpds = {
        "u": {"cpd": "categorical_embedded_softmax", "fit": fit},
        "t": {"cpd": "categorical_table"},
        "v": {"cpd": "categorical_embedded_softmax", "embedding_dim": 4, "hidden_dims": [16, 16], "fit": fit},
        "y": {"cpd": "categorical_embedded_softmax", "n_classes": 3, "activation": "tanh", "fit": fit},
        "z": {"cpd": "softmax_nn", "n_classes": 2, "fit": fit},
    }
    data = {k: x.float()[:, None] for k, x in {"u": u, "t": t, "v": v, "y": y, "z": z}.items()}
    return _fit(g, cpds, data)


def binned_model(within_bin="uniform", clip=False, rows=400, epochs=3, seed=0, dim=2):
    """softmax_nn in binned-continuous mode, root + child, D=dim."""
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edge("p", "q")
    p = torch.randn(rows, dim, generator=gen)
    q = torch.tanh(p) + 0.3 * torch.randn(rows, dim, generator=gen)
    fit = {"epochs": epochs, "batch_size": 128}
    conf = {"cpd": "softmax_nn", "n_classes": 6, "binning": "uniform", "within_bin": within_bin,
            "within_bin_clip": clip, "fit": fit}
    return _fit(g, {"p": dict(conf), "q": dict(conf)}, {"p": p, "q": q})


def kde_model(rows=300, seed=0, max_points=128):
    import networkx as nx

    gen = torch.Generator().manual_seed(seed)
    g = nx.DiGraph()
    g.add_edges_from([("p", "y"), ("p2", "y")])
    p = torch.randn(rows, 1, generator=gen)
    p2 = torch.randn(rows, 1, generator=gen)
    y = torch.sin(p) + 0.2 * p2 + 0.1 * torch.randn(rows, 1, generator=gen)
    cpds = {
        "p": {"cpd": "kde", "bandwidth": 0.3, "max_points": max_points},
        "p2": {"cpd": "linear_gaussian"},
        "y": {"cpd": "kde", "bandwidth": 0.5, "parent_bandwidth": 0.4, "max_points": max_points},
    }
    return _fit(g, cpds, {"p": p, "p2": p2, "y": y})
