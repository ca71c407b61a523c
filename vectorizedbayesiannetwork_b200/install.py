"""Drop-in installation into a live reference ``vbn`` package: replaces the registry entries of the
hot-path methods (vbn/core/registry.py:7-11) with the CUDA classes, so
``vbn.VBN(...).set_inference_method("importance_sampling")`` transparently runs on the B200 path.
The registry decorator refuses duplicate keys (registry.py:18-20), hence plain dict assignment."""
from __future__ import annotations

from .inference import (AncestralSampler, CategoricalExact, GaussianExact, GibbsSampler, ImportanceSampling,
                        LikelihoodWeighting, LoopyBeliefPropagation, MonteCarloMarginalization, RaoBlackwellizedMarginalization,
                        ResampledImportanceSampling)

_ORIGINAL = {}


def install(vbn_module=None) -> None:
    if vbn_module is None:
        import vbn as vbn_module  # the reference package must be importable
    reg = vbn_module.core.registry
    for key, cls in (("likelihood_weighting", LikelihoodWeighting),
                     ("importance_sampling", ImportanceSampling),
                     ("monte_carlo_marginalization", MonteCarloMarginalization),
                     ("gaussian_exact", GaussianExact), ("categorical_exact", CategoricalExact),
                     ("resampled_importance_sampling", ResampledImportanceSampling),
                     ("rao_blackwellized_marginalization", RaoBlackwellizedMarginalization),
                     ("lbp", LoopyBeliefPropagation)):
        _ORIGINAL.setdefault(("inference", key), reg.INFERENCE_REGISTRY.get(key))
        reg.INFERENCE_REGISTRY[key] = cls
    for key, cls in (("ancestral", AncestralSampler), ("gibbs", GibbsSampler)):
        _ORIGINAL.setdefault(("sampling", key), reg.SAMPLING_REGISTRY.get(key))
        reg.SAMPLING_REGISTRY[key] = cls


def uninstall(vbn_module=None) -> None:
    if vbn_module is None:
        import vbn as vbn_module
    reg = vbn_module.core.registry
    for (cat, key), cls in _ORIGINAL.items():
        if cls is not None:
            (reg.INFERENCE_REGISTRY if cat == "inference" else reg.SAMPLING_REGISTRY)[key] = cls
    _ORIGINAL.clear()
