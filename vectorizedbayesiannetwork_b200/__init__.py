"""B200-native (sm_100a) backend for VBN's batched posterior-inference hot path.

Public surface mirrors the reference for that path: ``VBN`` (set_inference_method /
infer_posterior / set_sampling_method / sample / get_cpd), ``Query``, the inference and sampling
registries, the five CPD kinds, ``CPDHandle``; plus ``install`` (drop-in into a live reference
``vbn`` package) and ``dist.Shard`` (multi-GPU).  No CPU fallback: everything numeric goes through
libvbn_cuda.so."""
from . import inference as _inference  # noqa: F401  (populates the registries)
from .core import (INFERENCE_REGISTRY, SAMPLING_REGISTRY, ConfigItem, CPDOutput, Query, StaticDAG, VBN)
from .cpd_handle import CPDHandle
from .cpds import (BaseCPD, CategoricalEmbeddedSoftmaxCPD, CategoricalTableCPD, GaussianNNCPD, KDECPD, LinearGaussianCPD, MDNCPD, RFFGaussianCPD, SoftmaxNNCPD,
                   cpd_from_spec, wrap_cpd)
from .dist import Shard, auto_shard
from .inference import (AncestralSampler, CategoricalExact, GaussianExact, GibbsSampler, ImportanceSampling,
                        LikelihoodWeighting, LoopyBeliefPropagation, MonteCarloMarginalization, RaoBlackwellizedMarginalization,
                        ResampledImportanceSampling)
from .install import install, uninstall

__all__ = [
    "VBN", "Query", "CPDOutput", "StaticDAG", "ConfigItem", "CPDHandle",
    "INFERENCE_REGISTRY", "SAMPLING_REGISTRY",
    "BaseCPD", "LinearGaussianCPD", "GaussianNNCPD", "MDNCPD", "SoftmaxNNCPD", "KDECPD", "RFFGaussianCPD", "CategoricalTableCPD", "CategoricalEmbeddedSoftmaxCPD",
    "cpd_from_spec", "wrap_cpd",
    "ImportanceSampling", "LikelihoodWeighting", "MonteCarloMarginalization", "AncestralSampler", "GibbsSampler", "LoopyBeliefPropagation",
    "GaussianExact", "CategoricalExact", "ResampledImportanceSampling", "RaoBlackwellizedMarginalization",
    "Shard", "auto_shard", "install", "uninstall",
]
