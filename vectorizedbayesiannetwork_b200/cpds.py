"""Host-side mirror of the reference CPD classes on the hot path (vbn/cpds/{linear_gaussian,
gaussian_nn,mdn,softmax_nn,kde}.py).  These objects only HOLD parameters and know how to pack
them for the CUDA schedule kernel; ``sample`` / ``log_prob`` / ``forward`` run on the GPU through
libvbn_cuda.so (see engine.py) -- there is no CPU implementation here.  Fitting is out of scope:
parameters come from a fitted reference model (``from_reference``), from a spec dict
(``from_spec``, the format of tests/golden) or from the caller.
"""
from __future__ import annotations

import itertools
import os
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L

LOG_2PI = math.log(2.0 * math.pi)
_UID = itertools.count(1)  # object ids can be recycled by CPython; cache keys use these instead


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def _f32(t) -> torch.Tensor:
    return torch.as_tensor(t).detach().to(device="cpu", dtype=torch.float32).contiguous()


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().astype(np.float32, copy=False).reshape(-1)


def _padded(a: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float32)
    out[: a.size] = a.reshape(-1)
    return out


@dataclass
class Packed:
    """What the plan compiler needs to know about one CPD."""

    kind: int
    dim: int
    n_par: int
    params: np.ndarray
    n_layers: int = 0
    act: int = 0
    n_out: int = 0
    k: int = 0
    layer_dim: List[int] = field(default_factory=list)
    aux: List[int] = field(default_factory=lambda: [0, 0, 0, 0])
    n_normals: int = 0   # per-row normals consumed by one sample()
    n_uniforms: int = 0  # per-row uniforms consumed by one sample()
    scratch: int = 0     # per-row scratch floats
    heavy: bool = False
    tc_blob: Optional[np.ndarray] = None  # tensor-core weight image (pack_mlp_tc), if eligible
    tc_k1: int = 0
    tc_n3: int = 0


def pack_mlp(layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], input_dim: int) -> Tuple[np.ndarray, List[int]]:
    """Layout documented in csrc/vbn_schedule.cuh: hidden layers transposed [in][pad4(out)], the
    last layer in nn.Linear layout [out][pad4(in)], every bias padded to a multiple of 4."""
    if len(layers) > L.MAX_LAYERS:
        raise ValueError(f"MLP deeper than {L.MAX_LAYERS} Linear layers is not supported")
    chunks: List[np.ndarray] = []
    dims: List[int] = []
    last = len(layers) - 1
    cur_in = int(input_dim)
    for i, (w, b) in enumerate(layers):
        w, b = _f32(w), _f32(b)
        out_dim, in_dim = int(w.shape[0]), int(w.shape[1])
        if in_dim != cur_in:
            raise ValueError(f"layer {i} expects {in_dim} inputs, got {cur_in}")
        if i != last:
            buf = np.zeros((in_dim, _pad4(out_dim)), dtype=np.float32)
            buf[:, :out_dim] = w.numpy().T
        else:
            buf = np.zeros((out_dim, _pad4(in_dim)), dtype=np.float32)
            buf[:, :in_dim] = w.numpy()
        chunks.append(buf.reshape(-1))
        chunks.append(_padded(b.numpy(), _pad4(out_dim)))
        dims.append(out_dim)
        cur_in = out_dim
    fast = len(layers) == 3 and dims[0] == 32 and dims[1] == 32
    if not fast:
        widest = max([int(input_dim)] + dims[:-1]) if layers else 0
        if widest > L.MAX_GENERIC_WIDTH:
            raise ValueError(
                f"MLP width {widest} exceeds the generic device path limit {L.MAX_GENERIC_WIDTH}"
            )
    return np.concatenate(chunks) if chunks else np.zeros(0, np.float32), dims


def _split_tf32(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """w = hi + lo with hi, lo representable in tf32 (10 explicit mantissa bits), round to nearest."""
    def rn(x):
        bits = x.astype(np.float32).view(np.uint32)
        return ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    hi = rn(w)
    lo = rn(w.astype(np.float32) - hi)
    return hi, lo


def _core_matrix_image(w: np.ndarray) -> np.ndarray:
    """[N][K] (K-major) -> the no-swizzle UMMA shared-memory image: core matrices of 8 rows x 16 bytes
    (4 floats), K-adjacent core matrices 128 B apart, 8-row groups K*32 B apart
    (csrc/vbn_schedule_tc.cuh make_b_desc)."""
    n, k = w.shape
    assert n % 8 == 0 and k % 4 == 0
    return np.ascontiguousarray(w.reshape(n // 8, 8, k // 4, 4).transpose(0, 2, 1, 3)).reshape(-1)


def pack_mlp_tc(layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], input_dim: int, l1_fma: bool = False,
                relu2: bool = False):
    """Weight image for the tcgen05 kernel, or None when the MLP is not [Dp<=32 -> 32 -> 32 -> O<=32].
    Layout (floats): W1hi, W1lo [32][K1+8]; W2hi, W2lo [32][40]; W3hi, W3lo [N3][40] -- core-matrix images of
    nn.Linear's [out][in] = [N][K] K-major weight with the bias appended as column K (then 7 zero columns).
    ``l1_fma`` (gaussian_nn / mdn nodes, whose ops carry their parent slots in the descriptor) and Dp <= 4: the first
    layer runs on the FP32 pipe, so the W1 images are replaced by the plain block W1^T[4][32] (rows >= Dp zero),
    b1[32], and K1 is reported as 0 (csrc/vbn_schedule_tc.cuh hidden1_fma).
    ``relu2`` (ReLU MLPs): the kernel stores 2 relu(h) = h + |h| as the hidden activations (one packed add instead of
    two max), so W2 and W3 -- not the biases -- are halved here: an exact power-of-two scaling on both sides, every
    product and the accumulator are bit-identical to relu(h) * w."""
    if len(layers) != 3:
        return None
    (w1, b1), (w2, b2), (w3, b3) = [(_f32(w).numpy(), _f32(b).numpy()) for w, b in layers]
    dp, n_out = int(input_dim), int(w3.shape[0])
    if w1.shape != (32, dp) or w2.shape != (32, 32) or w3.shape[1] != 32 or not (1 <= dp <= 32) or not (1 <= n_out <= 32):
        return None
    k1 = (dp + 7) & ~7
    n3 = (n_out + 15) & ~15
    w1p = np.zeros((32, k1), np.float32)
    w1p[:, :dp] = w1
    w3p = np.zeros((n3, 32), np.float32)
    w3p[:n_out] = w3
    chunks = []
    half = np.float32(0.5 if relu2 else 1.0)
    mma_layers = [(w1p, b1), (w2 * half, b2), (w3p * half, _padded(b3, n3))]
    if l1_fma and dp <= 4:
        plain = np.zeros((5, 32), np.float32)
        plain[:dp] = w1.T
        plain[4] = b1
        chunks.append(plain.reshape(-1))
        mma_layers = mma_layers[1:]
        k1 = 0
    for w, b in mma_layers:
        # the bias rides in an extra K = 8 step: column K holds b, columns K+1..K+7 are zero; the kernel pairs it
        # with a constant A block (1, 0, ..., 0) so the first MMA of a layer writes the bias into the accumulator
        ext = np.zeros((w.shape[0], w.shape[1] + 8), np.float32)
        ext[:, : w.shape[1]] = w
        ext[:, w.shape[1]] = b
        hi, lo = _split_tf32(ext)
        chunks += [_core_matrix_image(hi), _core_matrix_image(lo)]
    return np.concatenate(chunks), k1, n3


def _layers_from_module(net) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    return [(m.weight, m.bias) for m in net if hasattr(m, "weight") and hasattr(m, "bias")]


def _with_tc(pk: Packed, layers, input_dim: int) -> Packed:
    img = pack_mlp_tc(layers, input_dim, l1_fma=pk.kind in (L.OP_GNN, L.OP_MDN) and os.environ.get("VBN_TC_L1FMA", "1") != "0",
                      relu2=pk.act == L.ACT["relu"])
    if img is not None:
        pk.tc_blob, pk.tc_k1, pk.tc_n3 = img
    return pk


class BaseCPD:
    """Signature mirror of vbn/core/base.py:28-81 (sample / log_prob / forward)."""

    kind = "base"

    def __init__(self, input_dim: int, output_dim: int, device=None) -> None:
        self.input_dim = int(input_dim)
        self.output_dim = int(output_dim)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self._packed: Optional[Packed] = None
        self._version = 0
        self._uid = next(_UID)

    # --- packing -----------------------------------------------------------------------
    def pack(self) -> Packed:
        if self._packed is None:
            self._packed = self._pack()
        return self._packed

    def _pack(self) -> Packed:  # pragma: no cover - abstract
        raise NotImplementedError

    def invalidate(self) -> None:
        """Call after changing parameters in place: drops the packed blob and cached plans."""
        self._packed = None
        self._version += 1

    # --- public CPD API (GPU) ------------------------------------------------------------
    def sample(self, parents: Optional[torch.Tensor], n_samples: int, *, noise=None, seed=None) -> torch.Tensor:
        from .engine import cpd_sample

        return cpd_sample(self, parents, int(n_samples), noise=noise, seed=seed)

    def log_prob(self, x: torch.Tensor, parents: Optional[torch.Tensor]) -> torch.Tensor:
        from .engine import cpd_log_prob

        return cpd_log_prob(self, x, parents)

    def forward(self, parents: Optional[torch.Tensor], n_samples: int):
        from .core import CPDOutput

        samples = self.sample(parents, n_samples)
        log_prob = self.log_prob(samples, parents)
        return CPDOutput(samples=samples, log_prob=log_prob, pdf=torch.exp(log_prob))

    __call__ = forward

    def to_spec(self) -> dict:  # pragma: no cover - abstract
        raise NotImplementedError

    def param_width(self) -> int:
        """Floats per row written by the parameter read-out (engine.cpd_params); 0 = none."""
        return 0

    def params(self, parents: Optional[torch.Tensor]) -> torch.Tensor:
        from .engine import cpd_params

        return cpd_params(self, parents)


# ------------------------------------------------------------------------------------------
class LinearGaussianCPD(BaseCPD):
    """vbn/cpds/linear_gaussian.py:14-217 (buffers _weight[Dp,D], _bias[D], _var[D])."""

    kind = "linear_gaussian"

    def __init__(self, input_dim, output_dim, weight, bias, var, min_scale: float = 1e-3, device=None):
        super().__init__(input_dim, output_dim, device)
        self._weight = _f32(weight).reshape(self.input_dim, self.output_dim)
        self._bias = _f32(bias).reshape(self.output_dim)
        self._var = _f32(var).reshape(self.output_dim)
        self.min_scale = float(min_scale)

    def _scale(self) -> torch.Tensor:
        return torch.sqrt(self._var.clamp(min=self.min_scale**2))  # linear_gaussian.py:163-165

    def param_width(self) -> int:
        return 2 * self.output_dim

    def _pack(self) -> Packed:
        scale = self._scale()
        params = np.concatenate(
            [_np(self._weight), _np(self._bias), _np(scale), _np(2 * torch.log(scale)), _np(scale**2)]
        )
        return Packed(kind=L.OP_LG, dim=self.output_dim, n_par=self.input_dim, params=params,
                      n_normals=self.output_dim)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], c["weight"], c["bias"], c["var"], c["min_scale"], device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        return cls(cpd.input_dim, cpd.output_dim, cpd._weight, cpd._bias, cpd._var, cpd.min_scale, device)

    def to_spec(self):
        return {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "min_scale": self.min_scale, "weight": self._weight, "bias": self._bias, "var": self._var}


# ------------------------------------------------------------------------------------------
class GaussianNNCPD(BaseCPD):
    """vbn/cpds/gaussian_nn.py:37-288."""

    kind = "gaussian_nn"

    def __init__(self, input_dim, output_dim, *, layers=None, loc=None, log_scale=None,
                 mean_x=None, std_x=None, mean_y=None, std_y=None, activation="relu",
                 min_scale: float = 1e-3, device=None):
        super().__init__(input_dim, output_dim, device)
        d, dp = self.output_dim, self.input_dim
        self.activation = str(activation)
        self.min_scale = float(min_scale)
        self.mean_x = _f32(mean_x if mean_x is not None else torch.zeros(dp)).reshape(dp)
        self.std_x = _f32(std_x if std_x is not None else torch.ones(dp)).reshape(dp)
        self.mean_y = _f32(mean_y if mean_y is not None else torch.zeros(d)).reshape(d)
        self.std_y = _f32(std_y if std_y is not None else torch.ones(d)).reshape(d)
        if dp == 0:
            self._loc = _f32(loc).reshape(d)
            self._log_scale = _f32(log_scale).reshape(d)
            self.layers = None
        else:
            self.layers = [(_f32(w), _f32(b)) for w, b in layers]

    def param_width(self) -> int:
        return 2 * self.output_dim

    def _pack(self) -> Packed:
        d, dp = self.output_dim, self.input_dim
        if dp == 0:
            # root: fixed Normal(loc*std_y+mean_y, (softplus(log_scale)+min)*std_y)
            # (gaussian_nn.py:244-254); same density family as linear_gaussian with Dp = 0
            loc = self._loc * self.std_y + self.mean_y
            scale = (F.softplus(self._log_scale) + self.min_scale) * self.std_y
            params = np.concatenate([_np(loc), _np(scale), _np(2 * torch.log(scale)), _np(scale**2)])
            return Packed(kind=L.OP_LG, dim=d, n_par=0, params=params, n_normals=d)
        head = np.concatenate([_np(self.mean_x), _np(self.std_x), _np(self.mean_y), _np(self.std_y),
                               np.array([self.min_scale], np.float32)])
        mlp, dims = pack_mlp(self.layers, dp)
        if dims[-1] != 2 * d:
            raise ValueError("gaussian_nn net must output 2*output_dim values")
        params = np.concatenate([_padded(head, _pad4(2 * dp + 2 * d + 1)), mlp])
        return _with_tc(Packed(kind=L.OP_GNN, dim=d, n_par=dp, params=params, n_layers=len(dims),
                               act=L.ACT[self.activation], n_out=2 * d, layer_dim=dims, n_normals=d,
                               scratch=2 * d, heavy=True), self.layers, dp)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], layers=c.get("layers"), loc=c.get("loc"),
                   log_scale=c.get("log_scale"), mean_x=c["mean_x"], std_x=c["std_x"],
                   mean_y=c["mean_y"], std_y=c["std_y"], activation=c["activation"],
                   min_scale=c["min_scale"], device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        root = cpd.input_dim == 0
        return cls(cpd.input_dim, cpd.output_dim,
                   layers=None if root else _layers_from_module(cpd.net),
                   loc=cpd._loc if root else None, log_scale=cpd._log_scale if root else None,
                   mean_x=cpd.mean_x, std_x=cpd.std_x, mean_y=cpd.mean_y, std_y=cpd.std_y,
                   activation=cpd.activation, min_scale=cpd.min_scale, device=device)

    def to_spec(self):
        out = {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
               "min_scale": self.min_scale, "activation": self.activation, "mean_x": self.mean_x,
               "std_x": self.std_x, "mean_y": self.mean_y, "std_y": self.std_y}
        if self.input_dim == 0:
            out.update(loc=self._loc, log_scale=self._log_scale)
        else:
            out["layers"] = self.layers
        return out


# ------------------------------------------------------------------------------------------
class RFFGaussianCPD(BaseCPD):
    """vbn/cpds/rff_gaussian.py:14-291: random-Fourier-feature ridge regression with a
    constant Gaussian noise scale.  Inference-time state only (buffers of a fitted reference object)."""

    kind = "rff_gaussian"

    def __init__(self, input_dim, output_dim, *, n_features, rff_w=None, rff_b=None, coef=None, bias=None, var=None,
                 mean_x=None, std_x=None, mean_y=None, std_y=None, min_scale: float = 1e-3, stats_ready: bool = True,
                 device=None):
        super().__init__(input_dim, output_dim, device)
        d, dp, f = self.output_dim, self.input_dim, int(n_features)
        self.n_features = f
        self.min_scale = float(min_scale)
        self._stats_ready = bool(stats_ready)
        self.mean_x = _f32(mean_x if mean_x is not None else torch.zeros(dp)).reshape(dp)
        self.std_x = _f32(std_x if std_x is not None else torch.ones(dp)).reshape(dp)
        self.mean_y = _f32(mean_y if mean_y is not None else torch.zeros(d)).reshape(d)
        self.std_y = _f32(std_y if std_y is not None else torch.ones(d)).reshape(d)
        self._rff_w = _f32(rff_w if rff_w is not None else torch.zeros(f, dp)).reshape(f, dp)
        self._rff_b = _f32(rff_b if rff_b is not None else torch.zeros(f)).reshape(f)
        self._coef = _f32(coef if coef is not None else torch.zeros(f, d)).reshape(f, d)
        self._bias = _f32(bias if bias is not None else torch.zeros(d)).reshape(d)
        self._var = _f32(var if var is not None else torch.ones(d)).reshape(d)

    def _scale(self) -> torch.Tensor:
        return torch.sqrt(self._var.clamp(min=self.min_scale**2))  # rff_gaussian.py:181-183

    def param_width(self) -> int:
        return 2 * self.output_dim

    def _pack(self) -> Packed:
        if not self._stats_ready:  # rff_gaussian.py:76-78
            raise RuntimeError("RFFGaussianCPD is not fitted yet.")
        d, dp, f = self.output_dim, self.input_dim, self.n_features
        scale = self._scale()
        if dp == 0:
            # root: Normal(mean_y, scale) (rff_gaussian.py:188-191, 256-259) -- linear_gaussian with Dp = 0
            params = np.concatenate([_np(self.mean_y), _np(scale), _np(2 * torch.log(scale)), _np(scale**2)])
            return Packed(kind=L.OP_LG, dim=d, n_par=0, params=params, n_normals=d)
        head = np.array([f, math.sqrt(2.0 / float(f)), 0.0, 0.0], np.float32)
        stats = _padded(np.concatenate([_np(self.mean_x), _np(self.std_x)]), _pad4(2 * dp))
        tail = _padded(np.concatenate([_np(self._bias), _np(self.mean_y), _np(self.std_y), _np(scale),
                                       _np(2 * torch.log(scale)), _np(scale**2)]), _pad4(6 * d))
        wb = torch.cat([self._rff_w, self._rff_b.reshape(f, 1)], dim=1)  # per feature: w_f[Dp], b_f
        params = np.concatenate([head, stats, tail, _np(wb), _np(self._coef)])
        return Packed(kind=L.OP_RFF, dim=d, n_par=dp, params=_padded(params, _pad4(params.size)), n_normals=d,
                      scratch=dp + d, heavy=True)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], n_features=c["n_features"], rff_w=c["rff_w"], rff_b=c["rff_b"],
                   coef=c["coef"], bias=c["bias"], var=c["var"], mean_x=c["mean_x"], std_x=c["std_x"],
                   mean_y=c["mean_y"], std_y=c["std_y"], min_scale=c["min_scale"],
                   stats_ready=c.get("stats_ready", True), device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        return cls(cpd.input_dim, cpd.output_dim, n_features=cpd.n_features, rff_w=cpd._rff_w, rff_b=cpd._rff_b,
                   coef=cpd._coef, bias=cpd._bias, var=cpd._var, mean_x=cpd.mean_x, std_x=cpd.std_x,
                   mean_y=cpd.mean_y, std_y=cpd.std_y, min_scale=cpd.min_scale,
                   stats_ready=bool(cpd._stats_ready.item()), device=device)

    def to_spec(self):
        return {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "n_features": self.n_features, "min_scale": self.min_scale, "mean_x": self.mean_x,
                "std_x": self.std_x, "mean_y": self.mean_y, "std_y": self.std_y, "rff_w": self._rff_w,
                "rff_b": self._rff_b, "coef": self._coef, "bias": self._bias, "var": self._var,
                "stats_ready": self._stats_ready}


# ------------------------------------------------------------------------------------------
class MDNCPD(BaseCPD):
    """vbn/cpds/mdn.py:37-272."""

    kind = "mdn"

    def __init__(self, input_dim, output_dim, n_components, *, layers=None, logits=None, loc=None,
                 log_scale=None, activation="relu", min_scale: float = 1e-3, device=None):
        super().__init__(input_dim, output_dim, device)
        self.n_components = int(n_components)
        self.activation = str(activation)
        self.min_scale = float(min_scale)
        k, d = self.n_components, self.output_dim
        if self.input_dim == 0:
            self._logits = _f32(logits).reshape(k)
            self._loc = _f32(loc).reshape(k, d)
            self._log_scale = _f32(log_scale).reshape(k, d)
            self.layers = None
        else:
            self.layers = [(_f32(w), _f32(b)) for w, b in layers]

    def param_width(self) -> int:
        return self.n_components * (1 + 2 * self.output_dim)

    def _pack(self) -> Packed:
        k, d, dp = self.n_components, self.output_dim, self.input_dim
        n_out = k + 2 * k * d
        head = _padded(np.array([self.min_scale], np.float32), 4)
        if dp == 0:
            rest = torch.cat([self._loc, self._log_scale], dim=1).reshape(-1)  # per k: loc[D], raw[D]
            consts = np.concatenate([_np(self._logits), _np(rest)])
            params = np.concatenate([head, _padded(consts, _pad4(n_out))])
            return Packed(kind=L.OP_MDN, dim=d, n_par=0, params=params, n_layers=0, n_out=n_out, k=k,
                          n_normals=d, n_uniforms=1, scratch=n_out, heavy=True)
        mlp, dims = pack_mlp(self.layers, dp)
        if dims[-1] != n_out:
            raise ValueError("mdn net must output K*(2D+1) values")
        return _with_tc(Packed(kind=L.OP_MDN, dim=d, n_par=dp, params=np.concatenate([head, mlp]),
                               n_layers=len(dims), act=L.ACT[self.activation], n_out=n_out, k=k,
                               layer_dim=dims, n_normals=d, n_uniforms=1, scratch=n_out, heavy=True),
                        self.layers, dp)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], c["n_components"], layers=c.get("layers"),
                   logits=c.get("logits"), loc=c.get("loc"), log_scale=c.get("log_scale"),
                   activation=c["activation"], min_scale=c["min_scale"], device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        root = cpd.input_dim == 0
        return cls(cpd.input_dim, cpd.output_dim, cpd.n_components,
                   layers=None if root else _layers_from_module(cpd.net),
                   logits=cpd._logits if root else None, loc=cpd._loc if root else None,
                   log_scale=cpd._log_scale if root else None, activation=cpd.activation,
                   min_scale=cpd.min_scale, device=device)

    def to_spec(self):
        out = {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
               "min_scale": self.min_scale, "activation": self.activation,
               "n_components": self.n_components}
        if self.input_dim == 0:
            out.update(logits=self._logits, loc=self._loc, log_scale=self._log_scale)
        else:
            out["layers"] = self.layers
        return out


# ------------------------------------------------------------------------------------------
class SoftmaxNNCPD(BaseCPD):
    """vbn/cpds/softmax_nn.py:40-759 (inference-time state only)."""

    kind = "softmax_nn"

    def __init__(self, input_dim, output_dim, n_classes, *, layers=None, logits=None,
                 root_log_probs=None, root_ready=False, bin_edges=None, class_values=None,
                 sample_values=None, is_discrete=None, bins_ready=True, activation="relu",
                 temperature: float = 1.0, min_bin_width: float = 1e-12, within_bin="uniform",
                 within_bin_scale: float = 0.25, within_bin_clip: bool = False, device=None):
        super().__init__(input_dim, output_dim, device)
        d, k = self.output_dim, int(n_classes)
        self.n_classes = k
        self.activation = str(activation)
        self.temperature = float(temperature)
        self.min_bin_width = float(min_bin_width)
        self.within_bin = str(within_bin)
        if self.within_bin not in L.WITHIN_BIN:
            raise ValueError(f"Unknown within_bin '{within_bin}'")
        self.within_bin_scale = float(within_bin_scale)
        self.within_bin_clip = bool(within_bin_clip)
        self._bin_edges = _f32(bin_edges).reshape(d, k + 1)
        self._class_values = _f32(class_values).reshape(d, k)
        self._sample_values = _f32(sample_values).reshape(d, k)
        self._is_discrete = torch.as_tensor(is_discrete).detach().cpu().bool().reshape(d)
        self._bins_ready = bool(bins_ready)
        self._root_ready = bool(root_ready)
        self._root_log_probs = _f32(root_log_probs if root_log_probs is not None else torch.zeros(d, k)).reshape(d, k)
        if self.input_dim == 0:
            self._logits = _f32(logits if logits is not None else torch.zeros(d, k)).reshape(d, k)
            self.layers = None
        else:
            self.layers = [(_f32(w), _f32(b)) for w, b in layers]

    def param_width(self) -> int:
        return self.output_dim * self.n_classes

    def _ensure_bins_ready(self) -> None:
        if not self._bins_ready:  # softmax_nn.py:178-180
            raise RuntimeError("Bins not initialized. Call fit(...) before sampling.")

    def _pack(self) -> Packed:
        self._ensure_bins_ready()
        d, k, dp = self.output_dim, self.n_classes, self.input_dim
        head = np.concatenate([
            np.array([self.min_bin_width, self.within_bin_scale, self.temperature, 0.0], np.float32),
            _np(self._bin_edges), _np(self._class_values), _np(self._sample_values),
            _np(self._is_discrete.float()),
        ])
        head = _padded(head, _pad4(4 + d * (k + 1) + 2 * d * k + d))
        aux = [L.WITHIN_BIN[self.within_bin], int(self.within_bin_clip), int(bool(self._is_discrete.any())), 0]
        gaussian = self.within_bin == "gaussian"
        common = dict(kind=L.OP_SNN, dim=d, k=k, n_out=d * k, aux=aux, scratch=d * k, heavy=True,
                      # per row: D picks, plus D within-bin uniforms unless the law is gaussian (normals) or every dim is
                      # discrete (no within-bin draw at all: keeps four consecutive discrete nodes on one Philox block)
                      n_normals=d if gaussian else 0,
                      n_uniforms=d if (gaussian or bool(self._is_discrete.all())) else 2 * d)
        if dp == 0:
            t = self.temperature
            if self._root_ready:  # softmax_nn.py:636-641
                logits = torch.log_softmax(self._root_log_probs / t, dim=-1)
            else:
                logits = self._logits / t
            params = np.concatenate([head, _padded(_np(logits), _pad4(d * k))])
            return Packed(n_par=0, params=params, n_layers=0, **common)
        mlp, dims = pack_mlp(self.layers, dp)
        if dims[-1] != d * k:
            raise ValueError("softmax_nn net must output D*C logits")
        return _with_tc(Packed(n_par=dp, params=np.concatenate([head, mlp]), n_layers=len(dims),
                               act=L.ACT[self.activation], layer_dim=dims, **common), self.layers, dp)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], c["n_classes"], layers=c.get("layers"),
                   logits=c.get("logits"), root_log_probs=c.get("root_log_probs"),
                   root_ready=c.get("root_ready", False), bin_edges=c["bin_edges"],
                   class_values=c["class_values"], sample_values=c["sample_values"],
                   is_discrete=c["is_discrete"], bins_ready=c.get("bins_ready", True),
                   activation=c["activation"], temperature=c.get("temperature", 1.0),
                   min_bin_width=c["min_bin_width"], within_bin=c["within_bin"],
                   within_bin_scale=c["within_bin_scale"], within_bin_clip=c["within_bin_clip"],
                   device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        root = cpd.input_dim == 0
        return cls(cpd.input_dim, cpd.output_dim, cpd.n_classes,
                   layers=None if root else _layers_from_module(cpd.net),
                   logits=cpd._logits if root else None, root_log_probs=cpd._root_log_probs,
                   root_ready=bool(cpd._root_ready.item()), bin_edges=cpd._bin_edges,
                   class_values=cpd._class_values, sample_values=cpd._sample_values,
                   is_discrete=cpd._is_discrete, bins_ready=bool(cpd._bins_ready.item()),
                   activation=cpd.activation, temperature=cpd.temperature,
                   min_bin_width=cpd.min_bin_width, within_bin=cpd.within_bin,
                   within_bin_scale=cpd.within_bin_scale, within_bin_clip=cpd.within_bin_clip,
                   device=device)

    def to_spec(self):
        out = {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
               "n_classes": self.n_classes, "activation": self.activation,
               "temperature": self.temperature, "min_bin_width": self.min_bin_width,
               "within_bin": self.within_bin, "within_bin_scale": self.within_bin_scale,
               "within_bin_clip": self.within_bin_clip, "bin_edges": self._bin_edges,
               "class_values": self._class_values, "sample_values": self._sample_values,
               "is_discrete": self._is_discrete, "bins_ready": self._bins_ready,
               "root_ready": self._root_ready, "root_log_probs": self._root_log_probs}
        if self.input_dim == 0:
            out["logits"] = self._logits
        else:
            out["layers"] = self.layers
        return out


# ------------------------------------------------------------------------------------------
class KDECPD(BaseCPD):
    """vbn/cpds/kde.py:13-182 (stored points _parents[N,Dp], _targets[N,Dx])."""

    kind = "kde"

    def __init__(self, input_dim, output_dim, *, parents=None, targets=None, bandwidth: float = 1.0,
                 parent_bandwidth: Optional[float] = None, min_scale: float = 1e-3, device=None):
        super().__init__(input_dim, output_dim, device)
        self.bandwidth = float(bandwidth)
        self.parent_bandwidth = float(parent_bandwidth) if parent_bandwidth is not None else float(bandwidth)
        self.min_scale = float(min_scale)
        self._targets = None if targets is None else _f32(targets).reshape(-1, self.output_dim)
        n = 0 if self._targets is None else self._targets.shape[0]
        if parents is None or self.input_dim == 0:
            self._parents = torch.zeros(n, 0)
        else:
            self._parents = _f32(parents).reshape(n, self.input_dim)
        self._dev_points = None  # (parents, targets) on the device for the stand-alone kernel

    def _check_fitted(self) -> None:
        if self._targets is None:  # kde.py:64-66
            raise RuntimeError("KDECPD is not fitted yet.")

    def scales(self) -> Tuple[float, float]:
        sy = max(self.bandwidth, 1e-3) + self.min_scale  # kde.py:105-109
        sp = max(self.parent_bandwidth, 1e-3) + self.min_scale
        return sy, sp

    def _pack(self) -> Packed:
        self._check_fitted()
        d, dp = self.output_dim, self.input_dim
        n = int(self._targets.shape[0])
        sy, sp = self.scales()
        head = np.array([0.5 / (sy * sy), 0.5 / (sp * sp), -0.5 * d * (LOG_2PI + 2.0 * math.log(sy)),
                         max(self.bandwidth, 1e-3) + self.min_scale, math.log(float(n)), 0, 0, 0], np.float32)
        params = np.concatenate([head, _padded(_np(self._parents), _pad4(n * dp)),
                                 _padded(_np(self._targets), _pad4(n * d))])
        return Packed(kind=L.OP_KDE, dim=d, n_par=dp, params=params, k=n, n_normals=d, n_uniforms=1,
                      heavy=True)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], parents=c.get("parents"), targets=c.get("targets"),
                   bandwidth=c["bandwidth"], parent_bandwidth=c["parent_bandwidth"],
                   min_scale=c["min_scale"], device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        return cls(cpd.input_dim, cpd.output_dim, parents=cpd._parents, targets=cpd._targets,
                   bandwidth=cpd.bandwidth, parent_bandwidth=cpd.parent_bandwidth,
                   min_scale=cpd.min_scale, device=device)

    def to_spec(self):
        return {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "bandwidth": self.bandwidth, "parent_bandwidth": self.parent_bandwidth,
                "min_scale": self.min_scale, "parents": self._parents, "targets": self._targets}


# ------------------------------------------------------------------------------------------
class CategoricalTableCPD(BaseCPD):
    """vbn/cpds/categorical_table.py:24-417: count table with Dirichlet smoothing.
    Inference-time state only (counts already smoothed by the reference's fit).  D = 1 on the GPU path."""

    kind = "categorical_table"

    def __init__(self, input_dim, output_dim, *, counts, class_values, class_mask, parent_values=(),
                 parent_strides=(), n_classes=None, device=None):
        super().__init__(input_dim, output_dim, device)
        self._counts = _f32(counts)
        d = self.output_dim
        self._class_values = _f32(class_values).reshape(d, -1)
        self._sample_values = self._class_values
        self._class_mask = torch.as_tensor(class_mask).detach().cpu().bool().reshape(d, -1)
        self.n_classes = int(n_classes if n_classes is not None else self._class_values.shape[1])
        self._parent_values = [_f32(v).reshape(-1) for v in parent_values]
        self._parent_strides = [int(v) for v in parent_strides]
        if len(self._parent_values) != self.input_dim:
            raise ValueError("categorical_table needs one support vector per parent dim")

    @property
    def n_support(self) -> int:
        return int(self._class_mask[0].sum())

    def param_width(self) -> int:
        return self.n_classes  # probs[C] of the row's parent configuration (op_params, VBN_OP_TAB)

    def logits_table(self) -> torch.Tensor:
        """[n_cfg, C] log-probabilities: log_softmax(log(clamp(counts / sum, 1e-12)))  (:359-369, :413)."""
        probs = self._counts[0] / self._counts[0].sum(dim=-1, keepdim=True).clamp_min(1e-12)
        return torch.log_softmax(torch.log(probs.clamp_min(1e-12)), dim=-1)

    def probs(self, parents: Optional[torch.Tensor]) -> torch.Tensor:
        """softmax(_logits_from_parents) rows [B, S, 1, C] (host lookup: the parameters ARE the table)."""
        table = torch.exp(self.logits_table())
        if parents is None or self.input_dim == 0:
            return table[:1].view(1, 1, 1, -1)
        p3 = torch.as_tensor(parents).detach().cpu().float()
        p3 = p3.unsqueeze(1) if p3.dim() == 2 else p3
        idx = torch.zeros(p3.shape[:2], dtype=torch.long)
        for k, (sup, stride) in enumerate(zip(self._parent_values, self._parent_strides)):
            col = p3[..., k].contiguous()
            pos = torch.searchsorted(sup, col).clamp(max=sup.numel() - 1)
            if not bool((sup[pos] == col).all()):
                raise ValueError("Found values outside support.")  # categorical_table.py:12-21
            idx = idx + pos * stride
        return table[idx].unsqueeze(2)

    def _pack(self) -> Packed:
        if self.output_dim != 1:
            raise ValueError("categorical_table with output_dim > 1 has no CUDA implementation yet")
        if self.n_classes > TABLE_MAX_CLASSES or any(v.numel() > TABLE_MAX_CLASSES for v in self._parent_values):
            raise ValueError(f"categorical_table supports at most {TABLE_MAX_CLASSES} classes per variable on the GPU")
        if not bool(self._class_mask[0].all()):
            # masked (padding) classes keep their clamp(1e-12) mass in the reference's softmax; they
            # are kept as classes here too, but can never match an observed value
            pass
        logp = self.logits_table()
        params = _tab_params(self.n_classes, logp, self._parent_values, self._parent_strides,
                             self._class_values[0], torch.where(self._class_mask[0], self._class_values[0],
                                                                torch.full_like(self._class_values[0], float("nan"))),
                             strict=True)
        return Packed(kind=L.OP_TAB, dim=1, n_par=self.input_dim, params=params, k=self.n_classes,
                      n_normals=0, n_uniforms=1, scratch=0, heavy=False)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], counts=c["counts"], class_values=c["class_values"],
                   class_mask=c["class_mask"], parent_values=c.get("parent_values", ()),
                   parent_strides=c.get("parent_strides", ()), n_classes=c.get("n_classes"), device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        return cls(cpd.input_dim, cpd.output_dim, counts=cpd._counts, class_values=cpd._class_values,
                   class_mask=cpd._class_mask, parent_values=cpd._parent_values or (),
                   parent_strides=cpd._parent_strides or (), n_classes=cpd.n_classes, device=device)

    def to_spec(self):
        return {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "n_classes": self.n_classes, "counts": self._counts, "class_values": self._class_values,
                "class_mask": self._class_mask, "parent_values": self._parent_values,
                "parent_strides": self._parent_strides}


# ------------------------------------------------------------------------------------------
class CategoricalEmbeddedSoftmaxCPD(CategoricalTableCPD):
    """vbn/cpds/categorical_embedded_softmax.py:49-511: a categorical CPD whose logits are
    MLP(concat of per-parent embeddings).  Its parents are strictly discrete (:36-46, :280-293), so there is one
    logit vector per parent configuration: the CPD compiles to the same lookup op as categorical_table
    (VBN_OP_TAB, strict parent support).  The table is evaluated ONCE, through this library's GPU MLP path,
    over all configurations; D = 1 on the GPU path."""

    kind = "categorical_embedded_softmax"

    def __init__(self, input_dim, output_dim, *, n_classes, class_values, class_mask, parent_values=(),
                 embeddings=None, layers=None, logits=None, activation="relu", stats_ready: bool = True, device=None):
        BaseCPD.__init__(self, input_dim, output_dim, device)
        d = self.output_dim
        self.n_classes = int(n_classes)
        self.activation = str(activation)
        self._stats_ready = bool(stats_ready)
        self._class_values = _f32(class_values).reshape(d, -1)
        self._sample_values = self._class_values  # categorical_embedded_softmax.py:357-360
        self._class_mask = torch.as_tensor(class_mask).detach().cpu().bool().reshape(d, -1)
        self._parent_values = [_f32(v).reshape(-1) for v in parent_values]
        if len(self._parent_values) != self.input_dim:
            raise ValueError("categorical_embedded_softmax needs one support vector per parent dim")
        cards = [int(v.numel()) for v in self._parent_values]
        strides, acc = [], 1
        for card in reversed(cards):  # our own configuration index: last parent fastest
            strides.append(acc)
            acc *= card
        self._parent_strides = strides[::-1]
        if self.input_dim == 0:
            self._logits = _f32(logits).reshape(d, -1)
            self.embeddings, self.layers = None, None
        else:
            self._logits = None
            self.embeddings = [_f32(e) for e in embeddings]
            self.layers = [(_f32(w), _f32(b)) for w, b in layers]
        self._table: Optional[torch.Tensor] = None

    def invalidate(self) -> None:
        super().invalidate()
        self._table = None

    def logits_table(self) -> torch.Tensor:
        """[n_cfg, C] log-probabilities log_softmax(masked logits) of every parent configuration."""
        if not self._stats_ready:  # :136-138
            raise RuntimeError("CategoricalEmbeddedSoftmaxCPD is not fitted yet.")
        if self._table is not None:
            return self._table
        if self.output_dim != 1:
            raise ValueError("categorical_embedded_softmax with output_dim > 1 has no CUDA implementation yet")
        c = self.n_classes
        if self.input_dim == 0:
            # root parameters (:471-475, :497-500): a [1, C] parameter transform, like categorical_table's counts
            self._table = torch.log_softmax(self._logits[:1], dim=-1)
            return self._table
        if not bool(self._class_mask[0].all()):
            raise ValueError("categorical_embedded_softmax with padded classes has no CUDA implementation yet")
        cards = [int(v.numel()) for v in self._parent_values]
        n_cfg = int(np.prod(cards))
        if n_cfg > TABLE_MAX_CONFIGS:
            raise ValueError(f"categorical_embedded_softmax: {n_cfg} parent configurations exceed {TABLE_MAX_CONFIGS}")
        grid = torch.cartesian_prod(*[torch.arange(k) for k in cards]).reshape(n_cfg, len(cards))
        feats = torch.cat([emb[grid[:, k]] for k, emb in enumerate(self.embeddings)], dim=-1)  # :306-314
        # the MLP forward + log_softmax run on the GPU: a discrete-mode softmax_nn over the embedded features
        # has exactly this density (softmax_nn.py:581-759 with temperature 1, discrete dims add no within-bin term)
        ar = torch.arange(c, dtype=torch.float32).reshape(1, c)
        net = SoftmaxNNCPD(feats.shape[1], 1, c, layers=self.layers, bin_edges=torch.arange(c + 1).float().reshape(1, -1) - 0.5,
                           class_values=ar, sample_values=ar, is_discrete=torch.ones(1, dtype=torch.bool),
                           activation=self.activation, device=self.device)
        from .engine import cpd_log_prob

        cols = [_f32(cpd_log_prob(net, torch.full((n_cfg, 1), float(k)), feats)).reshape(n_cfg) for k in range(c)]
        self._table = torch.stack(cols, dim=1)
        return self._table

    def _pack(self) -> Packed:
        logp = self.logits_table()
        if self.n_classes > TABLE_MAX_CLASSES or any(v.numel() > TABLE_MAX_CLASSES for v in self._parent_values):
            raise ValueError(f"categorical_embedded_softmax supports at most {TABLE_MAX_CLASSES} classes per variable on the GPU")
        params = _tab_params(self.n_classes, logp, self._parent_values, self._parent_strides,
                             self._class_values[0], torch.where(self._class_mask[0], self._class_values[0],
                                                                torch.full_like(self._class_values[0], float("nan"))),
                             strict=True)
        return Packed(kind=L.OP_TAB, dim=1, n_par=self.input_dim, params=params, k=self.n_classes,
                      n_normals=0, n_uniforms=1, scratch=0, heavy=False)

    @classmethod
    def from_spec(cls, c, device=None):
        return cls(c["input_dim"], c["output_dim"], n_classes=c["n_classes"], class_values=c["class_values"],
                   class_mask=c["class_mask"], parent_values=c.get("parent_values", ()),
                   embeddings=c.get("embeddings"), layers=c.get("layers"), logits=c.get("logits"),
                   activation=c.get("activation", "relu"), stats_ready=c.get("stats_ready", True), device=device)

    @classmethod
    def from_reference(cls, cpd, device=None):
        root = cpd.input_dim == 0
        return cls(cpd.input_dim, cpd.output_dim, n_classes=cpd.n_classes, class_values=cpd._class_values,
                   class_mask=cpd._class_mask, parent_values=cpd._parent_values or (),
                   embeddings=None if root else [e.weight for e in cpd.embeddings],
                   layers=None if root else _layers_from_module(cpd.net),
                   logits=cpd._logits if root else None, activation=cpd.activation,
                   stats_ready=bool(cpd._stats_ready.item()), device=device)

    def to_spec(self):
        return {"kind": self.kind, "input_dim": self.input_dim, "output_dim": self.output_dim,
                "n_classes": self.n_classes, "activation": self.activation, "stats_ready": self._stats_ready,
                "class_values": self._class_values, "class_mask": self._class_mask,
                "sample_values": self._sample_values, "parent_values": self._parent_values,
                "logits": self._logits, "embeddings": self.embeddings, "layers": self.layers}


# ------------------------------------------------------------------------------------------
# discrete-parent lookup tables (VBN_OP_TAB)
# ------------------------------------------------------------------------------------------
_TABLE_CACHE: Dict[tuple, Packed] = {}
TABLE_MAX_CONFIGS = 4096


TABLE_MAX_CLASSES = 64
TABLE_KINDS = ("categorical_table", "categorical_embedded_softmax")  # CPDs whose parameters ARE a lookup table


def _is_plain_discrete(cpd) -> bool:
    if isinstance(cpd, CategoricalTableCPD):
        return cpd.output_dim == 1 and cpd.n_support <= TABLE_MAX_CLASSES
    return (isinstance(cpd, SoftmaxNNCPD) and cpd.output_dim == 1 and cpd._bins_ready
            and bool(cpd._is_discrete.all()) and cpd.n_classes <= TABLE_MAX_CLASSES)


def _support_of(cpd) -> torch.Tensor:
    """Values a plain discrete node can take (sorted)."""
    if isinstance(cpd, CategoricalTableCPD):
        return cpd._class_values[0][cpd._class_mask[0]]
    return cpd._class_values[0]


def _tab_params(c: int, logp: torch.Tensor, parent_supports: Sequence[torch.Tensor], strides: Sequence[int],
                sample_values: torch.Tensor, class_values: torch.Tensor, strict: bool) -> np.ndarray:
    """VBN_OP_TAB parameter block (layout: csrc/vbn_schedule.cuh op_tab)."""
    n_cfg = int(logp.shape[0])
    widest = max([c] + [int(v.numel()) for v in parent_supports])
    cpad = (widest + 7) & ~7
    cdf = torch.cumsum(torch.exp(logp.double()), dim=1).float()
    head = [np.array([c, n_cfg, cpad, 1.0 if strict else 0.0], np.float32)]
    for sup, stride in zip(parent_supports, strides):
        head.append(np.concatenate([np.array([int(sup.numel()), stride, 0, 0], np.float32), _padded(_np(sup), cpad)]))
    head += [_padded(_np(sample_values), cpad), _padded(_np(class_values), cpad), _np(cdf), _np(logp)]
    params = np.concatenate(head)
    params = _padded(params, _pad4(params.size))
    if c <= 4:
        # 128-bit rows for the plain table op (csrc op_tab_plain): cdf4 = {c_0 .. c_{C-2}, +inf .., total} and logp4
        # (class k in lane k), each [n_cfg][4], behind the block (plan._tab_plain_fields derives the offsets)
        cdf_np = cdf.detach().cpu().numpy().astype(np.float32).reshape(n_cfg, c)
        cdf4 = np.full((n_cfg, 4), np.inf, np.float32)
        cdf4[:, : c - 1] = cdf_np[:, : c - 1]
        cdf4[:, 3] = cdf_np[:, c - 1]
        logp4 = np.zeros((n_cfg, 4), np.float32)
        logp4[:, :c] = logp.detach().cpu().numpy().astype(np.float32).reshape(n_cfg, c)
        params = np.concatenate([params, cdf4.reshape(-1), logp4.reshape(-1)])
    return params


def table_eligible(cpd, parent_cpds: Sequence[BaseCPD]) -> bool:
    """A discrete-mode softmax_nn node (D = 1, <= 8 classes) whose parents are all such nodes: its
    logits take one value per parent configuration."""
    if not _is_plain_discrete(cpd) or not all(_is_plain_discrete(p) for p in parent_cpds):
        return False
    if len(parent_cpds) != cpd.input_dim:
        return False
    if not isinstance(cpd, SoftmaxNNCPD):
        return False  # categorical_table nodes pack themselves
    n_cfg = 1
    for pc in parent_cpds:
        n_cfg *= int(_support_of(pc).numel())
    return n_cfg <= TABLE_MAX_CONFIGS


def pack_table(cpd, parent_cpds: Sequence[BaseCPD], log_prob_fn) -> Packed:
    """Builds the VBN_OP_TAB block: the node's own log-density for every (parent configuration, class),
    evaluated by ``log_prob_fn(cpd, x[n,1], parents[n,Dp] | None) -> [n]`` -- the library's GPU path, so
    the table holds exactly the numbers a per-row evaluation would produce."""
    key = (cpd._uid, cpd._version, tuple((pc._uid, pc._version) for pc in parent_cpds))
    hit = _TABLE_CACHE.get(key)
    if hit is not None:
        return hit
    base = cpd.pack()
    c = cpd.n_classes
    supports = [_support_of(pc) for pc in parent_cpds]
    cards = [int(v.numel()) for v in supports]
    n_cfg = int(np.prod(cards)) if cards else 1
    strides, acc = [], 1
    for card in reversed(cards):
        strides.append(acc)
        acc *= card
    strides = strides[::-1]
    if cards:
        grids = torch.cartesian_prod(*supports).reshape(n_cfg, len(cards))
    else:
        grids = None
    cols = []
    for k in range(c):
        x = cpd._class_values[0, k].repeat(n_cfg).reshape(n_cfg, 1)
        cols.append(_f32(log_prob_fn(cpd, x, grids)).reshape(n_cfg))
    logp = torch.stack(cols, dim=1)                       # [n_cfg, C]
    params = _tab_params(c, logp, supports, strides, cpd._sample_values[0], cpd._class_values[0], strict=False)
    pk = Packed(kind=L.OP_TAB, dim=1, n_par=len(cards), params=params, k=c,
                n_normals=base.n_normals, n_uniforms=base.n_uniforms, scratch=0, heavy=False)
    if len(_TABLE_CACHE) > 4096:
        _TABLE_CACHE.clear()
    _TABLE_CACHE[key] = pk
    return pk


CPD_CLASSES = {
    "linear_gaussian": LinearGaussianCPD,
    "gaussian_nn": GaussianNNCPD,
    "mdn": MDNCPD,
    "softmax_nn": SoftmaxNNCPD,
    "kde": KDECPD,
    "categorical_table": CategoricalTableCPD,
    "rff_gaussian": RFFGaussianCPD,
    "categorical_embedded_softmax": CategoricalEmbeddedSoftmaxCPD,
}

_REFERENCE_CLASS_NAMES = {
    "LinearGaussianCPD": LinearGaussianCPD,
    "GaussianNNCPD": GaussianNNCPD,
    "MDNCPD": MDNCPD,
    "SoftmaxNNCPD": SoftmaxNNCPD,
    "KDECPD": KDECPD,
    "CategoricalTableCPD": CategoricalTableCPD,
    "RFFGaussianCPD": RFFGaussianCPD,
    "CategoricalEmbeddedSoftmaxCPD": CategoricalEmbeddedSoftmaxCPD,
}


def cpd_from_spec(c: dict, device=None) -> BaseCPD:
    kind = c["kind"]
    if kind not in CPD_CLASSES:
        raise ValueError(f"CPD kind '{kind}' is outside the accelerated path")
    return CPD_CLASSES[kind].from_spec(c, device)


def wrap_cpd(obj, device=None) -> BaseCPD:
    """Accepts one of our CPDs or a reference ``vbn.cpds.*`` module and returns ours."""
    if isinstance(obj, BaseCPD):
        return obj
    name = type(obj).__name__
    if name in _REFERENCE_CLASS_NAMES and hasattr(obj, "input_dim"):
        return _REFERENCE_CLASS_NAMES[name].from_reference(obj, device)
    raise TypeError(
        f"CPD type '{name}' has no CUDA implementation (supported: {sorted(CPD_CLASSES)}); "
        "there is no CPU fallback"
    )
