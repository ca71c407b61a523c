"""Two ways to run the product's device code in tests:
  * "cuda": the real thing -- libvbn_cuda.so on a B200 (tests marked ``gpu``);
  * "emu" : TEST-ONLY host emulation (tests/emu): the schedule kernel's device source compiled by
            g++ and driven through the same Python host code on CPU tensors, so the kernel logic
            and the host plumbing are checked against the oracle / golden vectors in a container
            without a GPU.  It is not part of the package and is never used outside tests.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libvbn_emu.so")


def build_emu() -> str:
    srcs = [os.path.join(EMU_DIR, "vbn_emu.cpp"), os.path.join(EMU_DIR, "cuda_shim.h"),
            os.path.join(ROOT, "include", "vbn_cuda.h")]
    csrc = os.path.join(ROOT, "vectorizedbayesiannetwork_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_LIB) for s in srcs):
        return EMU_LIB
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", csrc,
           "-I", EMU_DIR, os.path.join(EMU_DIR, "vbn_emu.cpp"), "-o", EMU_LIB]
    subprocess.run(cmd, check=True, capture_output=True)
    return EMU_LIB


class Backend:
    def __init__(self, name: str, device: torch.device):
        self.name = name
        self.device = device


@contextlib.contextmanager
def _null_device(_dev=None):
    yield


def _activate_emu(monkeypatch) -> Backend:
    from vectorizedbayesiannetwork_b200 import _lib as L
    from vectorizedbayesiannetwork_b200 import engine as E

    lib = C.CDLL(build_emu())
    for name, (res, args) in L.EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    monkeypatch.setattr(L, "_lib", lib)
    monkeypatch.setattr(E, "require_cuda", lambda device=None: torch.device("cpu"))
    monkeypatch.setattr(E, "_stream_ptr", lambda dev: 0)
    monkeypatch.setattr(torch.cuda, "device", _null_device)
    E._CPD_PLANS.clear()
    return Backend("emu", torch.device("cpu"))


def _activate_cuda() -> Backend:
    from vectorizedbayesiannetwork_b200 import _lib as L
    from vectorizedbayesiannetwork_b200 import engine as E

    assert torch.cuda.is_available(), "gpu test on a box without CUDA"
    L._lib = None
    L.load()
    E._CPD_PLANS.clear()
    return Backend("cuda", torch.device("cuda", 0))


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "emu":
        yield _activate_emu(monkeypatch)
        from vectorizedbayesiannetwork_b200 import engine as E

        E._CPD_PLANS.clear()
    else:
        yield _activate_cuda()
