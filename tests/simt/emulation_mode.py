"""Opt-in dry run of the GPU test suite on the host emulation:   THB_SIMT_EMULATION=1 python -m pytest tests -m gpu
(tests/conftest.py calls enable()).  The CUDA library is replaced by tests/simt's emulation build (+ numpy stand-ins for the dense DMMA
Cholesky), the engine's CUDA guard and the pinned-memory statistics read are replaced, and "cuda" devices map to the CPU inside torch's
factory functions / .cuda() / .to().  A dry run of host code and kernel logic, not a parity gate: the numbers of the dense Cholesky
come from numpy here, and nothing about the device's arithmetic or memory model is exercised."""
import ctypes as C
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def np_at(ptr, shape, dtype=np.float64):
    n = int(np.prod(shape))
    addr = ptr if isinstance(ptr, int) else ptr.value
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


class EmulatedLib:
    """Every entry point the emulation library exports -> the emulated kernels; symbolic analysis (host C++) -> the real library; the dense
    Cholesky (DMMA, not emulated) -> numpy with the same contract (thb200.h: damping fused into the load, info = failing pivot)."""

    def __init__(self, emu, real, signatures):
        self._emu, self._real, self._sig = emu, real, signatures

    def __getattr__(self, name):
        if name in self._sig and hasattr(self._emu, name) and not name.startswith("thb_potr"):
            return getattr(self._emu, name)
        if name.startswith("thb_symbolic_"):
            return getattr(self._real, name)
        raise AttributeError(f"{name}: a CUDA entry point without an emulation or stand-in")

    @staticmethod
    def _damped(Aptr, alpha, beta, B, n):
        M = np_at(Aptr, (B, n, n)).copy()
        M = np.tril(M) + np.transpose(np.tril(M, -1), (0, 2, 1))
        idx = np.arange(n)
        if alpha is not None and getattr(alpha, "value", alpha) is not None:
            M[:, idx, idx] = M[:, idx, idx] * (1.0 + np_at(alpha, (B,))[:, None]) + np_at(beta, (B,))[:, None]
        return M

    def thb_potrf_workspace_bytes(self, B, n):
        return int(B) * int(n) * int(n) * 8 + 256

    def thb_potrf_f64(self, Aptr, alpha, beta, info, B, n, ws, ws_bytes, stream):
        M = self._damped(Aptr, alpha, beta, B, n)
        L, inf = np_at(ws, (B, n, n)), np_at(info, (B,), np.int32)
        for i in range(B):
            try:
                L[i] = np.linalg.cholesky(M[i]); inf[i] = 0
            except np.linalg.LinAlgError:
                inf[i] = 1
        return 0

    def thb_potrs_f64(self, rhs, x, B, n, ws, ws_bytes, stream):
        L, r, out = np_at(ws, (B, n, n)), np_at(rhs, (B, n)), np_at(x, (B, n))
        for i in range(B):
            out[i] = np.linalg.solve(L[i].T, np.linalg.solve(L[i], r[i]))
        return 0

    def thb_potrf_potrs_f64(self, Aptr, rhs, alpha, beta, x, info, B, n, ws, ws_bytes, stream):
        self.thb_potrf_f64(Aptr, alpha, beta, info, B, n, ws, ws_bytes, stream)
        if not np_at(info, (B,), np.int32).any():
            self.thb_potrs_f64(rhs, x, B, n, ws, ws_bytes, stream)
        return 0


def load_emulated_lib():
    from theseus_b200 import _lib
    real = _lib.load()
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(HERE, "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return EmulatedLib(lib, real, _lib.SIGNATURES)


def patch_host(setattr_fn, lib):
    """Replace the loader, the CUDA guards and the pinned-memory statistics read.  setattr_fn(obj, name, value): monkeypatch.setattr or setattr."""
    from theseus_b200 import _lib, engine as engine_mod, geometry as geometry_mod, optimizer as optimizer_mod
    setattr_fn(_lib, "load", lambda: lib)
    setattr_fn(_lib, "stream_ptr", lambda: None)
    setattr_fn(engine_mod, "_require_cuda_device", lambda device: None)
    setattr_fn(geometry_mod, "_require_cuda", lambda t, what: None)
    # the one host read per LM iteration goes through pinned memory and a stream synchronisation on the GPU
    def _read_stats(self, stats, B):
        if self.process_group is not None:      # same collective as the product's _read_stats (the sharded-loop test relies on it)
            from theseus_b200.distributed import reduce_counts
            stats[1] = B
            reduce_counts(stats, self.process_group)
            return int(stats[0]) == int(stats[1])
        return int(stats[0]) == B
    setattr_fn(optimizer_mod.LevenbergMarquardt, "_read_stats", _read_stats)


def _is_cuda(d):
    return d is not None and (str(d).startswith("cuda") or (isinstance(d, torch.device) and d.type == "cuda"))


def enable():
    """Global (process-wide) switch used by the opt-in dry run of the GPU tests."""
    patch_host(setattr, load_emulated_lib())
    torch.Tensor.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *args, **kwargs):
        args = tuple("cpu" if _is_cuda(a) and not isinstance(a, torch.dtype) else a for a in args)
        if _is_cuda(kwargs.get("device")):
            kwargs["device"] = "cpu"
        return _to(self, *args, **kwargs)
    torch.Tensor.to = to
    for fname in ("ones", "zeros", "empty", "full", "tensor", "randn", "rand", "arange", "eye", "as_tensor", "linspace", "ones_like", "zeros_like",
                  "empty_like", "full_like", "randint"):
        orig = getattr(torch, fname)

        def wrapped(*args, __orig=orig, **kwargs):
            if _is_cuda(kwargs.get("device")):
                kwargs["device"] = "cpu"
            return __orig(*args, **kwargs)
        setattr(torch, fname, wrapped)
    from theseus_b200 import core as core_mod
    _oto = core_mod.Objective.to

    def objective_to(self, *args, **kwargs):
        args = tuple("cpu" if _is_cuda(a) and not isinstance(a, torch.dtype) else a for a in args)
        if _is_cuda(kwargs.get("device")):
            kwargs["device"] = "cpu"
        return _oto(self, *args, **kwargs)
    core_mod.Objective.to = objective_to
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.is_available = lambda: False
