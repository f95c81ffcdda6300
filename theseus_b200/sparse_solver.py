"""BaspachoSparseSolver (alias BlockSparseSolver): batched block-sparse Cholesky on the GPU.

Mirror of theseus/optimizer/linear/baspacho_sparse_solver.py:23-160 + BaspachoSolveFunction.forward
(optimizer/autograd/baspacho_sparse_autograd.py:21-65): reset() builds the block structure of AtA and the symbolic
decomposition once; solve() = add_MtM (thb_gram into the factor storage) -> damp(alpha, beta) -> factor -> Atb -> solve,
all in fp64 like the reference (baspacho_sparse_autograd.py:41,65), result cast back to the objective dtype.
It is also what CholmodSparseSolver / LUCudaSparseSolver map to (same linear system).
"""
import ctypes as C
from typing import Any, Dict, Optional, Type, Union

import numpy as np
import torch

from . import _lib
from .core import Objective
from .optimizer import (Linearization, LinearSolver, SparseLinearization, convert_to_alpha_beta_damping_tensors)
from .sparse import analyze, gram_out_offsets
from .structure import ata_block_structure, build_gram_plan


class BaspachoSparseSolver(LinearSolver):
    def __init__(self, objective: Objective, linearization_cls: Optional[Type[Linearization]] = None,
                 linearization_kwargs: Optional[Dict[str, Any]] = None, num_solver_contexts=1, batch_size: Optional[int] = None,
                 auto_reset: bool = True, dev: Optional[str] = None, ordering: str = "mindeg", **kwargs):
        linearization_cls = linearization_cls or SparseLinearization
        if not linearization_cls == SparseLinearization:
            raise RuntimeError(
                "BaspachoSparseSolver only works with theseus.optimizer.SparseLinearization,"
                + f" got {type(self.linearization)}")
        super().__init__(objective, linearization_cls, linearization_kwargs, **kwargs)
        self.linearization: SparseLinearization = self.linearization
        self._ordering = ordering
        self._plan = None
        self._dev = None
        self.reset()

    @classmethod
    def from_structure(cls, structure, ordering: str = "mindeg") -> "BaspachoSparseSolver":
        """Solver over a bare CSR structure with hand-filled `linearization.A_val / b` -- the pattern of the reference's own
        solver tests (void Objective + filled linearization, tests/theseus_tests/optimizer/linear/test_baspacho_sparse_solver.py:15-41)."""
        class _Lin:
            def __init__(self, S):
                self._S, self.A_val, self.b = S, None, None
                self.A_row_ptr, self.A_col_ind = S.A_row_ptr, S.A_col_ind
                self.num_rows, self.num_cols = S.num_rows, S.num_cols

            def structure(self):
                return self._S
        self = cls.__new__(cls)
        self.linearization = _Lin(structure)
        self._ordering, self._plan, self._dev = ordering, None, None
        self.reset()
        return self

    # ---- symbolic phase (baspacho_sparse_solver.py:58-113) ----
    def reset(self, **kwargs):
        if self._plan is not None:
            return
        S = self.linearization.structure()
        param_size, ptrs, inds = ata_block_structure(S)
        self.param_size, self.block_ptrs, self.block_inds = param_size, ptrs, inds  # the reference's SymbolicDecomposition inputs
        self._plan = analyze(param_size, ptrs, inds, ordering=self._ordering)
        self._gram_arrays = build_gram_plan(S, out_offsets=gram_out_offsets(self._plan), pos=self._plan.pos)

    @property
    def symbolic_stats(self):
        return dict(self._plan.stats)

    def _device_plan(self, device):
        if self._dev is not None and self._dev["device"] == device:
            return self._dev
        P = self._plan
        dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in P.arrays.items()}
        st = _lib.SparsePlanStruct(N=P.N, num_levels=int(P.stats["levels"]), max_dim=int(P.dims.max()) if P.N else 0, reserved=0,
                                   n=P.n, data_size=P.data_size, winv_size=P.winv_size,
                                   **{k: dev[k].data_ptr() for k in dev})
        g = self._gram_arrays
        gdev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in g.items() if isinstance(v, np.ndarray)}
        gst = _lib.GramPlan(
            num_entries=int(g["ent_blk"].shape[0]), ent_blk=gdev["ent_blk"].data_ptr(), ent_p=gdev["ent_p"].data_ptr(),
            ent_q=gdev["ent_q"].data_ptr(), blk_out=gdev["blk_out"].data_ptr(), blk_ld=gdev["blk_ld"].data_ptr(),
            blk_mirror=gdev["blk_mirror"].data_ptr(), blk_cptr=gdev["blk_cptr"].data_ptr(), c_off=gdev["c_off"].data_ptr(),
            c_stride=gdev["c_stride"].data_ptr(), c_rows=gdev["c_rows"].data_ptr(), c_bpa=gdev["c_bpa"].data_ptr(),
            c_bpb=gdev["c_bpb"].data_ptr(), n=int(g["n"]), col_cptr=gdev["col_cptr"].data_ptr(), cc_off=gdev["cc_off"].data_ptr(),
            cc_stride=gdev["cc_stride"].data_ptr(), cc_rows=gdev["cc_rows"].data_ptr(), cc_row0=gdev["cc_row0"].data_ptr())
        self._dev = dict(device=device, plan=st, keep=dev, gram=gst, gkeep=gdev, bufs={})
        return self._dev

    # ---- numeric phase (baspacho_sparse_autograd.py:21-65) ----
    def solve(self, damping: Optional[Union[float, torch.Tensor]] = None, ellipsoidal_damping: bool = True,
              damping_eps: float = 1e-8, **kwargs) -> torch.Tensor:
        lin = self.linearization
        A_val, b = lin.A_val, lin.b
        if A_val is None:
            raise RuntimeError("linearize() must be called before solve()")
        out_dtype = A_val.dtype
        if A_val.dtype != torch.float64:
            A_val, b = A_val.double(), b.double()
        B = A_val.shape[0]
        device = A_val.device
        d = self._device_plan(device)
        P = self._plan
        lib = _lib.load()
        s = _lib.stream_ptr()
        key = (B,)
        if d["bufs"].get("key") != key:
            d["bufs"] = dict(key=key,
                             factor=torch.empty(B, P.data_size, dtype=torch.float64, device=device),
                             winv=torch.empty(B, P.winv_size, dtype=torch.float64, device=device),
                             Atb=torch.empty(B, P.n, dtype=torch.float64, device=device),
                             work=torch.empty(B, P.n, dtype=torch.float64, device=device),
                             info=torch.empty(B, dtype=torch.int32, device=device))
        bufs = d["bufs"]
        factor, winv, Atb, work, info = bufs["factor"], bufs["winv"], bufs["Atb"], bufs["work"], bufs["info"]
        nnz, m = A_val.shape[1], b.shape[1]
        # add_MtM + tmat_vec: every structurally non-zero block of L's pattern that is not in AtA must start at zero
        _lib.check(lib.thb_fill_zero(_lib.ptr(factor), factor.numel() * 8, s), "fill_zero")
        _lib.check(lib.thb_gram_f64(C.byref(d["gram"]), B, _lib.ptr(A_val), nnz, _lib.ptr(b), m, _lib.ptr(factor), P.data_size,
                                    _lib.ptr(Atb), None, s), "gram(sparse)")
        if damping is not None:
            alpha, beta = convert_to_alpha_beta_damping_tensors(damping, damping_eps, ellipsoidal_damping, B, device, torch.float64)
            _lib.check(lib.thb_sparse_damp_f64(C.byref(d["plan"]), _lib.ptr(factor), _lib.ptr(alpha), _lib.ptr(beta), B, s), "sparse_damp")
            self._keep = (alpha, beta)
        _lib.check(lib.thb_sparse_factor_f64(C.byref(d["plan"]), _lib.ptr(factor), _lib.ptr(winv), _lib.ptr(info), B, s), "sparse_factor")
        x = torch.empty(B, P.n, dtype=torch.float64, device=device)
        _lib.check(lib.thb_sparse_solve_f64(C.byref(d["plan"]), _lib.ptr(factor), _lib.ptr(winv), _lib.ptr(Atb), _lib.ptr(x), _lib.ptr(work), B, s),
                   "sparse_solve")
        self._last = (A_val, b)
        bad = info.nonzero()
        if bad.numel() > 0:
            k = int(bad[0, 0])
            raise RuntimeError(f"block-sparse Cholesky: batch element {k}: matrix is not positive definite (pivot {int(info[k])})")
        return x.to(out_dtype)


BlockSparseSolver = BaspachoSparseSolver
CholmodSparseSolver = BaspachoSparseSolver   # same linear system; the CPU per-item CHOLMOD loop has no place on the GPU path
LUCudaSparseSolver = BaspachoSparseSolver    # AtA + damping is SPD: the LU variant is served by the same Cholesky
