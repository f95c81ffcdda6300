"""The linear solve as ONE differentiable op of (A_val, b):  x = (AtA + D)^-1 At b.

Mirror of the reference's solver autograd functions -- BaspachoSolveFunction, CholmodSolveFunction, LUCudaSolveFunction
(theseus/optimizer/autograd/*_sparse_autograd.py) -- and of what torch autograd does for the dense solver there.  The
backward is the closed form derived in baspacho_sparse_autograd.py:68-115:
    H = (AtA + D)^-1 grad_x,   b_grad = A H,   A_grad = (b - A x) (x) H - A H (x) x - 2 alpha (H o x) A   on the pattern of A
computed by thb_solve_backward_f64 (one fused kernel; the reference loops over the m rows in Python, common.py:27-36).
H reuses the factor of the forward pass; if the solver has factorised another system in between (the case the
reference's num_solver_contexts pool exists for, lu_cuda_sparse_solver.py:40-77) the factor is rebuilt from the saved inputs.

A solver "backend" provides:  _solve_nograd(A_val, b, damping, ellipsoidal, eps) -> (x, (A64, b64, x64, alpha, beta)),
_numeric(A64, b64, alpha, beta), _substitute(rhs) -> x, _factor_stamp, linearization (A_row_ptr, A_col_ind, num_rows, num_cols).
"""
import numpy as np
import torch

from . import _lib


def solve_backward(structure_dev, A_val, b, x, H, alpha, detach_hessian: bool):
    """(A_grad [B,nnz], b_grad [B,m]) given H = (AtA + D)^-1 grad_x  (optimizer/autograd/common.py:11-48)."""
    row_ptr, col_ind, num_rows, num_cols = structure_dev
    B = A_val.shape[0]
    A_grad, b_grad = torch.empty_like(A_val), torch.empty_like(b)
    if alpha is not None and not bool((alpha > 0).any()):
        alpha = None
    if alpha is not None and detach_hessian:
        raise RuntimeError("detach_hessian is only meant for an undamped Gauss-Newton step (common.py:38-41)")
    _lib.check(_lib.load().thb_solve_backward_f64(B, num_rows, num_cols, _lib.ptr(row_ptr), _lib.ptr(col_ind), _lib.ptr(A_val), _lib.ptr(b),
                                                  _lib.ptr(x), _lib.ptr(H.contiguous()), _lib.ptr(alpha), int(detach_hessian),
                                                  _lib.ptr(A_grad), _lib.ptr(b_grad), _lib.stream_ptr()), "solve_backward")
    return A_grad, b_grad


def csr_on_device(lin, device):
    """(A_row_ptr, A_col_ind, m, n) with int64 device index tensors, cached on the linearization object."""
    c = getattr(lin, "_csr_dev", None)
    if c is None or c[0].device != device:
        S = lin if hasattr(lin, "A_row_ptr") else lin.engine.structure  # DenseLinearization: the CSR lives in the engine's structure
        c = (torch.from_numpy(np.ascontiguousarray(S.A_row_ptr, dtype=np.int64)).to(device),
             torch.from_numpy(np.ascontiguousarray(S.A_col_ind, dtype=np.int64)).to(device), int(lin.num_rows), int(lin.num_cols))
        lin._csr_dev = c
    return c


class LinearSolveFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A_val, b, solver, damping, ellipsoidal_damping, damping_eps, detach_hessian):
        x, (A64, b64, x64, alpha, beta) = solver._solve_nograd(A_val, b, damping, ellipsoidal_damping, damping_eps)
        ctx.solver, ctx.stamp, ctx.detach = solver, solver._factor_stamp, detach_hessian
        # the LM control kernel updates the damping tensor IN PLACE after every step: keep this iteration's values
        alpha, beta = (alpha.clone(), beta.clone()) if alpha is not None else (None, None)
        ctx.saved = (A64, b64, x64, alpha, beta)
        ctx.in_dtypes = (A_val.dtype, b.dtype)
        return x

    @staticmethod
    def backward(ctx, grad_x):
        solver = ctx.solver
        A64, b64, x64, alpha, beta = ctx.saved
        if solver._factor_stamp != ctx.stamp:
            solver._numeric(A64, b64, alpha, beta)
            ctx.stamp = solver._factor_stamp
        H = solver._substitute(grad_x.double())
        A_grad, b_grad = solve_backward(csr_on_device(solver.linearization, A64.device), A64, b64, x64, H, alpha, ctx.detach)
        return A_grad.to(ctx.in_dtypes[0]), b_grad.to(ctx.in_dtypes[1]), None, None, None, None, None


def wants_grad(A_val, b) -> bool:
    return torch.is_grad_enabled() and (A_val.requires_grad or b.requires_grad)
