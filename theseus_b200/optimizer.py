"""Linearization / LinearSolver / NonlinearLeastSquares (GaussNewton, LevenbergMarquardt): the host-side
mirror of theseus/optimizer/{linearization,dense_linearization,sparse_linearization}.py,
theseus/optimizer/linear/{linear_solver,dense_solver}.py and theseus/optimizer/nonlinear/*.py.

Same class names, constructor / solve / optimize kwargs, info fields and error behaviour as the reference
(a non-PD system surfaces as RuntimeError, caught by the loop under no_grad -> status FAIL,
nonlinear_least_squares.py:138-152), so the parity tests read like the reference's own tests.  The
arithmetic is the CUDA library; the Python below only sequences kernel launches.
"""
import ctypes as C
import math
import warnings
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Optional, Type, Union

import numpy as np
import torch

from . import _lib
from .core import Objective
from .structure import ata_block_structure


# ------------------------------------------------------------------------------------------------ linearization
class VariableOrdering:
    """theseus/optimizer/variable_ordering.py:11-60 (default order = first appearance)."""

    def __init__(self, objective: Objective, default_order: bool = True):
        self.objective = objective
        self._var_order = list(objective.optim_vars.values()) if default_order else []
        self._var_name_to_index = {v.name: i for i, v in enumerate(self._var_order)}

    def index_of(self, key: str) -> int:
        return self._var_name_to_index[key]

    def append(self, var):
        """variable_ordering.py:38-48."""
        if var.name in self._var_name_to_index:
            raise ValueError(f"Variable {var.name} has already been added to the order.")
        if var.name not in self.objective.optim_vars:
            raise ValueError(f"Variable {var.name} is not an optimization variable for the objective.")
        self._var_order.append(var)
        self._var_name_to_index[var.name] = len(self._var_order) - 1

    def remove(self, var):
        self._var_order.remove(var)
        del self._var_name_to_index[var.name]
        self._var_name_to_index = {v.name: i for i, v in enumerate(self._var_order)}

    def extend(self, variables):
        for var in variables:
            self.append(var)

    def __getitem__(self, index):
        return self._var_order[index]

    def __iter__(self):
        return iter(self._var_order)

    def __len__(self):
        return len(self._var_order)

    @property
    def complete(self):
        return len(self._var_order) == self.objective.size_variables()


class Linearization:
    """theseus/optimizer/linearization.py:16-87."""

    def __init__(self, objective: Objective, ordering: Optional[VariableOrdering] = None, **kwargs):
        self.objective = objective
        self.ordering = ordering or VariableOrdering(objective)
        if not self.ordering.complete:
            raise ValueError("Given variable ordering is not complete.")
        # the objective's engine (pointer tables, CSR structure, column layout of delta) is compiled for ONE variable order: a custom
        # order is recorded on the objective here and the engine is (re)built for it on first use (core.Objective.engine)
        self._ordering_names = tuple(v.name for v in self.ordering)
        objective._engine_ordering = None if list(self._ordering_names) == list(objective.optim_vars.keys()) else self._ordering_names
        self.var_dims = [v.dof() for v in self.ordering]
        self.var_start_cols = list(np.concatenate([[0], np.cumsum(self.var_dims)[:-1]]).astype(int)) if self.var_dims else []
        self.num_cols = int(sum(self.var_dims))
        self.num_rows = objective.dim()

    @property
    def engine(self):
        return self.objective.engine(self._ordering_names)

    def linearize(self, _detach_hessian: bool = False, differentiable: bool = False):
        """differentiable=True (backward modes): the Jacobian values / residuals come out as autograd tensors built from the
        cost functions' torch.func Jacobians (core.AutoDiffCostFunction), not from the fused kernels."""
        if not self.ordering.complete:
            raise RuntimeError("Attempted to linearize an objective with an incomplete variable order.")
        self._differentiable = differentiable
        self._linearize_hessian_impl(_detach_hessian=_detach_hessian)

    # linearization.py:62-87: what a subclass provides
    def _linearize_jacobian_impl(self):
        raise NotImplementedError

    def _linearize_hessian_impl(self, _detach_hessian: bool = False):
        raise NotImplementedError

    def _ata_impl(self) -> torch.Tensor:
        raise NotImplementedError

    def _atb_impl(self) -> torch.Tensor:
        raise NotImplementedError

    def Av(self, v: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def diagonal_scaling(self, v: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def hessian_approx(self):
        return self.AtA

    @property
    def AtA(self) -> torch.Tensor:
        return self._ata_impl()

    @property
    def Atb(self) -> torch.Tensor:
        return self._atb_impl()


class SparseLinearization(Linearization):
    """theseus/optimizer/sparse_linearization.py:19-198: batch-shared CSR pattern, A_val [B,nnz], b [B,m]."""

    def __init__(self, objective: Objective, ordering: Optional[VariableOrdering] = None, **kwargs):
        super().__init__(objective, ordering)
        from .structure import build_structure
        idx = {v.name: i for i, v in enumerate(self.ordering)}
        S = build_structure(self.var_dims, [(cf.dim(), [idx[v.name] for v in cf.optim_vars]) for cf in objective.cost_functions.values()])
        self.A_row_ptr, self.A_col_ind = S.A_row_ptr, S.A_col_ind
        self.cost_function_block_pointers = S.block_pointers
        self.cost_function_row_block_starts = S.row_block_starts
        self.cost_function_stride = S.stride
        self._structure = S
        self.A_val: torch.Tensor = None
        self.b: torch.Tensor = None
        self._Atb = None
        self._AtA_diag = None
        self.detached_hessian = False

    def _linearize_jacobian_impl(self):
        self._Atb = None
        self._AtA_diag = None
        if getattr(self, "_differentiable", False):
            self.A_val, self.b = self.engine.linearize_sparse_differentiable()
        else:
            self.A_val, self.b = self.engine.linearize_sparse()

    def _linearize_hessian_impl(self, _detach_hessian: bool = False):
        self._linearize_jacobian_impl()
        self.detached_hessian = _detach_hessian

    def _ata_impl(self):
        raise NotImplementedError("AtA is not yet implemented for SparseLinearization.")  # same as the reference

    def _compute_atb_diag(self):
        eng = self.engine
        B = eng.batch_size
        Atb = eng.buf("sp_Atb", (B, self.num_cols))
        diag = eng.buf("sp_diag", (B, self.num_cols))
        eng.atb(self.A_val, self.b, Atb, diag)
        self._Atb, self._AtA_diag = Atb, diag

    def _atb_impl(self) -> torch.Tensor:
        if self._Atb is None:
            self._compute_atb_diag()
        return self._Atb.unsqueeze(2)

    def diagonal_scaling(self, v: torch.Tensor) -> torch.Tensor:
        if self._AtA_diag is None:
            self._compute_atb_diag()
        return self._AtA_diag * v

    def Av(self, v: torch.Tensor) -> torch.Tensor:
        eng = self.engine
        rp = eng.buf_const("A_row_ptr", self.A_row_ptr)
        ci = eng.buf_const("A_col_ind", self.A_col_ind)
        out = torch.empty(v.shape[0], self.num_rows, dtype=v.dtype, device=v.device)
        _lib.check(eng.lib.thb_mat_vec_f64(v.shape[0], self.num_rows, self.num_cols, _lib.ptr(rp), _lib.ptr(ci),
                                            _lib.ptr(self.A_val), _lib.ptr(v.contiguous()), _lib.ptr(out), _lib.stream_ptr()), "mat_vec")
        return out

    def structure(self):
        return self._structure


class DenseLinearization(Linearization):
    """theseus/optimizer/dense_linearization.py:16-80.  AtA [B,n,n], Atb [B,n,1]; A [B,m,n] is materialised
    only on request (property `A`), the hot path never forms the 99%-zero dense Jacobian."""

    def __init__(self, objective: Objective, ordering: Optional[VariableOrdering] = None, **kwargs):
        super().__init__(objective, ordering)
        self.b: torch.Tensor = None
        self._A_val = None
        self._AtA = None
        self._Atb = None
        self._diag = None

    def _linearize_jacobian_impl(self):
        if getattr(self, "_differentiable", False):
            self._A_val, self.b = self.engine.linearize_sparse_differentiable()
        else:
            self._A_val, self.b = self.engine.linearize_sparse()

    def _linearize_hessian_impl(self, _detach_hessian: bool = False):
        self._linearize_jacobian_impl()
        self.detached_hessian = _detach_hessian
        eng = self.engine
        B, n = eng.batch_size, self.num_cols
        self._AtA = eng.buf("AtA", (B, n, n))
        self._Atb = eng.buf("Atb", (B, n))
        self._diag = eng.buf("AtA_diag", (B, n))
        eng.gram_dense(self._A_val, self.b, self._AtA, self._Atb, self._diag)

    @property
    def A(self) -> torch.Tensor:
        """Dense Jacobian [B,m,n] scattered from the CSR values (debug / parity only)."""
        S = self.engine.structure
        B = self._A_val.shape[0]
        A = torch.zeros(B, self.num_rows, self.num_cols, dtype=self._A_val.dtype, device=self._A_val.device)
        rows = torch.from_numpy(np.repeat(np.arange(S.num_rows), np.diff(S.A_row_ptr))).to(A.device)
        cols = torch.from_numpy(S.A_col_ind).to(A.device)
        A[:, rows, cols] = self._A_val
        return A

    def hessian_approx(self):
        return self._AtA

    def _ata_impl(self) -> torch.Tensor:
        return self._AtA

    def _atb_impl(self) -> torch.Tensor:
        return self._Atb.unsqueeze(2)

    def Av(self, v: torch.Tensor) -> torch.Tensor:
        """A v from the CSR values (dense_linearization.py:73-74 does A.bmm(v) on the 99 %-zero dense Jacobian)."""
        eng = self.engine
        S = eng.structure
        rp, ci = eng.buf_const("A_row_ptr", S.A_row_ptr), eng.buf_const("A_col_ind", S.A_col_ind)
        A64, v64 = self._A_val.double().contiguous(), v.double().contiguous()
        out = torch.empty(v.shape[0], self.num_rows, dtype=torch.float64, device=v.device)
        _lib.check(eng.lib.thb_mat_vec_f64(v.shape[0], self.num_rows, self.num_cols, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(A64), _lib.ptr(v64),
                                            _lib.ptr(out), _lib.stream_ptr()), "mat_vec")
        return out.to(v.dtype)

    def diagonal_scaling(self, v: torch.Tensor) -> torch.Tensor:
        return v * self._diag


# ------------------------------------------------------------------------------------------------ linear solvers
def convert_to_alpha_beta_damping_tensors(damping, damping_eps: float, ellipsoidal_damping: bool, batch_size: int, device, dtype):
    """theseus/optimizer/linear/utils.py:14-33."""
    damping = torch.as_tensor(damping).to(device=device, dtype=dtype)
    if damping.ndim > 1:
        raise ValueError("Damping must be a float or a 1-D tensor.")
    if damping.ndim == 0 or damping.shape[0] == 1 and batch_size != 1:
        damping = damping.repeat(batch_size)
    return (damping, damping_eps * torch.ones_like(damping)) if ellipsoidal_damping else (torch.zeros_like(damping), damping)


class LinearSolver:
    """theseus/optimizer/linear/linear_solver.py:15-37."""

    def __init__(self, objective: Objective, linearization_cls: Optional[Type[Linearization]] = None,
                 linearization_kwargs: Optional[Dict[str, Any]] = None, **kwargs):
        linearization_kwargs = linearization_kwargs or {}
        self.linearization: Linearization = linearization_cls(objective, **linearization_kwargs)

    def reset(self, **kwargs):
        pass

    def solve(self, damping=None, **kwargs) -> torch.Tensor:
        raise NotImplementedError


class DenseSolver(LinearSolver):
    """theseus/optimizer/linear/dense_solver.py:19-123."""

    def __init__(self, objective: Objective, linearization_cls: Optional[Type[Linearization]] = None,
                 linearization_kwargs: Optional[Dict[str, Any]] = None, check_singular: bool = False):
        linearization_cls = linearization_cls or DenseLinearization
        if linearization_cls != DenseLinearization:
            raise RuntimeError(
                "DenseSolver only works with theseus.nonlinear.DenseLinearization, "
                f"but {linearization_cls} was provided.")
        super().__init__(objective, linearization_cls, linearization_kwargs)
        self._check_singular = check_singular
        self._ws = None

    def solve(self, damping: Optional[Union[float, torch.Tensor]] = None, ellipsoidal_damping: bool = True,
              damping_eps: float = 1e-8, **kwargs) -> torch.Tensor:
        lin = self.linearization
        A_val, b = getattr(lin, "_A_val", None), getattr(lin, "b", None)
        if A_val is not None and b is not None:
            from .autograd import LinearSolveFunction, wants_grad
            if wants_grad(A_val, b):  # differentiable solve: x as a function of the Jacobian values and residuals
                return LinearSolveFunction.apply(A_val, b, self, damping, ellipsoidal_damping, damping_eps,
                                                 bool(getattr(lin, "detached_hessian", False)))
        return self._apply_damping_and_solve(lin.Atb, lin.AtA, damping=damping, ellipsoidal_damping=ellipsoidal_damping,
                                             damping_eps=damping_eps)

    # ---- backend of autograd.LinearSolveFunction: the same system rebuilt from (A_val, b) in fp64 ----
    def _solve_nograd(self, A_val, b, damping, ellipsoidal_damping, damping_eps):
        out_dtype = A_val.dtype
        A64, b64 = A_val.detach().double().contiguous(), b.detach().double().contiguous()
        alpha = beta = None
        if damping is not None:
            alpha, beta = convert_to_alpha_beta_damping_tensors(damping, damping_eps, ellipsoidal_damping, A64.shape[0], A64.device, torch.float64)
        Atb = self._numeric(A64, b64, alpha, beta)
        x = self._substitute(Atb)
        return x.to(out_dtype), (A64, b64, x, alpha, beta)

    def _numeric(self, A_val, b, alpha, beta):
        eng = self.linearization.engine
        lib = _lib.load()
        B, n = A_val.shape[0], self.linearization.num_cols
        s = _lib.stream_ptr()
        st = getattr(self, "_bw", None)
        if st is None or st["AtA"].shape[0] != B or st["AtA"].device != A_val.device:
            st = dict(AtA=torch.zeros(B, n, n, dtype=torch.float64, device=A_val.device),
                      Atb=torch.empty(B, n, dtype=torch.float64, device=A_val.device),
                      ws=torch.empty(int(lib.thb_potrf_workspace_bytes(B, n)), dtype=torch.uint8, device=A_val.device),
                      info=torch.empty(B, dtype=torch.int32, device=A_val.device))
            self._bw = st
        plan = eng.gram_plan_dense()
        _lib.check(lib.thb_gram_f64(C.byref(plan), B, _lib.ptr(A_val), A_val.shape[1], _lib.ptr(b), b.shape[1], _lib.ptr(st["AtA"]), n * n,
                                    _lib.ptr(st["Atb"]), None, s), "gram")
        _lib.check(lib.thb_potrf_f64(_lib.ptr(st["AtA"]), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(st["info"]), B, n, _lib.ptr(st["ws"]),
                                     st["ws"].numel(), s), "potrf")
        self._factor_stamp = getattr(self, "_factor_stamp", 0) + 1
        self._keep_bw = (A_val, b, alpha, beta)
        bad = st["info"].nonzero()
        if bad.numel() > 0:
            k = int(bad[0, 0])
            raise torch.linalg.LinAlgError(f"linalg.cholesky: (Batch element {k}): The factorization could not be completed because the "
                                           f"input is not positive-definite (the leading minor of order {int(st['info'][k])} is not positive-definite).")
        return st["Atb"]

    def _substitute(self, rhs: torch.Tensor) -> torch.Tensor:
        st = self._bw
        B, n = rhs.shape
        rhs = rhs.contiguous()
        x = torch.empty(B, n, dtype=torch.float64, device=rhs.device)
        _lib.check(_lib.load().thb_potrs_f64(_lib.ptr(rhs), _lib.ptr(x), B, n, _lib.ptr(st["ws"]), st["ws"].numel(), _lib.stream_ptr()), "potrs")
        return x

    def _apply_damping_and_solve(self, Atb: torch.Tensor, AtA: torch.Tensor, damping=None, ellipsoidal_damping: bool = True,
                                 damping_eps: float = 1e-8) -> torch.Tensor:
        """Damping is fused into the factorisation's load of AtA (no second [B,n,n] tensor, dense_solver.py:38-78)."""
        if AtA.ndim != 3 or AtA.shape[1] != AtA.shape[2]:
            raise ValueError("Matrix must have a 3 dimensions, the first one being a batch dimension, and be square.")
        B, n, _ = AtA.shape
        out_dtype = AtA.dtype
        if AtA.dtype != torch.float64:
            AtA, Atb = AtA.double(), Atb.double()  # the fp64 kernel also serves fp32 objectives (result cast back)
        alpha = beta = None
        if damping is not None:
            alpha, beta = convert_to_alpha_beta_damping_tensors(damping, damping_eps, ellipsoidal_damping, B, AtA.device, AtA.dtype)
        lib = _lib.load()
        need = int(lib.thb_potrf_workspace_bytes(B, n))
        if self._ws is None or self._ws.numel() < need or self._ws.device != AtA.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=AtA.device)
        x = torch.empty(B, n, dtype=torch.float64, device=AtA.device)
        info = torch.empty(B, dtype=torch.int32, device=AtA.device)
        AtA_c, rhs = AtA.contiguous(), Atb.reshape(B, n).contiguous()
        _lib.check(lib.thb_potrf_potrs_f64(_lib.ptr(AtA_c), _lib.ptr(rhs), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(x), _lib.ptr(info),
                                           B, n, _lib.ptr(self._ws), self._ws.numel(), _lib.stream_ptr()), "potrf_potrs")
        self._last = (AtA_c, rhs, alpha, beta)
        self._last_info = info
        if getattr(self, "defer_info_check", False):  # CUDA-graph capture: no host sync here, check_info() is called after the replay
            return x.to(out_dtype)
        return self.check_info(x.to(out_dtype))

    def check_info(self, x=None):
        info = self._last_info
        bad = info.nonzero()
        if bad.numel() > 0:  # same visible behaviour as torch.linalg.cholesky: a RuntimeError subclass
            k = int(bad[0, 0])
            raise torch.linalg.LinAlgError(
                f"linalg.cholesky: (Batch element {k}): The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {int(info[k])} is not positive-definite).")
        return x


class CholeskyDenseSolver(DenseSolver):
    """theseus/optimizer/linear/dense_solver.py:144-161."""

    def __init__(self, objective: Objective, linearization_cls: Optional[Type[Linearization]] = DenseLinearization,
                 linearization_kwargs: Optional[Dict[str, Any]] = None, check_singular: bool = False):
        super().__init__(objective, linearization_cls, linearization_kwargs, check_singular=check_singular)


class LUDenseSolver(CholeskyDenseSolver):
    """theseus/optimizer/linear/dense_solver.py:125-141.  AtA (+damping) is SPD, so the LU variant is served by the
    same Cholesky kernel."""


# ------------------------------------------------------------------------------------------------ nonlinear
class BackwardMode(Enum):
    UNROLL = 0
    IMPLICIT = 1
    TRUNCATED = 2
    DLM = 3

    @staticmethod
    def resolve(key) -> "BackwardMode":
        if isinstance(key, BackwardMode):
            return key
        if not isinstance(key, str):
            raise ValueError("Backward mode must be th.BackwardMode or string.")
        try:
            return BackwardMode[key.upper()]
        except KeyError:
            raise ValueError(f"Unrecognized backward mode f{key}.Valid choices are unroll, implicit, truncated, dlm.")


class NonlinearOptimizerStatus(Enum):
    START = 0
    CONVERGED = 1
    MAX_ITERATIONS = 2
    FAIL = -1


@dataclass
class NonlinearOptimizerParams:
    abs_err_tolerance: float
    rel_err_tolerance: float
    max_iterations: int
    step_size: float

    def update(self, params_dict):
        for param, value in params_dict.items():
            if hasattr(self, param):
                setattr(self, param, value)
            else:
                raise ValueError(f"Invalid nonlinear optimizer parameter {param}.")


@dataclass
class OptimizerInfo:
    best_solution: Optional[Dict[str, torch.Tensor]]
    status: np.ndarray


@dataclass
class NonlinearOptimizerInfo(OptimizerInfo):
    converged_iter: torch.Tensor
    best_iter: torch.Tensor
    err_history: Optional[torch.Tensor]
    last_err: torch.Tensor
    best_err: torch.Tensor
    state_history: Optional[Dict[str, torch.Tensor]]


class NonlinearLeastSquares:
    """theseus/optimizer/nonlinear/nonlinear_least_squares.py:57-396 + nonlinear_optimizer.py:87-294."""
    _MAX_ALL_REJECT_ATTEMPTS = 3

    def __init__(self, objective: Objective, *args, linear_solver_cls: Optional[Type[LinearSolver]] = None, vectorize: bool = False,
                 linearization_cls: Optional[Type[Linearization]] = None, linearization_kwargs: Optional[Dict[str, Any]] = None,
                 linear_solver_kwargs: Optional[Dict[str, Any]] = None, abs_err_tolerance: float = 1e-10,
                 rel_err_tolerance: float = 1e-8, max_iterations: int = 20, step_size: float = 1.0, process_group=None,
                 cuda_graph: bool = False, **kwargs):
        self.objective = objective
        # cuda_graph=True: the iteration body (linearize -> solve -> retract -> error -> accept/reject -> commit) is captured once
        # into a CUDA graph and replayed; per iteration the host only reads 16 bytes of statistics (no reference analogue).
        self.cuda_graph = cuda_graph
        self._graph = None
        # Batch sharding over GPUs (no reference analogue): every rank owns a slice of the batch; the only
        # batch-global decisions of the loop (all-rejected retry, all-converged exit, mean-error test) are all-reduced.
        self.process_group = process_group
        self.params = NonlinearOptimizerParams(abs_err_tolerance, rel_err_tolerance, max_iterations, step_size)
        linear_solver_cls = linear_solver_cls or CholeskyDenseSolver
        linear_solver_kwargs = linear_solver_kwargs or {}
        self.linear_solver = linear_solver_cls(objective, linearization_cls=linearization_cls,
                                               linearization_kwargs=linearization_kwargs, **linear_solver_kwargs)
        self.ordering = self.linear_solver.linearization.ordering
        self._tmp_optim_vars = tuple(v.copy(new_name=v.name) for v in self.ordering)
        self._objectives_version = objective._structure_version

    def set_params(self, **kwargs):
        self.params.update(kwargs)

    # ---- public entry (optimizer/optimizer.py:40-53) ----
    def optimize(self, **kwargs) -> OptimizerInfo:
        if self._objectives_version != self.objective._structure_version:
            raise RuntimeError("The objective was modified after optimizer construction, which is currently not supported.")
        kwargs.pop("__FROM_THESEUS_LAYER_TOKEN__", None)
        dev = self.objective.device
        if isinstance(dev, torch.device) and dev.type == "cuda" and torch.cuda.is_available():
            # the kernels are enqueued on torch's current stream OF THE CURRENT DEVICE: make the objective's device current for the call
            # (an objective on cuda:1 while cuda:0 is current would otherwise launch into the wrong context)
            with torch.cuda.device(dev):
                return self._optimize_impl(**kwargs)
        return self._optimize_impl(**kwargs)

    # ---- bookkeeping (nonlinear_optimizer.py:109-213) ----
    @torch.no_grad()
    def _check_convergence(self, err: torch.Tensor, last_err: torch.Tensor):
        if self.params.abs_err_tolerance <= 0 and self.params.rel_err_tolerance <= 0:
            # |x| < tol is identically False for tol <= 0: no device->host sync needed for the fixed-iteration runs
            return None
        if self.process_group is not None:
            from .distributed import global_mean_abs_error
            mean_abs = global_mean_abs_error(err, self.process_group)[0]  # mean over the GLOBAL batch
        else:
            mean_abs = err.abs().mean()
        if mean_abs < self.params.abs_err_tolerance:
            return torch.ones_like(err).bool()
        err_change = last_err - err
        return (err_change.abs() < self.params.abs_err_tolerance).logical_or(
            (err_change / last_err).abs() < self.params.rel_err_tolerance)

    def _init_info(self, track_best_solution: bool, track_err_history: bool, track_state_history: bool) -> NonlinearOptimizerInfo:
        with torch.no_grad():
            last_err = self.objective.engine().error_metric("cur")
        B = self.objective.batch_size
        best_err = last_err.clone() if track_best_solution else None
        err_history = None
        if track_err_history:
            err_history = torch.ones(B, self.params.max_iterations + 1) * math.inf
            err_history[:, 0] = last_err.clone().cpu()
        state_history = None
        if track_state_history:
            state_history = {}
            for var in self.objective.optim_vars.values():
                state_history[var.name] = torch.ones(B, *var.shape[1:], self.params.max_iterations + 1) * math.inf
                state_history[var.name][..., 0] = var.tensor.detach().clone().cpu()
        best_solution = None
        if track_best_solution:
            best_solution = {v.name: v.tensor.detach().clone().cpu() for v in self.objective.optim_vars.values()}
        return NonlinearOptimizerInfo(
            best_solution=best_solution, last_err=last_err, best_err=best_err,
            status=np.array([NonlinearOptimizerStatus.START] * B),
            converged_iter=torch.zeros_like(last_err, dtype=torch.long), best_iter=torch.zeros_like(last_err, dtype=torch.long),
            err_history=err_history, state_history=state_history)

    def _update_info(self, info: NonlinearOptimizerInfo, current_iter: int, err: torch.Tensor, converged_indices):
        if converged_indices is None:
            info.converged_iter += 1
        else:
            info.converged_iter += 1 - converged_indices.long()
        if info.err_history is not None:
            info.err_history[:, current_iter + 1] = err.clone().cpu()
        if info.state_history is not None:
            for var in self.objective.optim_vars.values():
                info.state_history[var.name][..., current_iter + 1] = var.tensor.detach().clone().cpu()
        if info.best_solution is not None:
            good = err < info.best_err
            info.best_iter[good] = current_iter
            gc = good.cpu()
            for var in self.objective.optim_vars.values():
                info.best_solution[var.name][gc] = var.tensor.detach().clone()[good].cpu()
            info.best_err = torch.minimum(info.best_err, err)

    def _split_backward_iters(self, **kwargs):
        mode = kwargs["backward_mode"]
        if mode == BackwardMode.TRUNCATED:
            if "backward_num_iterations" not in kwargs:
                raise ValueError("backward_num_iterations expected but not received.")
            bwd = min(kwargs["backward_num_iterations"], self.params.max_iterations)
        else:
            bwd = {BackwardMode.UNROLL: self.params.max_iterations, BackwardMode.DLM: self.params.max_iterations,
                   BackwardMode.IMPLICIT: 1}[mode]
        return bwd, self.params.max_iterations - bwd

    # ---- the hot loop (nonlinear_least_squares.py:100-215) ----
    def _optimize_loop(self, num_iter: int, info: NonlinearOptimizerInfo, verbose: bool, end_iter_callback=None, **kwargs) -> int:
        eng = self.objective.engine()
        B = self.objective.batch_size
        if kwargs.pop("cuda_graph", self.cuda_graph) and self.params.abs_err_tolerance <= 0 and self.params.rel_err_tolerance <= 0 \
                and not torch.is_grad_enabled():
            return self._optimize_loop_graphed(num_iter, info, verbose, end_iter_callback, **kwargs)
        converged_indices = None  # None == all False (kept on the host side to avoid a sync per iteration)
        iters_done = 0
        it_ = 0
        all_reject_attempts = 0
        lin = self.linear_solver.linearization
        while it_ < num_iter:
            lin.linearize()
            run_err = None
            try:
                delta = self.compute_delta(**kwargs)
            except RuntimeError as e:
                run_err = e
            if self.process_group is not None:   # a failed solve on ANY rank ends the loop on every rank (collectives stay matched)
                from .distributed import any_rank_true
                if any_rank_true(run_err is not None, self.process_group, self.objective.device) and run_err is None:
                    run_err = RuntimeError("the linear solve failed on another rank of the process group")
            if run_err is not None:
                msg = f"There was an error while running the linear optimizer. Original error message: {run_err}."
                if torch.is_grad_enabled() or getattr(self, "_grad_mode_at_entry", False):
                    raise RuntimeError(msg + " Backward pass will not work. To obtain the best solution seen before the error, run with torch.no_grad()")
                warnings.warn(msg, RuntimeWarning)
                info.status[:] = NonlinearOptimizerStatus.FAIL
                return iters_done
            err, all_rejected = self._step(delta, info.last_err, converged_indices, **kwargs)
            if all_rejected:
                all_reject_attempts += 1
                if all_reject_attempts < NonlinearLeastSquares._MAX_ALL_REJECT_ATTEMPTS:
                    continue
            all_reject_attempts = 0
            with torch.no_grad():
                self._update_info(info, it_, err, converged_indices)
                if verbose:
                    print(f"Nonlinear optimizer. Iteration: {it_+1}. Error: {err.mean().item()}")
                converged_indices = self._check_convergence(err, info.last_err)
                if converged_indices is not None:
                    cpu_conv = converged_indices.cpu().numpy()
                    info.status[cpu_conv] = NonlinearOptimizerStatus.CONVERGED
                    all_conv = bool(cpu_conv.all())
                    if self.process_group is not None:  # every rank must take the same exit (collectives stay matched)
                        from .distributed import all_rejected as _count_equals_total
                        all_conv = _count_equals_total(int(cpu_conv.sum()), len(cpu_conv), self.process_group, err.device)
                    if all_conv:
                        break
                info.last_err = err
                if end_iter_callback is not None:
                    end_iter_callback(self, info, delta, it_)
            iters_done += 1
            it_ += 1
        info.status[info.status == NonlinearOptimizerStatus.START] = NonlinearOptimizerStatus.MAX_ITERATIONS
        return iters_done

    # ---- the same loop with the iteration body replayed from a CUDA graph ----
    def _iteration_body(self, prev_err: torch.Tensor, **kwargs):
        """One iteration on the current stream, no host reads: returns (delta, err, reject or None, stats or None)."""
        eng = self.objective.engine()
        self.linear_solver.linearization.linearize()
        delta = self.compute_delta(**kwargs)
        eng.retract_into(delta, self._tmp_optim_vars, float(self.params.step_size), None)
        err_new = eng.error_metric("tmp")
        reject, err, stats = self._complete_step(delta, err_new, prev_err, **kwargs)
        stats = stats if torch.is_tensor(stats) else None  # GN / non-adaptive LM: no accept test, nothing to read back
        eng.commit(reject)  # all-rejected == every item keeps its old value: the commit is unconditional
        prev_err.copy_(err)  # rejected items carry their previous error (the control kernel folds that in)
        return delta, err, reject, stats

    def _optimize_loop_graphed(self, num_iter: int, info: NonlinearOptimizerInfo, verbose: bool, end_iter_callback=None, **kwargs) -> int:
        eng = self.objective.engine()
        B = self.objective.batch_size
        solver = self.linear_solver
        d = getattr(self, "_damping", None)
        if d is not None and not torch.is_tensor(d):  # a python-float damping would be copied host->device inside the capture
            self._damping = torch.full((B,), float(d), dtype=self.objective.dtype, device=self.objective.device)
        eng._bind("cur"); eng._bind("tmp")  # pointer tables current BEFORE deciding whether the captured graph is still valid
        key = (B, eng.table_version, id(getattr(self, "_damping", None)) if torch.is_tensor(getattr(self, "_damping", None)) else None,
               tuple(sorted((k, repr(v)) for k, v in kwargs.items())))
        g = self._graph
        if g is None or g["key"] != key:
            prev = info.last_err.clone()
            solver.defer_info_check, self._capturing = True, True
            try:
                # warm-up on a side stream (allocations, plan uploads, lazy initialisation), restoring the state it advanced
                state = [v.tensor.clone() for v in self.ordering]
                damp = self._damping.clone() if torch.is_tensor(getattr(self, "_damping", None)) else None
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._iteration_body(prev, **kwargs)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                for v, t in zip(self.ordering, state):
                    v.tensor.copy_(t)
                if damp is not None:
                    self._damping.copy_(damp)
                prev.copy_(info.last_err)
                key = (key[0], eng.table_version) + key[2:]
                graph = torch.cuda.CUDAGraph()
                l0 = int(_lib.load().thb_launch_count())
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):  # other threads (NCCL watchdog, samplers) may touch CUDA
                    out = self._iteration_body(prev, **kwargs)
                body_launches = int(_lib.load().thb_launch_count()) - l0
            finally:
                solver.defer_info_check, self._capturing = False, False
            g = self._graph = dict(key=key, graph=graph, prev=prev, out=out, launches=body_launches)
        else:
            g["prev"].copy_(info.last_err)
        delta, err, reject, stats = g["out"]
        iters_done = it_ = all_reject_attempts = 0
        while it_ < num_iter:
            g["graph"].replay()
            _lib.replayed_launches += g["launches"]
            all_rejected = False
            if stats is not None:
                all_rejected = self._read_stats(stats, B)
            run_err = None
            try:
                solver.check_info()
            except RuntimeError as e:
                run_err = e
            if self.process_group is not None:
                from .distributed import any_rank_true
                if any_rank_true(run_err is not None, self.process_group, self.objective.device) and run_err is None:
                    run_err = RuntimeError("the linear solve failed on another rank of the process group")
            if run_err is not None:
                msg = f"There was an error while running the linear optimizer. Original error message: {run_err}."
                if getattr(self, "_grad_mode_at_entry", False):
                    raise RuntimeError(msg + " Backward pass will not work. To obtain the best solution seen before the error, run with torch.no_grad()")
                warnings.warn(msg, RuntimeWarning)
                info.status[:] = NonlinearOptimizerStatus.FAIL
                return iters_done
            if all_rejected:
                all_reject_attempts += 1
                if all_reject_attempts < NonlinearLeastSquares._MAX_ALL_REJECT_ATTEMPTS:
                    continue
            all_reject_attempts = 0
            with torch.no_grad():
                cur_err = err.clone()
                self._update_info(info, it_, cur_err, None)
                if verbose:
                    print(f"Nonlinear optimizer. Iteration: {it_+1}. Error: {cur_err.mean().item()}")
                info.last_err = cur_err
                if end_iter_callback is not None:
                    end_iter_callback(self, info, delta, it_)
            iters_done += 1
            it_ += 1
        info.status[info.status == NonlinearOptimizerStatus.START] = NonlinearOptimizerStatus.MAX_ITERATIONS
        return iters_done

    def _optimize_impl(self, track_best_solution: bool = False, track_err_history: bool = False, track_state_history: bool = False,
                       verbose: bool = False, backward_mode: Union[str, BackwardMode] = BackwardMode.UNROLL,
                       end_iter_callback=None, **kwargs) -> OptimizerInfo:
        backward_mode = BackwardMode.resolve(backward_mode)
        if torch.is_grad_enabled() and any(v.tensor.requires_grad for v in list(self.objective.optim_vars.values()) + list(self.objective.aux_vars.values())):
            return self._optimize_impl_differentiable(track_best_solution, track_err_history, track_state_history, verbose, backward_mode,
                                                      end_iter_callback, **kwargs)
        kwargs_plus = {**kwargs, "backward_mode": backward_mode}
        eng = self.objective.engine()
        eng.adopt_optim_vars()
        self.reset(**kwargs_plus)
        # nothing requires grad: the loop runs without a tape, but a failing linear solve is still reported the way the caller's grad
        # mode asks for (nonlinear_least_squares.py:138-151: RuntimeError with gradients enabled, warning + status FAIL under no_grad)
        self._grad_mode_at_entry = torch.is_grad_enabled()
        try:
            with torch.no_grad():
                info = self._init_info(track_best_solution, track_err_history, track_state_history)
                if verbose:
                    print(f"Nonlinear optimizer. Iteration: 0. Error: {info.last_err.mean().item()}")
                self._optimize_loop(num_iter=self.params.max_iterations, info=info, verbose=verbose,
                                    end_iter_callback=end_iter_callback, **kwargs)
        finally:
            self._grad_mode_at_entry = False
        info.converged_iter[torch.from_numpy(info.status == NonlinearOptimizerStatus.MAX_ITERATIONS).to(info.converged_iter.device)] = -1
        return info

    # ---- backward modes (nonlinear_least_squares.py:233-294, theseus_layer.py:45-97) ----
    def _optimize_impl_differentiable(self, track_best_solution, track_err_history, track_state_history, verbose, backward_mode,
                                      end_iter_callback, **kwargs) -> OptimizerInfo:
        """UNROLL: every iteration on the autograd tape.  TRUNCATED: the last `backward_num_iterations` only.  IMPLICIT: the whole
        loop under no_grad, then ONE differentiable, undamped Gauss-Newton step with the Hessian detached (implicit function
        theorem at the fixed point).  On the tape an iteration is: differentiable linearization (torch.func Jacobians of the
        AutoDiffCostFunctions) -> autograd.LinearSolveFunction (fused CUDA forward, closed-form CUDA backward) -> x + delta."""
        from .core import CostFunction
        if backward_mode == BackwardMode.DLM:
            raise NotImplementedError("BackwardMode.DLM is not built (SURVEY.md 8: out of scope)")
        if self.process_group is not None:
            # the taped loop takes local exits (all-converged / all-rejected decided per rank): refused rather than risking unmatched
            # collectives; shard the batch OUTSIDE the layer for training (each rank its own objective), or run under no_grad
            raise NotImplementedError("process_group is supported by the no-grad LM / GN loop only, not by the backward modes")
        def has_torch(cf):
            inner = getattr(cf, "cost_function", cf)
            return type(inner)._torch_error is not CostFunction._torch_error
        bad_cf = [cf.name for cf in self.objective.cost_functions.values() if not has_torch(cf)]
        bad_v = [v.name for v in self.ordering if v.KIND not in (0, 1, 2, 3, 4)]
        if bad_cf or bad_v:
            raise NotImplementedError(
                f"theseus_b200: cost functions {bad_cf[:3]} / variables {bad_v[:3]} have no torch restatement, so they cannot be put on "
                "the autograd tape of the backward modes -- wrap the call in torch.no_grad().")
        kwargs_plus = {**kwargs, "backward_mode": backward_mode}
        self.reset(**kwargs_plus)
        with torch.no_grad():
            info = self._init_info(track_best_solution, track_err_history, track_state_history)
        bwd_iters, nograd_iters = self._split_backward_iters(**kwargs_plus)
        if backward_mode == BackwardMode.UNROLL:
            self._optimize_loop_differentiable(bwd_iters, info, verbose, end_iter_callback, False, **kwargs)
            info.converged_iter[torch.from_numpy(info.status == NonlinearOptimizerStatus.MAX_ITERATIONS).to(info.converged_iter.device)] = -1
            return info
        with torch.no_grad():
            self.objective.engine().adopt_optim_vars()
            done0 = self._optimize_loop(num_iter=nograd_iters, info=info, verbose=verbose, end_iter_callback=end_iter_callback, **kwargs)
        with torch.no_grad():
            ginfo = self._init_info(track_best_solution, track_err_history, track_state_history)
        done1 = self._optimize_loop_differentiable(bwd_iters, ginfo, verbose, end_iter_callback, backward_mode == BackwardMode.IMPLICIT, **kwargs)
        self._merge_infos(ginfo, done0, done1, info)
        return info

    def _merge_infos(self, ginfo, done0: int, done1: int, info) -> None:
        """nonlinear_optimizer.py:215-271: fold the grad-loop's bookkeeping into the no-grad loop's."""
        total = min(done0 + done1, self.params.max_iterations)
        info.status = ginfo.status
        info.converged_iter = ginfo.converged_iter + done0
        info.converged_iter[torch.from_numpy(ginfo.status == NonlinearOptimizerStatus.MAX_ITERATIONS).to(info.converged_iter.device)] = -1
        info.last_err = ginfo.last_err
        if info.err_history is not None:
            info.err_history[:, done0:total + 1] = ginfo.err_history[:, :total + 1 - done0]
        if info.state_history is not None:
            for k in info.state_history:
                info.state_history[k][..., done0:total + 1] = ginfo.state_history[k][..., :total + 1 - done0]
        if info.best_solution is not None:
            better = (ginfo.best_err < info.best_err).cpu()
            for k in info.best_solution:
                info.best_solution[k][better] = ginfo.best_solution[k][better]
            info.best_iter[better.to(info.best_iter.device)] = ginfo.best_iter[better.to(info.best_iter.device)] + done0
            info.best_err = torch.minimum(info.best_err, ginfo.best_err)

    def _optimize_loop_differentiable(self, num_iter: int, info, verbose: bool, end_iter_callback, last_implicit_step: bool, **kwargs) -> int:
        """nonlinear_least_squares.py:100-215 with every value-carrying op on the autograd tape (retraction: lie_torch.retract)."""
        from . import lie_torch
        eng = self.objective.engine()
        lin = self.linear_solver.linearization
        S = eng.structure
        converged = None
        iters_done = it_ = rejects_in_a_row = 0
        while it_ < num_iter:
            lin.linearize(_detach_hessian=last_implicit_step, differentiable=True)
            try:
                if last_implicit_step:
                    try:
                        delta = self.linear_solver.solve()  # the implicit-function derivation wants a plain GN step
                    except RuntimeError:
                        if kwargs.get("__strict_implicit_final_gn__", False):
                            raise
                        delta = self.compute_delta(**kwargs)
                else:
                    delta = self.compute_delta(**kwargs)
            except RuntimeError as run_err:
                raise RuntimeError(f"There was an error while running the linear optimizer. Original error message: {run_err}. "
                                   "Backward pass will not work. To obtain the best solution seen before the error, run with torch.no_grad()")
            step = 1.0 if (last_implicit_step and not kwargs.get("__keep_final_step_size__", False)) else float(self.params.step_size)
            olds = [v.tensor for v in self.ordering]
            news = []
            for v, old, c0, d in zip(self.ordering, olds, S.var_start_cols, S.var_dims):
                new = lie_torch.retract(v.KIND, old if old.shape[0] == delta.shape[0] else old.expand(delta.shape[0], *old.shape[1:]),
                                        step * delta[:, int(c0):int(c0) + int(d)])
                if converged is not None and not last_implicit_step:
                    new = torch.where(converged.view(-1, *([1] * (new.ndim - 1))), old.expand_as(new), new)
                news.append(new)
            for v, new in zip(self.ordering, news):
                v.update(new)
            with torch.no_grad():
                err_new = eng.error_metric("cur")
            reject, err, all_rejected = (None, err_new, False) if last_implicit_step else \
                self._complete_step(delta.detach(), err_new, info.last_err, **kwargs)
            if reject is not None:
                if all_rejected:
                    for v, old in zip(self.ordering, olds):
                        v.update(old)
                    rejects_in_a_row += 1
                    if rejects_in_a_row < NonlinearLeastSquares._MAX_ALL_REJECT_ATTEMPTS:
                        continue
                    err = info.last_err
                else:
                    rj = reject.bool()
                    for v, old, new in zip(self.ordering, olds, news):
                        v.update(torch.where(rj.view(-1, *([1] * (new.ndim - 1))), old.expand_as(new), new))
            rejects_in_a_row = 0
            with torch.no_grad():
                self._update_info(info, it_, err, converged)
                if verbose:
                    print(f"Nonlinear optimizer. Iteration: {it_+1}. Error: {err.mean().item()}")
                converged = self._check_convergence(err, info.last_err)
                if converged is not None:
                    cpu_conv = converged.cpu().numpy()
                    info.status[cpu_conv] = NonlinearOptimizerStatus.CONVERGED
                    if bool(cpu_conv.all()):
                        break
                info.last_err = err
                if end_iter_callback is not None:
                    end_iter_callback(self, info, delta, it_)
            iters_done += 1
            it_ += 1
        info.status[info.status == NonlinearOptimizerStatus.START] = NonlinearOptimizerStatus.MAX_ITERATIONS
        return iters_done

    # ---- step (nonlinear_least_squares.py:296-365) ----
    def _step(self, delta: torch.Tensor, previous_err: torch.Tensor, converged_indices, **kwargs):
        eng = self.objective.engine()
        step = float(self.params.step_size)
        eng.retract_into(delta, self._tmp_optim_vars, step, converged_indices)
        err_new = eng.error_metric("tmp")
        reject, err, all_rejected = self._complete_step(delta, err_new, previous_err, **kwargs)
        if reject is not None and all_rejected:
            return previous_err, True
        eng.commit(reject)
        return err, False

    def reset(self, **kwargs) -> None:
        self.linear_solver.reset(**kwargs)

    def _complete_step(self, delta, new_err, previous_err, **kwargs):
        return None, new_err, False

    def compute_delta(self, **kwargs) -> torch.Tensor:
        raise NotImplementedError


class Vectorize:
    """theseus/core/vectorizer.py:38-90 by name only: the reference rewires an objective so that cost functions of equal schema are
    evaluated in one batched call.  Here that grouping IS the engine (engine.py: one fused kernel launch per schema group, always on),
    so constructing this marks the objective and changes nothing."""

    def __init__(self, objective: Objective, empty_cuda_cache: bool = False):
        self._objective = objective
        objective.vectorized = True


class LinearOptimizerStatus(Enum):
    START = 0
    CONVERGED = 1
    FAIL = -1


class LinearOptimizer:
    """theseus/optimizer/linear/linear_optimizer.py:25-83: ONE linearize -> solve -> retract of the objective (the solution of a linear
    least-squares problem), same constructor and failure convention as the reference."""

    def __init__(self, objective: Objective, linear_solver_cls: Type[LinearSolver], *args, vectorize: bool = False,
                 linearization_cls: Optional[Type[Linearization]] = None, linearization_kwargs: Optional[Dict[str, Any]] = None,
                 linear_solver_kwargs: Optional[Dict[str, Any]] = None, **kwargs):
        self.objective = objective
        self.linear_solver = linear_solver_cls(objective, linearization_cls=linearization_cls, linearization_kwargs=linearization_kwargs or {},
                                               **(linear_solver_kwargs or {}))

    def optimize(self, **kwargs) -> OptimizerInfo:
        return self._optimize_impl(**kwargs)

    def _optimize_impl(self, **kwargs) -> OptimizerInfo:
        info = OptimizerInfo(best_solution={}, status=np.array([LinearOptimizerStatus.START] * self.objective.batch_size))
        lin = self.linear_solver.linearization
        try:
            self.objective.engine(getattr(lin, "_ordering_names", None)).adopt_optim_vars()
            lin.linearize()
            delta = self.linear_solver.solve()
        except RuntimeError as run_err:
            msg = f"There was an error while running the linear optimizer. Original error message: {run_err}."
            if torch.is_grad_enabled():
                raise RuntimeError(msg + " Backward pass will not work. To obtain the best solution seen before the error, run with torch.no_grad()")
            warnings.warn(msg, RuntimeWarning)
            info.status[:] = LinearOptimizerStatus.FAIL
            return info
        self.objective.retract_vars_sequence(delta, lin.ordering)
        info.status[:] = LinearOptimizerStatus.CONVERGED
        for var in lin.ordering:
            info.best_solution[var.name] = var.tensor.clone().cpu()
        return info


class GaussNewton(NonlinearLeastSquares):
    """theseus/optimizer/nonlinear/gauss_newton.py:17-47."""

    def compute_delta(self, **kwargs) -> torch.Tensor:
        return self.linear_solver.solve()


class LevenbergMarquardt(NonlinearLeastSquares):
    """theseus/optimizer/nonlinear/levenberg_marquardt.py:51-201."""
    _MIN_DAMPING = 1.0e-7
    _MAX_DAMPING = 1.0e7

    def __init__(self, objective: Objective, *args, **kwargs):
        super().__init__(objective, *args, **kwargs)
        self._allows_ellipsoidal = True
        self._allows_adaptive = True
        self._damping: Union[float, torch.Tensor] = 0.001
        self._stats_host = None

    def reset(self, damping: float = 1e-3, adaptive_damping: bool = False, **kwargs) -> None:
        super().reset(**kwargs)
        if adaptive_damping:
            d = self._damping
            B = self.objective.batch_size
            if torch.is_tensor(d) and d.shape == (B,) and d.dtype == self.objective.dtype and d.device == torch.device(self.objective.device):
                d.fill_(damping)  # same buffer every call: its address is baked into a captured graph; the control kernel updates it in place
            else:
                self._damping = damping * torch.ones(B, device=self.objective.device, dtype=self.objective.dtype)
        else:
            self._damping = damping

    def compute_delta(self, ellipsoidal_damping: bool = False, damping_eps: Optional[float] = None, **kwargs) -> torch.Tensor:
        damping_eps = damping_eps if damping_eps is not None else 1e-8
        return self.linear_solver.solve(damping=self._damping, ellipsoidal_damping=ellipsoidal_damping, damping_eps=damping_eps)

    def _complete_step(self, delta, new_err, previous_err, adaptive_damping: bool = False, down_damping_ratio: float = 9.0,
                       up_damping_ratio: float = 11.0, damping_accept: float = 0.1, ellipsoidal_damping: bool = False, **kwargs):
        if not adaptive_damping:
            return None, new_err, False
        return self._check_accept(delta, new_err, previous_err, damping_accept, down_damping_ratio, up_damping_ratio, ellipsoidal_damping)

    @torch.no_grad()
    def _check_accept(self, delta, err, previous_err, damping_accept, down_damping_ratio, up_damping_ratio, ellipsoidal_damping):
        """levenberg_marquardt.py:172-201 as one kernel (thb_lm_control): rho test, damping update/clamp, reject mask,
        committed error; only the number of rejected items comes back to the host (for the all-rejected retry rule)."""
        lin = self.linear_solver.linearization
        lib = _lib.load()
        B, n = delta.shape
        dev = delta.device
        sfx = "f64" if delta.dtype == torch.float64 else "f32"
        Atb = lin.Atb.reshape(B, n)
        diag = lin.diagonal_scaling(torch.ones(B, n, dtype=delta.dtype, device=dev)) if ellipsoidal_damping else None
        reject = torch.empty(B, dtype=torch.uint8, device=dev)
        err_out = torch.empty_like(err)
        stats = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.check(getattr(lib, f"thb_lm_control_{sfx}")(
            _lib.ptr(delta), _lib.ptr(Atb), _lib.ptr(diag), B, n, float(self.params.step_size), _lib.ptr(previous_err), _lib.ptr(err),
            _lib.ptr(self._damping), 1 if ellipsoidal_damping else 0, float(damping_accept), float(down_damping_ratio),
            float(up_damping_ratio), _lib.ptr(reject), _lib.ptr(err_out), _lib.ptr(stats), _lib.stream_ptr()), "lm_control")
        self._keep_control = (delta, Atb, diag, err, previous_err)
        if getattr(self, "_capturing", False):  # CUDA-graph capture: the statistics are read after the replay (_read_stats)
            return reject, err_out, stats
        return reject, err_out, self._read_stats(stats, B)

    def _read_stats(self, stats, B) -> bool:
        """all-rejected? -- the ONE host read per iteration (16 bytes)."""
        if self._stats_host is None:
            self._stats_host = torch.empty(4, dtype=torch.int32).pin_memory()
        if self.process_group is not None:
            # the single per-iteration collective: [#rejected, #items] summed over ranks (NCCL, latency-bound)
            from .distributed import reduce_counts
            stats[1] = B
            reduce_counts(stats, self.process_group)
        self._stats_host.copy_(stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n_rej = int(self._stats_host[0])
        total = int(self._stats_host[1]) if self.process_group is not None else B
        return n_rej == total


class TrustRegion(NonlinearLeastSquares):
    """theseus/optimizer/nonlinear/trust_region.py:24-151 (Nocedal & Wright ch. 4): the step is accepted / the radius adapted from
    rho = actual / predicted reduction, the prediction from the linearization already on the device (Atb, A v)."""

    def __init__(self, objective: Objective, *args, **kwargs):
        super().__init__(objective, *args, **kwargs)
        if self.process_group is not None:   # its all-rejected decision is local (reject.all()): ranks could take different exits
            raise NotImplementedError("process_group is supported by GaussNewton / LevenbergMarquardt only, not by the trust-region optimizers")
        self._trust_region: torch.Tensor = None

    def reset(self, trust_region_init: float = 0.5, **kwargs) -> None:
        super().reset(**kwargs)
        self._trust_region = trust_region_init * torch.ones(self.objective.batch_size, 1, device=self.objective.device, dtype=self.objective.dtype)

    def _compute_delta_impl(self) -> torch.Tensor:
        raise NotImplementedError

    def compute_delta(self, **kwargs) -> torch.Tensor:
        return self._compute_delta_impl()

    @staticmethod
    def _squared_norm(tensor: torch.Tensor, keepdim: bool = True) -> torch.Tensor:
        return (tensor ** 2).sum(dim=1, keepdim=keepdim)

    def _predicted_error(self, previous_error: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
        lin = self.linear_solver.linearization
        Adelta = lin.Av(delta)
        grad = -lin.Atb.squeeze(2)
        return previous_error + (delta * grad).sum(dim=1) + 0.5 * TrustRegion._squared_norm(Adelta, keepdim=False)

    def _compute_rho(self, delta, previous_err, new_err) -> torch.Tensor:
        pred_err = self._predicted_error(previous_err, delta)
        return ((previous_err - new_err) / (previous_err - pred_err)).view(-1, 1)

    @torch.no_grad()
    def _complete_step(self, delta, new_err, previous_err, accept_threshold: float = 0.0, shrink_threshold: float = 0.25,
                       expand_threshold: float = 0.75, shrink_ratio: float = 0.25, expand_ratio: float = 2.0,
                       min_trust_region: float = 1.0e-5, max_trust_region: float = 1.0e5, **kwargs):
        good = (0.0 < shrink_ratio <= 1.0) and (expand_ratio >= 1.0) and (shrink_threshold < expand_threshold) and (accept_threshold < shrink_threshold)
        if not good:
            raise ValueError("Invalid parameters for TrustRegionMethod. Values must satisfy <accept/shrink>_threshold < expand_threshold, "
                             "shrink_ratio in (0, 1], and expand_ratio > 1.0.")
        rho = self._compute_rho(delta, previous_err, new_err)
        tr = torch.where(rho < shrink_threshold, self._trust_region * shrink_ratio, self._trust_region)
        tr = torch.where(rho > expand_threshold, tr * expand_ratio, tr)
        self._trust_region = tr.clamp(min_trust_region, max_trust_region)
        reject = (rho < accept_threshold).view(-1)
        err = torch.where(reject, previous_err, new_err)
        return reject, err, bool(reject.all())


class Dogleg(TrustRegion):
    """theseus/optimizer/nonlinear/dogleg.py:14-105: Gauss-Newton step if it lies inside the trust region, else the dogleg path
    between the Cauchy point and the Gauss-Newton point.  The factorisation, A v, retraction and error are the fused kernels; the
    path arithmetic is a handful of [B,n] vector operations."""
    EPS = 1e-7

    def _compute_delta_impl(self) -> torch.Tensor:
        tr2 = self._trust_region ** 2
        delta_gn = self.linear_solver.solve()
        if (TrustRegion._squared_norm(delta_gn) < tr2).all():
            return delta_gn
        lin = self.linear_solver.linearization
        delta_sd = lin.Atb.squeeze(2)
        Adelta_sd_norm_2 = TrustRegion._squared_norm(lin.Av(delta_sd))
        grad_norm_2 = TrustRegion._squared_norm(delta_sd)
        cauchy = grad_norm_2 / (Adelta_sd_norm_2 + Dogleg.EPS)
        delta_c = delta_sd * cauchy
        delta_c_norm_2 = grad_norm_2 * (cauchy ** 2)
        inside = delta_c_norm_2 <= tr2
        delta_dogleg = delta_c if bool(inside.all()) else torch.where(inside, delta_c, delta_c * self._trust_region / (delta_c_norm_2 + Dogleg.EPS).sqrt())
        if bool(inside.any()):
            diff = delta_gn - delta_c
            a = TrustRegion._squared_norm(diff)
            b = (2 * delta_c * diff).sum(dim=1, keepdim=True)
            c = delta_c_norm_2 - tr2
            disc = ((b ** 2) - 4 * a * c).clamp(Dogleg.EPS)
            tau = ((-b + disc.sqrt()) / (2 * a + Dogleg.EPS)).clamp(max=1.0)
            delta_dogleg = torch.where(inside, delta_c + tau * diff, delta_dogleg)
        return delta_dogleg

