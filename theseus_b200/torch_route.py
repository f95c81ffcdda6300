"""Batched torch route: the cost functions that are evaluated with torch.func instead of a fused kernel -- AutoDiffCostFunction, the
tactile costs, robust wrappers without a fused loss, SO2 costs, and EVERY cost function on the autograd tape of the backward modes --
grouped like the reference's Vectorize (theseus/core/vectorizer.py:222-332): cost functions with the same signature are stacked along the
batch dimension and evaluated by ONE vmap(jacrev) call, their Jacobians scattered into the batched CSR with one indexed copy.

The per-cost-function loop (CostFunction.generic_jacobians_error, engine.linearize_sparse_differentiable) stays the specification; this
module must reproduce it exactly (tests/test_torch_route.py compares both on every kind of objective the goldens hold).  The engine uses it
when THB_BATCHED_TORCH_ROUTE=1 (opt-in until it has run on a GPU: written after the round-1 GPU budget was spent).
"""
from typing import Callable, Dict, List, Sequence

import numpy as np
import torch

from .core import WEIGHT_SCALE, CostFunction, RobustCostFunction
from .embodied import MovingFrameBetween


def _weight(cf: CostFunction, w: torch.Tensor, err: torch.Tensor, jacs):
    """CostFunction._weight with the weight tensor passed in (cost_weight.py:81-90 Scale, :125-136 Diagonal)."""
    w = w.view(-1, 1) if cf.weight.WEIGHT_KIND == WEIGHT_SCALE else w
    err = err * w
    if jacs is not None:
        jacs = [J * w.unsqueeze(2) for J in jacs]
    return jacs, err


def _plain_jacobians_error(cf: CostFunction, opt_t, aux_t, w, differentiable: bool):
    """CostFunction.generic_jacobians_error on explicit, already expanded tensors."""
    from torch.func import jacrev, vmap

    def one(o, a):
        return cf._torch_error(tuple(x.unsqueeze(0) for x in o), tuple(x.unsqueeze(0) for x in a))[0]

    with torch.enable_grad():
        jacs = vmap(jacrev(one, argnums=0))(tuple(opt_t), tuple(aux_t))
        err = cf._torch_error(tuple(opt_t), tuple(aux_t))
    jacs = [type(v).project_tensor(t, j) for v, t, j in zip(cf.optim_vars, opt_t, jacs)]
    if differentiable:
        return _weight(cf, w, err, jacs)
    return _weight(cf, w.detach(), err.detach(), [j.detach() for j in jacs])


def _moving_frame_jacobians_error(cf: MovingFrameBetween, opt_t, aux_t, w, differentiable: bool):
    """MovingFrameBetween.generic_jacobians_error (the reference's Jacobian convention: no dlog factor) on explicit tensors."""
    from torch.func import jacrev, vmap
    from . import lie_torch
    k = cf.frame1.KIND
    B = opt_t[0].shape[0]

    def one(o):
        return cf._torch_frame_diff(tuple(x.unsqueeze(0) for x in o))[0]

    with torch.enable_grad():
        D = cf._torch_frame_diff(tuple(opt_t))
        dD = vmap(jacrev(one))(tuple(opt_t))
        err = lie_torch.local(k, aux_t[0], D)
    jacs = []
    for v, t, J in zip(cf.optim_vars, opt_t, dD):
        Jin = type(v).project_tensor(t, J.reshape(B, -1, *t.shape[1:]))
        Jin = Jin.reshape(B, *D.shape[1:], Jin.shape[-1]).movedim(-1, 1)
        jacs.append(lie_torch.velocity_to_tangent(k, D, Jin).transpose(1, 2))
    if differentiable:
        return _weight(cf, w, err, jacs)
    return _weight(cf, w.detach(), err.detach(), [j.detach() for j in jacs])


def _inner_jacobians_error(cf: CostFunction, opt_t, aux_t, w, differentiable: bool):
    if isinstance(cf, MovingFrameBetween):
        return _moving_frame_jacobians_error(cf, opt_t, aux_t, w, differentiable)
    return _plain_jacobians_error(cf, opt_t, aux_t, w, differentiable)


def _aux_variables(cf: CostFunction):
    """Variables whose tensors the torch form of `cf` reads besides the optimisation variables and the weight."""
    if isinstance(cf, RobustCostFunction):
        extra = [cf.log_loss_radius] + ([cf.gnc_control_val] if hasattr(cf, "gnc_control_val") else [])
        return list(cf.cost_function._torch_aux()) + extra
    return list(cf._torch_aux())


def stacked_jacobians_error(cf: CostFunction, opt_t, aux_t, w, differentiable: bool):
    """(weighted Jacobians, weighted error) of `cf`'s formula on explicit tensors of any leading size (the stacked batch of a group)."""
    if not isinstance(cf, RobustCostFunction):
        return _inner_jacobians_error(cf, opt_t, aux_t, w, differentiable)
    inner = cf.cost_function
    n_in = len(inner._torch_aux())
    jacs, err = _inner_jacobians_error(inner, opt_t, aux_t[:n_in], w, differentiable)
    loss_args = aux_t[n_in:]
    if cf.flatten_dims:
        sc = torch.sqrt(cf.loss.linearize((err ** 2).reshape(-1, 1), *[a.repeat_interleave(err.shape[1], dim=0) for a in loss_args])
                        + cf._EPS).reshape(err.shape)
    else:
        sc = torch.sqrt(cf.loss.linearize((err ** 2).sum(dim=1, keepdim=True), *loss_args) + cf._EPS)
    if not differentiable:
        sc = sc.detach()
    return [sc.unsqueeze(2) * J for J in jacs], sc * err


def stacked_error(cf: CostFunction, opt_t, aux_t, w):
    """Weighted error (robust: the entries whose squared norm is the loss value) on explicit tensors."""
    if not isinstance(cf, RobustCostFunction):
        return _weight(cf, w, cf._torch_error(tuple(opt_t), tuple(aux_t)), None)[1]
    inner = cf.cost_function
    n_in = len(inner._torch_aux())
    err = _weight(inner, w, inner._torch_error(tuple(opt_t), tuple(aux_t[:n_in])), None)[1]
    loss_args = aux_t[n_in:]
    if cf.flatten_dims:
        val = cf.loss.evaluate((err ** 2).reshape(-1, 1), *[a.repeat_interleave(err.shape[1], dim=0) for a in loss_args]).reshape(err.shape)
        return torch.sqrt(val + cf._EPS)
    val = cf.loss.evaluate((err ** 2).sum(dim=1, keepdim=True), *loss_args)
    return torch.ones_like(err) * torch.sqrt(val / cf.dim() + cf._EPS)


def signature(cf: CostFunction):
    """Cost functions with equal signatures compute the same formula on tensors of the same shapes: they can be stacked."""
    inner = cf.cost_function if isinstance(cf, RobustCostFunction) else cf
    w = cf.weight.weight_tensor().tensor
    return (type(cf), type(inner), id(getattr(inner, "_err_fn", None)), cf.dim(),
            tuple((type(v), tuple(v.tensor.shape[1:])) for v in cf.optim_vars),
            tuple(tuple(a.tensor.shape[1:]) for a in _aux_variables(cf)),
            type(cf.weight), tuple(w.shape[1:]),
            type(getattr(cf, "loss", None)), bool(getattr(cf, "flatten_dims", False)))


class TorchRoute:
    """Groups of stackable cost functions + the CSR positions of their Jacobian blocks (from structure.Structure)."""

    def __init__(self, costs: Sequence[CostFunction], cost_ids: Sequence[int], S):
        self.costs, self.S = costs, S
        groups: Dict[tuple, List[int]] = {}
        for f in cost_ids:
            groups.setdefault(signature(costs[f]), []).append(f)
        self.groups = list(groups.values())
        self._index: Dict[tuple, list] = {}

    def _indices(self, gi: int, device):
        """Per optimisation-variable slot: positions [K, d, dof] of the Jacobian entries inside a row of A_val; rows [K, d] of b."""
        key = (gi, str(device))
        if key not in self._index:
            S, ids = self.S, self.groups[gi]
            cf0 = self.costs[ids[0]]
            d = int(S.cost_dims[ids[0]])
            r = np.arange(d)[None, :, None]
            slots = []
            for kslot, v in enumerate(cf0.optim_vars):
                dof = v.dof()
                c = np.arange(dof)[None, None, :]
                off = np.array([int(S.row_block_starts[f]) for f in ids])[:, None, None]
                st = np.array([int(S.stride[f]) for f in ids])[:, None, None]
                p0 = np.array([int(S.block_pointers[f][kslot]) for f in ids])[:, None, None]
                slots.append(torch.from_numpy((off + r * st + p0 + c).reshape(-1).astype(np.int64)).to(device))
            rows = np.array([int(S.cost_row0[f]) for f in ids])[:, None] + np.arange(d)[None, :]
            self._index[key] = [slots, torch.from_numpy(rows.reshape(-1).astype(np.int64)).to(device)]
        return self._index[key]

    def _stack(self, gi: int, tensor_of: Callable, B: int):
        ids = self.groups[gi]
        ex = lambda t: t if t.shape[0] == B else t.expand((B,) + tuple(t.shape[1:]))
        cfs = [self.costs[f] for f in ids]
        n_opt = len(cfs[0].optim_vars)
        opt_t = [torch.cat([ex(tensor_of(cf.optim_vars[s])) for cf in cfs], dim=0) for s in range(n_opt)]
        auxs = [_aux_variables(cf) for cf in cfs]
        aux_t = [torch.cat([ex(a[s].tensor) for a in auxs], dim=0) for s in range(len(auxs[0]))]
        w = torch.cat([ex(cf.weight.weight_tensor().tensor) for cf in cfs], dim=0)
        return cfs[0], opt_t, aux_t, w

    def linearize(self, tensor_of: Callable, B: int, A_val: torch.Tensor, b: torch.Tensor, differentiable: bool):
        """Writes the Jacobian blocks and -error of every cost function of this route into A_val [B, nnz] / b [B, m] (in place).
        tensor_of(variable) -> the tensor to evaluate at (current or trial values)."""
        for gi, ids in enumerate(self.groups):
            K = len(ids)
            cf0, opt_t, aux_t, w = self._stack(gi, tensor_of, B)
            jacs, err = stacked_jacobians_error(cf0, opt_t, aux_t, w, differentiable)
            slots, rows = self._indices(gi, A_val.device)
            for idx, J in zip(slots, jacs):      # J [K*B, d, dof] -> [B, K*d*dof]
                A_val[:, idx] = J.reshape(K, B, -1).transpose(0, 1).reshape(B, -1)
            b[:, rows] = -err.reshape(K, B, -1).transpose(0, 1).reshape(B, -1)
        return A_val, b

    def half_squared_error(self, tensor_of: Callable, B: int) -> torch.Tensor:
        """sum over this route's cost functions of 0.5 * ||weighted error||^2 per batch item, [B]."""
        out = None
        for gi, ids in enumerate(self.groups):
            cf0, opt_t, aux_t, w = self._stack(gi, tensor_of, B)
            e = stacked_error(cf0, opt_t, aux_t, w)
            s = (e * e).sum(dim=1).reshape(len(ids), B).sum(dim=0) * 0.5
            out = s if out is None else out + s
        return out
