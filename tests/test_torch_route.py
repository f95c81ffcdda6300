"""theseus_b200/torch_route.py (cost functions of equal signature stacked along the batch, ONE vmap(jacrev) per group) against the
per-cost-function loop it replaces (CostFunction.generic_jacobians_error / generic_error, the loop of engine.linearize_sparse and
linearize_sparse_differentiable), on every kind of objective the goldens hold -- values and gradients."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import theseus_b200 as th
from theseus_b200.structure import build_structure
from theseus_b200.torch_route import TorchRoute, signature
from helpers import load, pgo_objective, ba_objective, se2_pg_objective


def _golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _structure(objective):
    ordering = list(objective.optim_vars.values())
    index = {v.name: i for i, v in enumerate(ordering)}
    costs = list(objective.cost_functions.values())
    S = build_structure([v.dof() for v in ordering], [(cf.dim(), [index[v.name] for v in cf.optim_vars]) for cf in costs])
    return costs, S


def _per_cost(costs, S, B, dtype, differentiable):
    """The specification: engine.linearize_sparse_differentiable's loop, one cost function at a time."""
    ex = lambda t: t if t.shape[0] == B else t.expand((B,) + tuple(t.shape[1:]))
    A_val = torch.zeros(B, S.nnz, dtype=dtype)
    b = torch.zeros(B, S.num_rows, dtype=dtype)
    err = torch.zeros(B, dtype=dtype)
    for f, cf in enumerate(costs):
        jacs, e = cf.generic_jacobians_error([ex(v.tensor) for v in cf.optim_vars], differentiable=differentiable)
        d, st, off = int(S.cost_dims[f]), int(S.stride[f]), int(S.row_block_starts[f])
        blk = A_val[:, off:off + d * st].view(B, d, st)
        for kslot, J in enumerate(jacs):
            p0 = int(S.block_pointers[f][kslot])
            blk[:, :, p0:p0 + J.shape[2]] = J
        b[:, int(S.cost_row0[f]):int(S.cost_row0[f]) + d] = -e
        ge = cf.generic_error([ex(v.tensor) for v in cf.optim_vars])
        err = err + (ge * ge).sum(dim=1) * 0.5
    return A_val, b, err


def _objectives():
    G = _golden_module()
    out = {}
    for name in ("pgo_small_lm", "pgo_small_welsch", "pgo_small_geman"):
        out[name] = pgo_objective(th, load(name), device="cpu")[0]
    out["ba_small_huber"] = ba_objective(th, load("ba_small_huber"), device="cpu")[0]
    out["se2_pg"] = se2_pg_objective(th, load("se2_kat"), device="cpu")[0]
    g = load("tactile_kat")
    inputs = {k: torch.from_numpy(g[k]) for k in ("obj", "eff", "eff_meas", "mfb_meas", "c_square", "eff_radius", "sdf", "sdf_origin", "sdf_cell")}
    out["tactile"] = G.tactile_problem(th, torch, inputs, device="cpu")[0]
    g = load("so2_kat")
    out["so2"] = G.so2_problem(th, torch, torch.from_numpy(g["lm_thetas0"]), torch.from_numpy(g["lm_meas"]),
                               [tuple(int(x) for x in e) for e in g["lm_edges"]], g["lm_w_edge"], float(g["lm_w_prior"]))[0]
    g = load("autodiff_lie")
    out["autodiff_lie"] = G.autodiff_lie_problem(th, torch, {k: torch.from_numpy(g[k]) for k in ("T3", "T2", "p", "q", "p2", "q2")})
    return out


OBJECTIVES = None


def _get(name):
    global OBJECTIVES
    if OBJECTIVES is None:
        OBJECTIVES = _objectives()
    return OBJECTIVES[name]


@pytest.mark.parametrize("name", ["pgo_small_lm", "pgo_small_welsch", "pgo_small_geman", "ba_small_huber", "se2_pg", "tactile", "so2",
                                  "autodiff_lie"])
def test_batched_route_equals_per_cost_loop(name):
    objective = _get(name)
    costs, S = _structure(objective)
    B = max(v.tensor.shape[0] for v in objective.optim_vars.values())
    dtype = objective.dtype
    A0, b0, e0 = _per_cost(costs, S, B, dtype, differentiable=False)
    route = TorchRoute(costs, list(range(len(costs))), S)
    assert sum(len(g) for g in route.groups) == len(costs)
    if name.startswith("pgo"):
        assert len(route.groups) == 2          # every Between edge in one group, the prior in another
    A1 = torch.zeros(B, S.nnz, dtype=dtype)
    b1 = torch.zeros(B, S.num_rows, dtype=dtype)
    route.linearize(lambda v: v.tensor, B, A1, b1, differentiable=False)
    # stacking changes the batch size torch's batched matmuls see, hence their blocking: equal to rounding, not bitwise
    np.testing.assert_allclose(A1.numpy(), A0.numpy(), rtol=1e-12, atol=1e-13 * float(A0.abs().max()))
    np.testing.assert_allclose(b1.numpy(), b0.numpy(), rtol=1e-12, atol=1e-13 * float(b0.abs().max()))
    np.testing.assert_allclose(route.half_squared_error(lambda v: v.tensor, B).numpy(), e0.numpy(), rtol=1e-13)


@pytest.mark.parametrize("name", ["pgo_small_welsch", "pgo_small_geman", "tactile", "so2"])
def test_batched_route_gradients_equal_per_cost_loop(name):
    """differentiable=True: gradients of a scalar of (A_val, b) w.r.t. the cost weights, the auxiliary variables and the variable values."""
    objective = _get(name)
    costs, S = _structure(objective)
    B = max(v.tensor.shape[0] for v in objective.optim_vars.values())
    dtype = objective.dtype
    leaves = {}
    for cf in costs:
        for v in list(cf.optim_vars) + list(cf.aux_vars) + [cf.weight.weight_tensor()]:
            if v.tensor.is_floating_point() and id(v) not in leaves:
                leaves[id(v)] = v
    leaves = list(leaves.values())
    gen = torch.Generator().manual_seed(3)
    cA = torch.randn(B, S.nnz, generator=gen, dtype=dtype)
    cb = torch.randn(B, S.num_rows, generator=gen, dtype=dtype)
    grads = []
    for mode in ("loop", "route"):
        for v in leaves:
            v.tensor = v.tensor.detach().clone().requires_grad_(True)
        if mode == "loop":
            A, b, _ = _per_cost(costs, S, B, dtype, differentiable=True)
        else:
            A, b = torch.zeros(B, S.nnz, dtype=dtype), torch.zeros(B, S.num_rows, dtype=dtype)
            TorchRoute(costs, list(range(len(costs))), S).linearize(lambda v: v.tensor, B, A, b, differentiable=True)
        loss = (A * cA).sum() + (b * cb).sum()
        gs = torch.autograd.grad(loss, [v.tensor for v in leaves], allow_unused=True)
        grads.append([None if g is None else g.numpy().copy() for g in gs])
        for v in leaves:
            v.tensor = v.tensor.detach()
    n_used = 0
    for v, g0, g1 in zip(leaves, grads[0], grads[1]):
        assert (g0 is None) == (g1 is None), v.name
        if g0 is not None:
            n_used += 1
            np.testing.assert_allclose(g1, g0, rtol=1e-9, atol=1e-11 * max(1.0, np.abs(g0).max()), err_msg=v.name)
    assert n_used >= 3


def test_signatures_separate_what_must_not_be_stacked():
    objective = _get("tactile")
    costs, _ = _structure(objective)
    by_type = {}
    for cf in costs:
        by_type.setdefault(type(cf).__name__, set()).add(signature(cf))
    assert all(len(s) == 1 for s in by_type.values()), {k: len(v) for k, v in by_type.items()}   # same class + same shapes -> one group
    assert len({s for ss in by_type.values() for s in ss}) == len(by_type)
    a = th.Vector(2, name="a")
    f1 = th.AutoDiffCostFunction([a], lambda optim_vars, aux_vars: optim_vars[0].tensor, 2, name="f1")
    f2 = th.AutoDiffCostFunction([a], lambda optim_vars, aux_vars: 2 * optim_vars[0].tensor, 2, name="f2")
    assert signature(f1) != signature(f2)      # different user functions are never stacked
