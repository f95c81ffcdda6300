"""Shared test helpers: load golden fixtures, build the same pose-graph problem for the oracle (numpy spec)
and for the product (theseus_b200 objective)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def pgo_spec(g, dtype=np.float64):
    """Oracle problem description from a pgo_*.npz fixture (layout of oracle/nls.py)."""
    poses0, edges, meas, edge_w = g["poses0"], g["edges"], g["meas"], g["edge_w"]
    spec = dict(dtype=np.dtype(dtype), vars=[], costs=[])
    for i in range(poses0.shape[0]):
        spec["vars"].append(dict(kind="SE3", dof=6, value=poses0[i].astype(dtype)))
    for e in range(edges.shape[0]):
        spec["costs"].append(dict(kind="between", group="SE3", vars=(int(edges[e, 0]), int(edges[e, 1])),
                                  aux=meas[e].astype(dtype), weight=("diag", edge_w[e].astype(dtype))))
    spec["costs"].append(dict(kind="local", group="SE3", vars=(0,), aux=poses0[0].astype(dtype),
                              weight=("scale", np.full((1, 1), float(g["prior_w"]), dtype=dtype))))
    return spec


def pgo_objective(th, g, device="cuda", dtype=None):
    """The product objective built exactly like examples/pose_graph/pose_graph_cube.py:56-83."""
    import torch
    dtype = dtype or torch.float64
    poses0, edges, meas, edge_w = g["poses0"], g["edges"], g["meas"], g["edge_w"]
    poses = [th.SE3(tensor=torch.from_numpy(poses0[i]).to(dtype), name=f"VERTEX_SE3__{i}") for i in range(poses0.shape[0])]
    objective = th.Objective(dtype=dtype)
    for e in range(edges.shape[0]):
        i, j = int(edges[e, 0]), int(edges[e, 1])
        z = th.SE3(tensor=torch.from_numpy(meas[e]).to(dtype), name=f"EDGE_SE3__{e}_{i}_{j}")
        w = th.DiagonalCostWeight(th.Variable(torch.from_numpy(edge_w[e]).to(dtype), name=f"EDGE_WEIGHT__{e}"))
        objective.add(th.Between(poses[i], poses[j], z, w, name=f"between_{e}"))
    prior = th.Difference(poses[0], th.SE3(tensor=torch.from_numpy(poses0[0]).to(dtype), name="VERTEX_SE3__0__PRIOR"),
                          th.ScaleCostWeight(torch.tensor(float(g["prior_w"]), dtype=dtype)), name="prior")
    objective.add(prior)
    objective.to(device)
    return objective, poses


def lm_kwargs_of(g):
    kw = eval(str(g["kwargs_json"]))  # written by tests/golden/make_golden.py (repr of a plain dict)
    method = kw.pop("method")
    iters = kw.pop("iters")
    return method, iters, kw


def decisive_iterations(err0, trace_err, tol=1e-7):
    """Number of leading LM iterations whose accept/reject decision is numerically well defined.

    rho = (err_prev - err_new) / den is 0/0 noise once the problem has converged to rounding level (the reference's
    own decisions become arbitrary there: pose-graph LM reaches its floor in 3-5 iterations).  An iteration is
    'decisive' if the reference reduced the error by more than `tol` relative, for every batch item."""
    prev = err0
    k = 0
    for it in range(trace_err.shape[0]):
        red = (prev - trace_err[it]) / prev
        if not np.all(red > tol):
            break
        prev = trace_err[it]
        k += 1
    return k
