"""Config C1: examples/simple_example.py (curve fit y = v exp(x), AutoDiffCostFunction, GaussNewton + CholeskyDenseSolver,
batch 1, fp32) through the product on the GPU vs the reference's own run (tests/golden/simple_example.npz)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load

pytestmark = pytest.mark.gpu


def test_simple_example_matches_reference():
    g = load("simple_example")
    x = th.Variable(torch.zeros(1, 20), name="x")
    y = th.Variable(torch.from_numpy(g["y"]), name="y")
    v = th.Vector(1, name="v")

    def error_fn(optim_vars, aux_vars):
        xx, yy = aux_vars
        return yy.tensor - optim_vars[0].tensor * torch.exp(xx.tensor)

    objective = th.Objective()
    objective.add(th.AutoDiffCostFunction([v], error_fn, 20, aux_vars=[x, y], cost_weight=th.ScaleCostWeight(1.0)))
    objective.to("cuda")
    layer = th.TheseusLayer(th.GaussNewton(objective, max_iterations=10))
    with torch.no_grad():
        sol, info = layer.forward(input_tensors={"x": torch.from_numpy(g["x"]).cuda(), "v": torch.ones(1, 1, device="cuda")},
                                  optimizer_kwargs=dict(track_err_history=True))
    np.testing.assert_allclose(sol["v"].cpu().numpy(), g["v_final"], rtol=1e-5)
    k = int(g["converged_iter"][0])
    assert int(info.converged_iter[0]) == k
    assert info.status[0] == th.NonlinearOptimizerStatus.CONVERGED
    np.testing.assert_allclose(info.err_history[0, :k + 1].numpy(), g["err_history"][0, :k + 1], rtol=1e-4, atol=1e-7)
