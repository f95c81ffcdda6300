"""The reference's plugin contract for cost functions and cost weights -- subclass th.CostFunction / th.CostWeight, register variables,
implement error() / jacobians() / dim() (theseus/core/cost_function.py:64-149, cost_weight.py:20-55) -- through the engine's generic route,
on the CPU with the CUDA library replaced by its host emulation (tests/simt).  Problems: tests/user_costs.py (restatements of the
reference's own linearization_test_utils.py:122-196 and nonlinear/common.py:14-315).  The GPU twins are in tests/test_gpu_zz_first_run.py."""
import warnings

import numpy as np
import pytest
import torch

import theseus_b200 as th
import user_costs
from test_simt_engine_emulation import emu_lib, emulated  # noqa: F401  (fixtures)


def test_dense_linearization_of_user_defined_costs_is_exact(emulated):
    """test_dense_linearization.py:15-31: A, b, AtA, Atb of the 6 x 10 system, custom column order, custom matrix weights."""
    objective, ordering, A, b = user_costs.mock_linear_system(th)
    lin = th.DenseLinearization(objective, ordering=ordering)
    lin.linearize()
    assert lin.b.ndim == 2 and lin.A.shape == A.shape
    np.testing.assert_allclose(lin.A.numpy(), A.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(lin.b.numpy(), b.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(lin.AtA.numpy(), (A.transpose(1, 2) @ A).numpy(), rtol=1e-14)
    np.testing.assert_allclose(lin.Atb.numpy().reshape(A.shape[0], -1), (A.transpose(1, 2) @ b.unsqueeze(2)).squeeze(2).numpy(), rtol=1e-14)


def test_sparse_linearization_of_user_defined_costs_is_exact(emulated):
    """test_sparse_linearization.py:15-49: CSR -> dense equals A; Atb, Av and diagonal_scaling against the dense forms."""
    objective, ordering, A, b = user_costs.mock_linear_system(th)
    lin = th.SparseLinearization(objective, ordering=ordering)
    lin.linearize()
    B = A.shape[0]
    rp, ci = np.asarray(lin.A_row_ptr), np.asarray(lin.A_col_ind)
    dense = np.zeros(A.shape)
    for r in range(A.shape[1]):
        dense[:, r, ci[rp[r]:rp[r + 1]]] = lin.A_val.numpy()[:, rp[r]:rp[r + 1]]
    np.testing.assert_allclose(dense, A.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(lin.b.numpy(), b.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(lin.Atb.numpy().reshape(B, -1), (A.transpose(1, 2) @ b.unsqueeze(2)).squeeze(2).numpy(), rtol=1e-14)
    v = torch.randn(B, A.shape[2], generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    np.testing.assert_allclose(lin.Av(v).numpy(), (A @ v.unsqueeze(2)).squeeze(2).numpy(), rtol=1e-13)
    diag = (A * A).sum(dim=1)
    np.testing.assert_allclose(lin.diagonal_scaling(v).numpy(), (diag * v).numpy(), rtol=1e-13)


def _check_info(info, batch_size, max_iterations, initial_error, objective):
    """nonlinear/common.py:91-97."""
    assert info.err_history.shape == (batch_size, max_iterations + 1)
    hist = info.err_history                      # float32 on the CPU, like the reference's (nonlinear_optimizer.py:131-150)
    assert torch.allclose(hist[:, 0], initial_error.cpu().to(hist.dtype))
    assert torch.equal(info.err_history.argmin(dim=1), info.best_iter + 1)
    last = objective.error_metric().cpu()
    assert torch.allclose(hist[:, int(info.converged_iter.max())], last.to(hist.dtype))


LM_GRID = [dict(damping=d, ellipsoidal_damping=e, adaptive_damping=a, damping_eps=0.0)
           for d in (0.0, 0.001, 0.1) for e in (True, False) for a in (True, False)]


@pytest.mark.parametrize("multivar", [False, True])
@pytest.mark.parametrize("case", [("gn", {})] + [("lm", kw) for kw in LM_GRID[::3]] + [("dogleg", {})])
def test_regression_with_user_defined_costs_recovers_coefficients(emulated, case, multivar):
    """test_gauss_newton.py / test_levenberg_marquardt.py:25-39 / test_dogleg.py through nonlinear/common.py:118-215 (a subset of the LM
    grid here; the GPU twin runs all of it)."""
    method, kw = case
    batch_size, iters = 8, 20
    objective, variables = user_costs.regression_problem(th, multivar, batch_size=batch_size, npoints=20)
    initial_error = objective.error_metric().clone()
    cls = {"gn": th.GaussNewton, "lm": th.LevenbergMarquardt, "dogleg": th.Dogleg}[method]
    opt = cls(objective)
    assert isinstance(opt.linear_solver, th.CholeskyDenseSolver)
    opt.set_params(max_iterations=iters)
    calls = []

    def cb(o, info, delta, it):
        assert o is opt and isinstance(info, th.OptimizerInfo) and torch.is_tensor(delta) and it == len(calls)
        calls.append(it)
    with torch.no_grad():
        info = opt.optimize(track_best_solution=True, track_err_history=True, end_iter_callback=cb, **kw)
    coeffs = torch.cat([v.tensor for v in variables], dim=1)
    np.testing.assert_allclose(coeffs.numpy(), np.ones((batch_size, 5)), rtol=1e-5, atol=1e-6)   # torch.allclose defaults, as in the reference
    _check_info(info, batch_size, iters, initial_error, objective)


def test_singular_system_of_a_user_defined_cost_fails_like_the_reference(emulated):
    """nonlinear/common.py:215-280: a singular system raises RuntimeError with gradients enabled and, under no_grad, warns and reports
    status FAIL for every batch item."""
    class ZeroJacobianCost(th.CostFunction):
        def __init__(self, var, cost_weight):
            super().__init__(cost_weight)
            self.var = var
            self.register_optim_var("var")

        def dim(self):
            return 1

        def error(self):
            return torch.ones_like(self.var.tensor)

        def jacobians(self):
            return [torch.zeros(self.var.tensor.shape[0], 1, 1, dtype=self.var.dtype)], self.error()

    objective = th.Objective(dtype=torch.float64)
    objective.add(ZeroJacobianCost(th.Vector(1, name="dummy", dtype=torch.float64), th.ScaleCostWeight(torch.ones(1, dtype=torch.float64))))
    objective.update({"dummy": torch.zeros(3, 1, dtype=torch.float64)})
    opt = th.GaussNewton(objective, max_iterations=5)
    with pytest.raises(RuntimeError):
        opt.optimize(track_best_solution=True)
    with pytest.warns(RuntimeWarning):
        with torch.no_grad():
            info = opt.optimize(track_best_solution=True, track_err_history=True)
    assert (info.status == th.NonlinearOptimizerStatus.FAIL).all()


def test_public_cost_function_interface_on_built_in_costs():
    """cost_function.py:85-122 on a built-in cost function (torch restatement, no library call): error / jacobians /
    weighted_error / weighted_jacobians_error are consistent with each other and with the cost weight."""
    gen = torch.Generator().manual_seed(3)
    a = th.rand_se3(4, generator=gen, dtype=torch.float64)
    b = th.rand_se3(4, generator=gen, dtype=torch.float64)
    z = th.rand_se3(4, generator=gen, dtype=torch.float64)
    w = th.DiagonalCostWeight(torch.linspace(0.5, 3.0, 6, dtype=torch.float64).view(1, 6))
    cf = th.Between(a, b, z, w)
    with torch.no_grad():
        e = cf.error()
        jacs, e2 = cf.jacobians()
        we = cf.weighted_error()
        wj, we2 = cf.weighted_jacobians_error()
    assert e.shape == (4, 6) and [tuple(J.shape) for J in jacs] == [(4, 6, 6), (4, 6, 6)]
    torch.testing.assert_close(e, e2)
    torch.testing.assert_close(we, e * w.diagonal.tensor)
    torch.testing.assert_close(we, we2)
    for J, WJ in zip(jacs, wj):
        torch.testing.assert_close(WJ, J * w.diagonal.tensor.unsqueeze(2))
    wj3, we3 = w.weight_jacobians_and_error(jacs, e)
    torch.testing.assert_close(we3, we)
    torch.testing.assert_close(wj3[1], wj[1])


class _TorchSolveSolver(th.LinearSolver):
    """A user-defined LinearSolver (theseus/optimizer/linear/linear_solver.py:15-37): reads the linearization's AtA / Atb, applies the
    damping of dense_solver.py:38-64 and solves with torch."""

    def __init__(self, objective, linearization_cls=None, linearization_kwargs=None, **kwargs):
        super().__init__(objective, linearization_cls or th.DenseLinearization, linearization_kwargs, **kwargs)

    def solve(self, damping=None, ellipsoidal_damping=True, damping_eps=1e-8, **kwargs):
        AtA, Atb = self.linearization.AtA, self.linearization.Atb
        if damping is not None:
            d = torch.as_tensor(damping, dtype=AtA.dtype, device=AtA.device).view(-1, 1)
            diag = torch.diagonal(AtA, dim1=1, dim2=2)
            AtA = AtA + torch.diag_embed(d * diag + damping_eps if ellipsoidal_damping else d.expand_as(diag))
        return torch.linalg.solve(AtA, Atb.reshape(AtA.shape[0], -1, 1)).squeeze(2)


@pytest.mark.parametrize("method", ["gn", "lm"])
def test_user_defined_linear_solver_gives_the_library_solvers_iterates(emulated, method):
    """linear_solver_cls is a plugin point (nonlinear_least_squares.py:88-96): a user's solver over DenseLinearization must walk the same
    iterates as CholeskyDenseSolver."""
    kw = dict(damping=0.01, adaptive_damping=True, ellipsoidal_damping=True) if method == "lm" else {}
    cls = th.LevenbergMarquardt if method == "lm" else th.GaussNewton
    runs = []
    for solver_cls in (_TorchSolveSolver, th.CholeskyDenseSolver):
        objective, variables = user_costs.regression_problem(th, False, batch_size=4, npoints=20)
        opt = cls(objective, linear_solver_cls=solver_cls, max_iterations=8, abs_err_tolerance=0, rel_err_tolerance=0)
        assert isinstance(opt.linear_solver, solver_cls)
        with torch.no_grad():
            info = opt.optimize(track_err_history=True, **kw)
        runs.append((info.err_history.clone(), variables[0].tensor.clone()))
    torch.testing.assert_close(runs[0][0], runs[1][0], rtol=1e-6, atol=1e-12)
    torch.testing.assert_close(runs[0][1], runs[1][1], rtol=1e-9, atol=1e-12)


def test_user_defined_linearization_with_a_singular_system_fails_like_the_reference(emulated):
    """nonlinear/common.py:215-280 as written there: the solver's linearization replaced by a user subclass of Linearization that
    returns AtA = 0, Atb = 1."""
    class ZeroHessian(th.optimizer.Linearization):
        def _linearize_jacobian_impl(self):
            pass

        def _linearize_hessian_impl(self, _detach_hessian=False):
            B = self.objective.batch_size
            self._AtA = torch.zeros(B, self.num_cols, self.num_cols, dtype=torch.float64)
            self._Atb = torch.ones(B, self.num_cols, 1, dtype=torch.float64)

        def _ata_impl(self):
            return self._AtA

        def _atb_impl(self):
            return self._Atb

    objective, variables = user_costs.regression_problem(th, False, batch_size=2, npoints=3)
    opt = th.GaussNewton(objective, max_iterations=5)
    opt.linear_solver.linearization = ZeroHessian(objective)
    with pytest.raises(RuntimeError):
        opt.optimize(track_best_solution=True)
    with pytest.warns(RuntimeWarning):
        with torch.no_grad():
            info = opt.optimize(track_best_solution=True, track_err_history=True)
    assert (info.status == th.NonlinearOptimizerStatus.FAIL).all()


def test_linear_optimizer_solves_a_linear_problem_in_one_step(emulated):
    """theseus/optimizer/linear/linear_optimizer.py:25-83 (tests/theseus_tests/optimizer/linear/test_linear_optimizer.py pattern): one
    linearize + solve + retract; for a linear least-squares objective that is the minimiser."""
    d = torch.float64
    x, y = th.Vector(3, name="x", dtype=d), th.Vector(3, name="y", dtype=d)
    tx = th.Vector(tensor=torch.tensor([[1.0, 2.0, 3.0], [0.5, -1.0, 2.0]], dtype=d), name="tx")
    ty = th.Vector(tensor=torch.tensor([[-1.0, 0.0, 4.0], [3.0, 3.0, 3.0]], dtype=d), name="ty")
    w = th.ScaleCostWeight(torch.tensor(2.0, dtype=d))
    objective = th.Objective(dtype=d)
    objective.add(th.Difference(x, tx, w, name="px"))
    objective.add(th.Difference(y, ty, w, name="py"))
    objective.add(th.Between(x, y, th.Vector(tensor=(ty.tensor - tx.tensor), name="z"), w, name="b"))
    objective.update({"x": torch.zeros(2, 3, dtype=d), "y": torch.ones(2, 3, dtype=d)})
    opt = th.LinearOptimizer(objective, th.CholeskyDenseSolver)
    with torch.no_grad():
        info = opt.optimize()
    assert (info.status == th.LinearOptimizerStatus.CONVERGED).all()
    torch.testing.assert_close(x.tensor, tx.tensor, rtol=0, atol=1e-12)
    torch.testing.assert_close(y.tensor, ty.tensor, rtol=0, atol=1e-12)
    torch.testing.assert_close(info.best_solution["x"], tx.tensor)
    assert float(objective.error_metric().abs().max()) < 1e-20
    th.Vectorize(objective)   # by name only: the engine always evaluates per schema group
    assert objective.vectorized


def test_lie_group_checks_at_construction_follow_the_reference():
    """manifold.py:44-68,123-146 + lie_group_check.py: SE2 / SO2 storage is validated at construction (strict: ValueError, else
    normalised with a warning; silenced / skipped by the contexts and by disable_checks); the SO3 / SE3 matrix checks live in torchlie and
    are off unless its enable_checks context is active (check_contexts.py:12-44) -- the reference's observable behaviour, checked here
    against values computed with the reference (see the comment per case)."""
    from theseus_b200 import geometry as G
    d = torch.float64
    bad2 = torch.tensor([[1.0, 2.0, 0.0, 3.0], [0.0, 0.0, 0.6, 0.8]], dtype=d)
    with pytest.raises(ValueError):
        th.SE2(tensor=bad2, strict_checks=True)
    with pytest.warns(UserWarning, match="has been normalized"):
        fixed = th.SE2(tensor=bad2)
    torch.testing.assert_close(fixed.tensor, torch.tensor([[1.0, 2.0, 0.0, 1.0], [0.0, 0.0, 0.6, 0.8]], dtype=d))
    with pytest.warns(UserWarning):
        assert th.SO2(tensor=torch.zeros(1, 2, dtype=d)).tensor.tolist() == [[1.0, 0.0]]      # so2.py:193-204: zero norm -> identity
    assert torch.equal(th.SE2(tensor=bad2, disable_checks=True).tensor, bad2)
    with th.no_lie_group_check(silent=True):
        assert torch.equal(th.SE2(tensor=bad2).tensor, bad2)
        with th.enable_lie_group_check():
            with pytest.raises(ValueError):
                th.SE2(tensor=bad2, strict_checks=True)
    with th.set_lie_group_check_enabled(False, silent=True):
        assert torch.equal(th.SO2(tensor=bad2[:, 2:]).tensor, bad2[:, 2:])
    with pytest.warns(RuntimeWarning, match="checks are disabled"):
        with th.no_lie_group_check(silent=True):
            with th.set_lie_group_check_enabled(False, silent=False):
                th.SE2(tensor=bad2)
    # SE3 / SO3: not validated by default (the reference keeps a scaled rotation block as it is) ...
    X = th.rand_se3(3, generator=torch.Generator().manual_seed(0), dtype=d).tensor
    bad3 = X.clone()
    bad3[:, :, :3] *= 1.01
    assert torch.equal(th.SE3(tensor=bad3).tensor, bad3) and torch.equal(th.SE3(tensor=bad3, strict_checks=True).tensor, bad3)
    # ... and validated / normalised by SVD under torchlie's enable_checks (so3_impl.py:30-48, 1133-1141)
    with G.enable_checks():
        with pytest.raises(ValueError):
            th.SO3(tensor=bad3[:, :, :3], strict_checks=True)
        with pytest.warns(UserWarning, match="has been normalized"):
            fixed3 = th.SE3(tensor=bad3)
        assert torch.equal(th.SE3(tensor=X).tensor, X)
    torch.testing.assert_close(fixed3.tensor, X, rtol=0, atol=1e-13)      # nearest rotation of 1.01 R is R; translation untouched


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_user_defined_cost_next_to_fused_groups_reproduces_the_reference_trace(emulated, solver):
    """A pose graph whose gauge prior is a USER-DEFINED cost function (helpers.user_local_cost_cls: error / Jacobian from the library's
    group methods) while the edges stay on the fused Between kernels: the LM trace must be the reference's (pgo_small_lm), i.e. the
    generic route and the fused groups fill one linear system consistently, for both solvers."""
    from helpers import load, pgo_objective, lm_kwargs_of
    g = load("pgo_small_lm")
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g, device="cpu", user_prior=True)
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization)
    opt = th.LevenbergMarquardt(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    assert len(objective.engine().generic) == 1
    errs, deltas = [], []

    def cb(o, info, delta, it):
        errs.append(info.last_err.numpy().copy()); deltas.append(delta.numpy().copy())
    with torch.no_grad():
        opt.optimize(end_iter_callback=cb, **kw)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-8)
    for it in range(2):
        rel = np.linalg.norm(deltas[it] - g["trace_delta"][it], axis=1) / np.linalg.norm(g["trace_delta"][it], axis=1)
        assert rel.max() < 1e-5, (it, rel)
    np.testing.assert_allclose(np.stack([p.tensor.numpy() for p in poses], 0), g["poses_final"], rtol=1e-6, atol=1e-6)


def test_user_defined_manifold_as_optimisation_variable_is_refused_loudly(emulated):
    """The one plugin point that is NOT accepted: the retract kernel knows the variable kinds of thb200.h."""
    class MyManifold(th.Manifold):
        def dof(self):
            return self.tensor.shape[1]

    objective, variables = user_costs.regression_problem(th, False, batch_size=2, npoints=2)
    v = MyManifold(torch.zeros(2, 5, dtype=torch.float64), name="mine")
    Cost = user_costs.squared_fit_cost_cls(th)
    bad = th.Objective(dtype=torch.float64)
    bad.add(Cost([v], th.ScaleCostWeight(torch.ones(1, dtype=torch.float64)), th.Variable(torch.ones(2, 5, dtype=torch.float64), name="p"),
                 th.Variable(torch.ones(2, 1, dtype=torch.float64), name="t"), name="c"))
    with pytest.raises(NotImplementedError, match="user-defined Manifold"):
        th.GaussNewton(bad).optimize()


def test_masked_jacobians_and_masked_variables():
    """core/cost_function.py:37-55, core/variable.py:134-148."""
    gen = torch.Generator().manual_seed(2)
    d = torch.float64
    a, b, z = (th.rand_se3(5, generator=gen, dtype=d) for _ in range(3))
    cf = th.Between(a, b, z, th.ScaleCostWeight(torch.ones(1, dtype=d)))
    mask = torch.tensor([True, False, True, True, False])
    with torch.no_grad():
        full_j, full_e = cf.jacobians()
        mj, me = th.masked_jacobians(cf, mask)
    assert me.shape == full_e.shape and mj[0].shape == full_j[0].shape
    torch.testing.assert_close(me[mask], full_e[mask])
    torch.testing.assert_close(mj[1][mask], full_j[1][mask])
    assert float(me[~mask].abs().max()) == 0.0 and float(mj[0][~mask].abs().max()) == 0.0
    assert a.tensor.shape[0] == 5
    with th.masked_variables([a, b], mask):
        assert a.tensor.shape[0] == 3 and b.tensor.shape[0] == 3
    assert a.tensor.shape[0] == 5
