"""User-defined cost functions / cost weights written against the reference's plugin contract (theseus/core/cost_function.py:64-149,
cost_weight.py:20-55: subclass, register the variables, implement error() / jacobians() / dim()) -- problem builders shared by the CPU
tests on the host emulation (tests/test_user_defined_costs.py) and the GPU tests (tests/test_gpu_zz_first_run.py).

The two problems restate the reference's own plugin-level tests:
  * the hand-written 6 x 10 linear system of tests/theseus_tests/optimizer/linearization_test_utils.py:122-196 (three cost functions of
    dims 1, 2, 3 over four variables of sizes 1..4, identity-times-k weights, column order v4 v3 v2 v1);
  * the quadratic-regression problem of tests/theseus_tests/optimizer/nonlinear/common.py:14-215: residuals
    (sum_i b_i p_i)^2 - (sum_i 1 * p_i)^2 with p = [point, 1], one cost function per data point, coefficients either one Vector(nvars)
    or nvars Vector(1)s; every GN / LM variant must recover b = 1."""
import torch


def weighted_sum_cost_cls(th):
    class WeightedSumCost(th.CostFunction):
        """e_i = (i + 1) * sum_j (j + 1) * sum(var_j);  d e_i / d var_j = (i + 1)(j + 1) * ones  (linearization_test_utils.py:58-90)."""

        def __init__(self, optim_vars, cost_weight, dim, name=None):
            super().__init__(cost_weight, name=name)
            for j, v in enumerate(optim_vars):
                setattr(self, f"optim_var_{j}", v)
                self.register_optim_var(f"optim_var_{j}")
            self._dim = dim

        def dim(self):
            return self._dim

        def error(self):
            total = sum((j + 1) * v.tensor.sum(dim=1, keepdim=True) for j, v in enumerate(self.optim_vars))       # [B, 1]
            rows = torch.arange(1, self._dim + 1, dtype=total.dtype, device=total.device).view(1, -1)
            return total * rows

        def jacobians(self):
            t0 = self.optim_vars[0].tensor
            rows = torch.arange(1, self._dim + 1, dtype=t0.dtype, device=t0.device).view(1, -1, 1)
            jacs = [(j + 1) * rows * torch.ones(t0.shape[0], self._dim, v.dof(), dtype=t0.dtype, device=t0.device)
                    for j, v in enumerate(self.optim_vars)]
            return jacs, self.error()
    return WeightedSumCost


def matrix_weight_cls(th):
    class MatrixWeight(th.CostWeight):
        """sqrt-information MATRIX mult * I (linearization_test_utils.py:98-119): a weight kind the library has no kernel for."""

        def __init__(self, dim, mult, dtype, name=None):
            super().__init__(name=name)
            self.sqrt = th.Variable(torch.eye(dim, dtype=dtype).unsqueeze(0) * mult, name=f"{self.name}__sqrt")
            self.register_aux_var("sqrt")

        def weight_error(self, error):
            return torch.matmul(self.sqrt.tensor, error.unsqueeze(2)).squeeze(2)

        def weight_jacobians_and_error(self, jacobians, error):
            return [torch.matmul(self.sqrt.tensor, J) for J in jacobians], self.weight_error(error)
    return MatrixWeight


def mock_linear_system(th, device="cpu", dtype=torch.float64, batch_size=4):
    """(objective, ordering, A [B,6,10], b [B,6]) of linearization_test_utils.py:122-196 with th.Vector variables."""
    Cost, W = weighted_sum_cost_cls(th), matrix_weight_cls(th)
    v = {k: th.Vector(k, name=f"v{k}", dtype=dtype) for k in (1, 2, 3, 4)}
    objective = th.Objective(dtype=dtype)
    objective.add(Cost([v[1], v[2]], W(1, 1.0, dtype, name="cov1"), 1, name="f1"))
    objective.add(Cost([v[1], v[3]], W(2, 2.0, dtype, name="cov2"), 2, name="f2"))
    objective.add(Cost([v[2], v[4]], W(3, 3.0, dtype, name="cov3"), 3, name="f3"))
    objective.to(device)
    ordering = th.VariableOrdering(objective, default_order=False)
    ordering.extend([v[4], v[3], v[2], v[1]])
    objective.update({f"v{k}": k * torch.ones(batch_size, k, dtype=dtype, device=device) for k in (1, 2, 3, 4)})
    # cost function k over (v_i, v_j): every row r of the block is k * (r + 1) * [1 * ones(i) | 2 * ones(j)], error = (r + 1) (i^2 + 2 j^2)
    A = torch.zeros(6, 10, dtype=dtype)
    col = {4: 0, 3: 4, 2: 7, 1: 9}
    b, row = [], 0
    for k, (i, j) in ((1, (1, 2)), (2, (1, 3)), (3, (2, 4))):
        for r in range(k):
            A[row, col[i]:col[i] + i] = k * (r + 1) * 1.0
            A[row, col[j]:col[j] + j] = k * (r + 1) * 2.0
            b.append(-k * (r + 1) * (i * i + 2.0 * j * j))
            row += 1
    A = A.unsqueeze(0).repeat(batch_size, 1, 1)
    b = torch.tensor(b, dtype=dtype).unsqueeze(0).repeat(batch_size, 1)
    return objective, ordering, A, b


def squared_fit_cost_cls(th):
    class SquaredFitCost(th.CostFunction):
        """e = (sum_i b_i p_i)^2 - target,  J = 2 (sum_i b_i p_i) p  (nonlinear/common.py:14-89); coefficients in one Vector or one per
        Vector(1)."""

        def __init__(self, optim_vars, cost_weight, point, target, name=None):
            super().__init__(cost_weight, name=name)
            for j, v in enumerate(optim_vars):
                setattr(self, f"optim_var_{j}", v)
                self.register_optim_var(f"optim_var_{j}")
            self.point, self.target = point, target
            self.register_aux_vars(["point", "target"])

        def dim(self):
            return 1

        def _z(self):
            coeffs = torch.cat([v.tensor for v in self.optim_vars], dim=1)
            return (self.point.tensor * coeffs).sum(dim=1, keepdim=True)

        def error(self):
            return self._z() ** 2 - self.target.tensor

        def jacobians(self):
            grad = (2.0 * self._z() * self.point.tensor).unsqueeze(1)          # [B, 1, nvars]
            if len(self.optim_vars) == 1:
                return [grad], self.error()
            return [grad[:, :, j:j + 1] for j in range(len(self.optim_vars))], self.error()
    return SquaredFitCost


def regression_problem(th, multivar, device="cpu", dtype=torch.float64, nvars=5, npoints=50, batch_size=32, seed=0):
    """nonlinear/common.py:118-215: (objective, variables); true coefficients all 1, initial value b_i = i."""
    Cost = squared_fit_cost_cls(th)
    gen = torch.Generator().manual_seed(seed)
    if multivar:
        variables = [th.Vector(1, name=f"coeff{i}", dtype=dtype) for i in range(nvars)]
    else:
        variables = [th.Vector(nvars, name="coefficients", dtype=dtype)]
    w = th.ScaleCostWeight(torch.ones(1, dtype=dtype))
    objective = th.Objective(dtype=dtype)
    for k in range(npoints):
        p = torch.cat([torch.randn(batch_size, nvars - 1, generator=gen, dtype=dtype), torch.ones(batch_size, 1, dtype=dtype)], dim=1)
        target = p.sum(dim=1, keepdim=True) ** 2
        objective.add(Cost(variables, w, th.Variable(p, name=f"point_{k}"), th.Variable(target, name=f"target_{k}"), name=f"residual_point_{k}"))
    objective.to(device)
    if multivar:
        objective.update({f"coeff{i}": i * torch.ones(batch_size, 1, dtype=dtype, device=device) for i in range(nvars)})
    else:
        objective.update({"coefficients": torch.arange(nvars, dtype=dtype, device=device).repeat(batch_size, 1)})
    return objective, variables
