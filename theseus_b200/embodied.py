"""`th.eb` namespace of the reference (theseus/embodied/__init__.py): Between / Local / Reprojection have fused CUDA schemas;
MovingFrameBetween runs on the torch path (torch.func Jacobians + tangent-space projection, like an AutoDiffCostFunction)."""
from typing import Optional

from .core import Between, CostFunction, CostWeight, Difference, Local, Reprojection  # noqa: F401
from .geometry import LieGroup


class MovingFrameBetween(CostFunction):
    """theseus/embodied/measurements/moving_frame_between.py:14-77:
        e = log(Z^-1 ((F1^-1 P1)^-1 (F2^-1 P2)))        optim vars: frame1, frame2, pose1, pose2 (SE2 or SE3); aux: measurement.
    No fused kernel (the engine's schemas hold at most two variables): torch path.  NOTE the reference's Jacobians are those of the
    group-valued D = (F1^-1 P1)^-1 (F2^-1 P2) in ITS tangent space (moving_frame_between.py:46-65 chains the `between` Jacobians and
    stops there) -- the d log factor of the final `measurement.local(D)` is not applied.  For drop-in parity the same quantity is
    computed here: Euclidean torch.func Jacobian of D, input side projected like every AutoDiff Jacobian, output side converted from a
    velocity dD to tangent coordinates vee(D^-1 dD)  (tests/test_torch_restatements.py compares with the reference's values)."""

    def __init__(self, frame1: LieGroup, frame2: LieGroup, pose1: LieGroup, pose2: LieGroup, measurement: LieGroup,
                 cost_weight: CostWeight, name: Optional[str] = None):
        if len(set(x.__class__.__name__ for x in (frame1, frame2, pose1, pose2, measurement))) > 1:
            raise ValueError("Inconsistent types between input variables.")
        super().__init__(cost_weight, name=name)
        self.frame1, self.frame2, self.pose1, self.pose2 = frame1, frame2, pose1, pose2
        self.register_optim_vars(["frame1", "frame2", "pose1", "pose2"])
        self.measurement = measurement
        self.register_aux_vars(["measurement"])

    def dim(self) -> int:
        return self.frame1.dof()

    def _torch_error(self, optim_tensors, aux_tensors):
        from . import lie_torch
        k = self.frame1.KIND
        f1, f2, p1, p2 = optim_tensors
        d = lie_torch.between(k, lie_torch.between(k, f1, p1), lie_torch.between(k, f2, p2))
        return lie_torch.local(k, aux_tensors[0], d)

    def _torch_frame_diff(self, optim_tensors):
        from . import lie_torch
        k = self.frame1.KIND
        f1, f2, p1, p2 = optim_tensors
        return lie_torch.between(k, lie_torch.between(k, f1, p1), lie_torch.between(k, f2, p2))

    def generic_jacobians_error(self, optim_tensors, differentiable: bool = False):
        import torch
        from torch.func import jacrev, vmap
        from . import lie_torch
        k = self.frame1.KIND
        aux = self.measurement.tensor
        B = max([t.shape[0] for t in optim_tensors] + [aux.shape[0]])
        ex = lambda t: t if t.shape[0] == B else t.expand((B,) + tuple(t.shape[1:]))
        opt_t = tuple(ex(t) for t in optim_tensors)

        def one(o):
            return self._torch_frame_diff(tuple(x.unsqueeze(0) for x in o))[0]

        with torch.enable_grad():
            D = self._torch_frame_diff(opt_t)
            dD = vmap(jacrev(one))(opt_t)                      # per variable: [B, *group_shape(out), *group_shape(in)]
            err = lie_torch.local(k, ex(aux), D)
        gs = D.ndim - 1                                         # 1 for SE2 storage [4], 2 for SE3 storage [3,4]
        jacs = []
        for v, t, J in zip(self.optim_vars, opt_t, dD):
            Jin = type(v).project_tensor(t, J.reshape(B, -1, *t.shape[1:]))          # input side -> tangent: [B, prod(out), dof]
            Jin = Jin.reshape(B, *D.shape[1:], Jin.shape[-1])                         # [B, *out, dof]
            Jin = Jin.movedim(-1, 1)                                                  # [B, dof, *out] : one velocity dD per column
            jacs.append(lie_torch.velocity_to_tangent(k, D, Jin).transpose(1, 2))     # [B, dof_out, dof_in]
        if differentiable:
            return self._weight(err, jacs)
        return self._weight(err.detach(), [j.detach() for j in jacs])

    def schema(self):
        return None, []
