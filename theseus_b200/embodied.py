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

    def _generic_unweighted(self, optim_tensors, differentiable: bool = False):
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
            return jacs, err
        return [j.detach() for j in jacs], err.detach()

    def schema(self):
        return None, []


class QuasiStaticPushingPlanar(CostFunction):
    """theseus/embodied/motionmodel/quasi_static_pushing_planar.py:19-297 (quasi-static pushing model of the tactile example, Zhou et
    al. 2017): object poses obj1, obj2 and end-effector poses eff1, eff2 (SE2) at consecutive times, aux c_square;
        e = D V - Vp,  V = [R2^T (t_o2 - t_o1), theta(o1^-1 o2)],  Vp = [R2^T (t_e2 - t_e1), 0],
        D = [[1, 0, -py], [0, 1, px], [-py, px, -c^2]],  (px, py) = R2^T (t_e2 - t_o2).
    Torch path (torch.func Jacobians + tangent-space projection = the reference's chained analytic Jacobians)."""

    def __init__(self, obj1, obj2, eff1, eff2, c_square, cost_weight: CostWeight, name: Optional[str] = None):
        from .geometry import Variable, as_variable
        super().__init__(cost_weight, name=name)
        self.obj1, self.obj2, self.eff1, self.eff2 = obj1, obj2, eff1, eff2
        self.register_optim_vars(["obj1", "obj2", "eff1", "eff2"])
        c_square = c_square if isinstance(c_square, Variable) else as_variable(c_square)
        if c_square.tensor.dtype != obj1.dtype:
            c_square.tensor = c_square.tensor.to(obj1.dtype)
        if c_square.tensor.squeeze().ndim > 1:
            raise ValueError("dt must be a 0-D or 1-D tensor.")
        c_square.tensor = c_square.tensor.view(-1, 1)
        self.c_square = c_square
        self.register_aux_vars(["c_square"])

    def dim(self) -> int:
        return 3

    def _torch_error(self, optim_tensors, aux_tensors):
        import torch
        o1, o2, e1, e2 = optim_tensors
        c2 = aux_tensors[0].view(-1)
        cos2, sin2 = o2[..., 2], o2[..., 3]

        def unrot(v):  # R2^T v
            return torch.stack((cos2 * v[..., 0] + sin2 * v[..., 1], -sin2 * v[..., 0] + cos2 * v[..., 1]), dim=-1)

        p = unrot(e2[..., :2] - o2[..., :2])
        v = unrot(o2[..., :2] - o1[..., :2])
        # theta of o1^-1 o2 (se2.py:318-339): cos = c1 c2 + s1 s2, sin = c1 s2 - s1 c2
        omega = torch.atan2(o1[..., 2] * sin2 - o1[..., 3] * cos2, o1[..., 2] * cos2 + o1[..., 3] * sin2)
        vp = unrot(e2[..., :2] - e1[..., :2])
        px, py = p[..., 0], p[..., 1]
        ex = v[..., 0] - py * omega - vp[..., 0]
        ey = v[..., 1] + px * omega - vp[..., 1]
        et = -py * v[..., 0] + px * v[..., 1] - c2 * omega
        return torch.stack((ex, ey, et), dim=-1)

    def schema(self):
        return None, []


class EffectorObjectContactPlanar(CostFunction):
    """theseus/embodied/collision/eff_obj_contact.py:21-126 (+ SignedDistanceField2D.signed_distance, collision/signed_distance_field.py:
    163-241): the end effector (a disc of radius eff_radius at eff.xy) touches the object whose signed distance field is given in the
    object frame:  e = | sdf(R_obj^T (t_eff - t_obj)) - eff_radius |,  dim 1.  sdf = bilinear interpolation of sdf_data [Bs, rows, cols]
    (cell (r,c) at origin + (c, r) * cell_size), 0 outside the grid.  Torch path; autograd of the bilinear form is exactly the
    reference's analytic gradient, the sign flip for dist < radius is d|.|."""

    def __init__(self, obj, eff, sdf_origin, sdf_data, sdf_cell_size, eff_radius, cost_weight: CostWeight, name: Optional[str] = None,
                 use_huber_loss: bool = False):
        import torch
        from .geometry import Point2, Variable, as_variable
        if use_huber_loss:
            raise NotImplementedError("Jacobians for huber loss are not yet implemented.")  # same as the reference (eff_obj_contact.py:49-52)
        super().__init__(cost_weight, name=name)
        self.obj, self.eff = obj, eff
        self.sdf_origin = sdf_origin if isinstance(sdf_origin, Point2) else Point2(tensor=sdf_origin)
        self.sdf_data = as_variable(sdf_data)
        if self.sdf_data.tensor.ndim != 3:
            raise ValueError("Argument sdf_data to SignedDistanceField2D must be a batch of matrices.")
        if isinstance(sdf_cell_size, Variable):
            self.sdf_cell_size = sdf_cell_size
        else:
            self.sdf_cell_size = Variable((sdf_cell_size if torch.is_tensor(sdf_cell_size) else torch.tensor(float(sdf_cell_size))).view(-1, 1))
        self.eff_radius = as_variable(eff_radius)
        if self.eff_radius.tensor.squeeze().ndim > 1:
            raise ValueError("eff_radius must be a 0-D or 1-D tensor.")
        self.eff_radius.tensor = self.eff_radius.tensor.view(-1, 1)
        for v in (self.sdf_cell_size, self.eff_radius, self.sdf_data):
            if v.tensor.dtype != obj.dtype:
                v.tensor = v.tensor.to(obj.dtype)
        self.register_optim_vars(["obj", "eff"])
        self.register_aux_vars(["sdf_origin", "sdf_data", "sdf_cell_size", "eff_radius"])

    def dim(self) -> int:
        return 1

    def _torch_error(self, optim_tensors, aux_tensors):
        import torch
        o, e = optim_tensors
        origin, data, cell, radius = aux_tensors
        cell, radius = cell.view(-1), radius.view(-1)
        dx, dy = e[..., 0] - o[..., 0], e[..., 1] - o[..., 1]
        px = o[..., 2] * dx + o[..., 3] * dy       # eff position in the object frame (SE2.transform_to)
        py = -o[..., 3] * dx + o[..., 2] * dy
        nrows, ncols = data.shape[-2], data.shape[-1]
        oob = (px < origin[..., 0]) | (px > origin[..., 0] + (ncols - 1.0) * cell) | (py < origin[..., 1]) | (py > origin[..., 1] + (nrows - 1.0) * cell)
        col, row = (px - origin[..., 0]) / cell, (py - origin[..., 1]) / cell
        lr, lc = torch.floor(row), torch.floor(col)
        hr, hc = lr + 1.0, lc + 1.0
        lri, lci = lr.long().clamp(0, nrows - 1), lc.long().clamp(0, ncols - 1)
        hri, hci = hr.long().clamp(0, nrows - 1), hc.long().clamp(0, ncols - 1)
        bi = torch.arange(data.shape[0], device=data.device)
        g = lambda r, c: data[bi, r, c]
        dist = (hr - row) * (hc - col) * g(lri, lci) + (row - lr) * (hc - col) * g(hri, lci) \
            + (hr - row) * (col - lc) * g(lri, hci) + (row - lr) * (col - lc) * g(hri, hci)
        dist = torch.where(oob, torch.zeros_like(dist), dist)    # sdf_boundary_value = 0 (signed_distance_field.py:26)
        return (dist - radius).abs().unsqueeze(-1)

    def schema(self):
        return None, []
