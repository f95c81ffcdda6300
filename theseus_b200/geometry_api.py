"""The rest of the reference's public interface of the geometry classes, installed on theseus_b200.geometry's classes at import:

  * TAPE ROUTE of the group operations: exp_map / log_map / inverse / compose / adjoint are CUDA kernels (csrc/thb_lie_ops.cu) and
    therefore not differentiable; when an operand is on the autograd tape or inside a torch.func transform (a user's
    AutoDiffCostFunction err_fn, which receives typed variables exactly like the reference's: cost_function.py:283-316), the same closed
    forms run as differentiable torch ops (lie_torch.py) on whatever device the tensor lives on.  Plain tensors keep the kernels.
  * accessors / actions / conversions that are a handful of tensor operations in the reference too (no kernel): rotation, translation,
    transform_to / transform_from (+ Jacobians, se3_impl.py:757-835, se2.py:399-468), rotate / unrotate, to_matrix, hat / vee, theta, xy,
    x / y / z, quaternion conversions, the Vector arithmetic of geometry/vector.py, rand / randn.

References: theseus/geometry/{se3,so3,se2,so2,vector,point_types}.py, torchlie/functional/{se3,so3}_impl.py."""
from typing import List, Optional, Union

import torch

from . import lie_torch
from .geometry import SE2, SE3, SO2, SO3, LieGroup, Point2, Point3, Variable, Vector


def on_tape(*tensors: torch.Tensor) -> bool:
    """True if any operand needs differentiable torch ops: inside vmap / jacrev, or requiring grad with grad mode on."""
    F = torch._C._functorch
    for t in tensors:
        if F.is_batchedtensor(t) or F.is_gradtrackingtensor(t) or (t.requires_grad and torch.is_grad_enabled()):
            return True
    return False


def _t(x) -> torch.Tensor:
    return x.tensor if isinstance(x, Variable) else x


def _skew(w: torch.Tensor) -> torch.Tensor:
    z = torch.zeros_like(w[..., 0])
    return torch.stack((torch.stack((z, -w[..., 2], w[..., 1]), -1), torch.stack((w[..., 2], z, -w[..., 0]), -1),
                        torch.stack((-w[..., 1], w[..., 0], z), -1)), -2)


# ---------------------------------------------------------------------------------------------------------------- tape route
_EXP = {0: lie_torch.se3_exp, 1: lambda w: lie_torch._so3_exp_parts(w)[0], 3: lie_torch.se2_exp, 4: lie_torch.so2_exp}
_LOG = {0: lie_torch.se3_log, 1: lambda R: lie_torch.so3_log(R)[0], 3: lie_torch.se2_log, 4: lie_torch.so2_log}
_INV = {0: lie_torch.se3_inverse, 1: lambda R: R.transpose(-1, -2), 3: lie_torch.se2_inverse, 4: lie_torch.so2_inverse}
_MUL = {0: lie_torch.se3_compose, 1: lambda A, B: A @ B, 3: lie_torch.se2_compose, 4: lie_torch.so2_compose}


def _adjoint_torch(kind: int, g: torch.Tensor) -> torch.Tensor:
    if kind == 0:    # se3_impl.py:531-538: [[R, [t]x R], [0, R]]
        R, t = g[..., :3], g[..., 3]
        top = torch.cat((R, _skew(t) @ R), dim=-1)
        return torch.cat((top, torch.cat((torch.zeros_like(R), R), dim=-1)), dim=-2)
    if kind == 1:
        return g
    if kind == 3:    # se2.py:309-316: [[R, (y, -x)^T], [0, 0, 1]]
        c, s, x, y = g[..., 2], g[..., 3], g[..., 0], g[..., 1]
        z, o = torch.zeros_like(c), torch.ones_like(c)
        return torch.stack((torch.stack((c, -s, y), -1), torch.stack((s, c, -x), -1), torch.stack((z, z, o), -1)), -2)
    return torch.ones(g.shape[0], 1, 1, dtype=g.dtype, device=g.device)


_SE2_NEAR_ZERO = {torch.float32: 3e-2, torch.float64: 1e-6}     # theseus/global_params.py:46-59


def _exp_with_jacobian(cls, tangent_vector: torch.Tensor, jacobians: List[torch.Tensor]):
    """LieGroup.exp_map(tangent, jacobians=[]) (lie_group.py:84-93): the group element and d exp / d tangent (right Jacobian).
    SO3 / SE3: one kernel for both (thb_so3_jexp / thb_se3_jexp: so3_impl.py:270-320, se3_impl.py:225-330); SE2 / SO2: the closed form
    of se2.py:239-300 / so2.py:222-234 in the few tensor operations the reference uses."""
    from . import _lib
    from .geometry import _require_cuda, _sfx
    LieGroup._check_jacobians_list(jacobians)
    kind = cls.KIND
    if kind in (0, 1):
        if on_tape(tangent_vector):
            raise NotImplementedError(f"{cls.__name__}.exp_map(..., jacobians=[...]) of a tensor on the autograd tape")
        _require_cuda(tangent_vector, f"{cls.__name__}.exp_map")
        t = tangent_vector.contiguous()
        n, gshape, name = (6, (3, 4), "se3") if kind == 0 else (3, (3, 3), "so3")
        out = torch.empty((t.shape[0],) + gshape, dtype=t.dtype, device=t.device)
        J = torch.empty(t.shape[0], n, n, dtype=t.dtype, device=t.device)
        _lib.check(getattr(_lib.load(), f"thb_{name}_jexp_{_sfx(t)}")(_lib.ptr(t), _lib.ptr(out), _lib.ptr(J), t.shape[0], _lib.stream_ptr()), f"{name}_jexp")
        jacobians.append(J)
        return cls(tensor=out, disable_checks=True)
    group = cls.exp_map(tangent_vector)         # SE2 / SO2: the element first (it may refuse the tensor), then the closed-form Jacobian
    if kind == 4:
        jacobians.append(torch.ones(tangent_vector.shape[0], 1, 1, dtype=tangent_vector.dtype, device=tangent_vector.device))
        return group
    u, theta = tangent_vector[:, :2], tangent_vector[:, 2]
    cosine, sine = theta.cos(), theta.sin()
    small = theta.abs() < _SE2_NEAR_ZERO[tangent_vector.dtype]
    one = torch.ones((), dtype=theta.dtype, device=theta.device)
    theta2, theta3 = theta ** 2, theta ** 3
    theta_nz, theta2_nz = torch.where(small, one, theta), torch.where(small, one, theta2)
    sbt = torch.where(small, 1 - theta2 / 6, sine / theta_nz)
    cm1bt = torch.where(small, -theta / 2 + theta3 / 24, (cosine - 1) / theta_nz)
    tmsbt2 = torch.where(small, theta - theta3 / 120, (theta - sine) / theta2_nz)
    cm1bt2 = torch.where(small, -0.5 + theta2 / 24, (cosine - 1) / theta2_nz)
    z, o = torch.zeros_like(theta), torch.ones_like(theta)
    jacobians.append(torch.stack((torch.stack((sbt, -cm1bt, tmsbt2 * u[:, 0] + cm1bt2 * u[:, 1]), -1),
                                  torch.stack((cm1bt, sbt, tmsbt2 * u[:, 1] - cm1bt2 * u[:, 0]), -1), torch.stack((z, z, o), -1)), -2))
    return group


def _install_tape_route(cls):
    kind = cls.KIND
    k_exp, k_log, k_inv, k_mul, k_adj = cls.exp_map, cls.log_map, cls.inverse, cls.compose, getattr(cls, "adjoint", None)

    def exp_map(tangent_vector: torch.Tensor, jacobians: Optional[List[torch.Tensor]] = None):
        if jacobians is not None:
            return _exp_with_jacobian(cls, tangent_vector, jacobians)
        if on_tape(tangent_vector):
            return cls(tensor=_EXP[kind](tangent_vector), disable_checks=True)
        return k_exp(tangent_vector)

    def log_map(self, jacobians: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        if jacobians is None and on_tape(self.tensor):
            return _LOG[kind](self.tensor)
        return k_log(self, jacobians)

    def inverse(self):
        if on_tape(self.tensor):
            return cls(tensor=_INV[kind](self.tensor), disable_checks=True)
        return k_inv(self)

    def compose(self, other):
        if on_tape(self.tensor, other.tensor):
            return cls(tensor=_MUL[kind](self.tensor, other.tensor), disable_checks=True)
        return k_mul(self, other)

    def adjoint(self) -> torch.Tensor:
        if on_tape(self.tensor) or k_adj is None:
            return _adjoint_torch(kind, self.tensor)
        return k_adj(self)

    cls.exp_map = staticmethod(exp_map)
    cls.log_map, cls.inverse, cls.compose, cls.adjoint = log_map, inverse, compose, adjoint


# ---------------------------------------------------------------------------------------------------------------- SE3 / SO3
def _se3_rotation(self) -> SO3:
    return SO3(tensor=self.tensor[..., :3], disable_checks=True)


def _se3_translation(self) -> Point3:
    return Point3(tensor=self.tensor[..., 3])


def _point_tensor(point, n: int, who: str) -> torch.Tensor:
    p = _t(point)
    if p.ndim != 2 or p.shape[1] != n:
        raise ValueError(f"{who}: points must have shape [batch, {n}]")
    return p


def _se3_transform_from(self, point: Union[Point3, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point3:
    """se3.py:249-256, se3_impl.py:757-777: R p + t;  d/dg = [R, -R [p]x], d/dp = R."""
    p = _point_tensor(point, 3, "SE3.transform_from")
    R, t = self.tensor[..., :3], self.tensor[..., 3]
    ret = (R @ p.unsqueeze(-1)).squeeze(-1) + t
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        Rb = R.expand(ret.shape[0], 3, 3)
        jacobians.extend([torch.cat((Rb, -(Rb @ _skew(p))), dim=-1), Rb])
    return Point3(tensor=ret)


def _se3_transform_to(self, point: Union[Point3, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point3:
    """se3.py:258-276, se3_impl.py:788-835: R^T (p - t);  d/dg = [-I, [ret]x], d/dp = R^T."""
    p = _point_tensor(point, 3, "SE3.transform_to")
    R, t = self.tensor[..., :3], self.tensor[..., 3]
    ret = (R.transpose(-1, -2) @ (p - t).unsqueeze(-1)).squeeze(-1)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        eye = torch.eye(3, dtype=ret.dtype, device=ret.device).expand(ret.shape[0], 3, 3)
        jacobians.extend([torch.cat((-eye, _skew(ret)), dim=-1), R.transpose(-1, -2).expand(ret.shape[0], 3, 3)])
    return Point3(tensor=ret)


def _se3_to_matrix(self) -> torch.Tensor:
    g = self.tensor
    last = torch.zeros(g.shape[0], 1, 4, dtype=g.dtype, device=g.device)
    last[..., 3] = 1
    return torch.cat((g, last), dim=-2)


def _se3_hat(tangent_vector: torch.Tensor) -> torch.Tensor:
    """se3_impl.py _hat_impl: [[ [w]x, v ], [0, 0]] for the tangent order [v, w]."""
    v, w = tangent_vector[..., :3], tangent_vector[..., 3:]
    top = torch.cat((_skew(w), v.unsqueeze(-1)), dim=-1)
    return torch.cat((top, torch.zeros_like(top[..., :1, :])), dim=-2)


def _se3_vee(matrix: torch.Tensor) -> torch.Tensor:
    R = matrix[..., :3, :3]
    w = 0.5 * torch.stack((R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]), dim=-1)
    return torch.cat((matrix[..., :3, 3], w), dim=-1)


def _so3_hat(tangent_vector: torch.Tensor) -> torch.Tensor:
    return _skew(tangent_vector)


def _so3_vee(matrix: torch.Tensor) -> torch.Tensor:
    return 0.5 * torch.stack((matrix[..., 2, 1] - matrix[..., 1, 2], matrix[..., 0, 2] - matrix[..., 2, 0], matrix[..., 1, 0] - matrix[..., 0, 1]), dim=-1)


def _so3_rotate(self, point: Union[Point3, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point3:
    """so3.py:287-294, so3_impl.py _transform / _jtransform: R p;  d/dR = -R [p]x, d/dp = R."""
    p = _point_tensor(point, 3, "SO3.rotate")
    R = self.tensor
    ret = (R @ p.unsqueeze(-1)).squeeze(-1)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        Rb = R.expand(ret.shape[0], 3, 3)
        jacobians.extend([-(Rb @ _skew(p)), Rb])
    return Point3(tensor=ret)


def _so3_unrotate(self, point: Union[Point3, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point3:
    """so3.py:296-303: R^T p;  d/dR = [ret]x, d/dp = R^T."""
    p = _point_tensor(point, 3, "SO3.unrotate")
    R = self.tensor
    ret = (R.transpose(-1, -2) @ p.unsqueeze(-1)).squeeze(-1)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        jacobians.extend([_skew(ret), R.transpose(-1, -2).expand(ret.shape[0], 3, 3)])
    return Point3(tensor=ret)


def _so3_to_quaternion(self) -> torch.Tensor:
    """so3.py to_quaternion ([w, x, y, z]): from the rotation vector, q = [cos(theta/2), sin(theta/2) axis], with the half-angle series
    near zero."""
    w = SO3(tensor=self.tensor, disable_checks=True).log_map() if not on_tape(self.tensor) and self.tensor.is_cuda else lie_torch.so3_log(self.tensor)[0]
    theta = w.norm(dim=-1, keepdim=True)
    half = 0.5 * theta
    small = theta < 1e-6
    k = torch.where(small, 0.5 - theta ** 2 / 48, torch.sin(half) / torch.where(small, torch.ones_like(theta), theta))
    return torch.cat((torch.cos(half), k * w), dim=-1)


def _unit_quaternion_to_so3(quaternion: torch.Tensor) -> SO3:
    """so3.py unit_quaternion_to_SO3: [w, x, y, z] -> rotation matrix."""
    if quaternion.ndim == 1:
        quaternion = quaternion.unsqueeze(0)
    q = quaternion / quaternion.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack((torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)), -1),
                     torch.stack((2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)), -1),
                     torch.stack((2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), -1)), -2)
    return SO3(tensor=R, disable_checks=True)


def _x_y_z_unit_quaternion_to_se3(x_y_z_quaternion: torch.Tensor) -> SE3:
    """se3.py x_y_z_unit_quaternion_to_SE3: [x, y, z, qw, qx, qy, qz]."""
    if x_y_z_quaternion.ndim == 1:
        x_y_z_quaternion = x_y_z_quaternion.unsqueeze(0)
    R = _unit_quaternion_to_so3(x_y_z_quaternion[:, 3:]).tensor
    return SE3(tensor=torch.cat((R, x_y_z_quaternion[:, :3, None]), dim=2), disable_checks=True)


def _se3_to_x_y_z_quaternion(self) -> torch.Tensor:
    return torch.cat((self.tensor[..., 3], _so3_to_quaternion(_se3_rotation(self))), dim=-1)


# ---------------------------------------------------------------------------------------------------------------- SE2 / SO2
def _rot2(p: torch.Tensor, c: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    return torch.stack((c * p[..., 0] - s * p[..., 1], s * p[..., 0] + c * p[..., 1]), dim=-1)


def _se2_rotation(self) -> SO2:
    return SO2(tensor=self.tensor[:, 2:], disable_checks=True)


def _se2_translation(self) -> Point2:
    return Point2(tensor=self.tensor[:, :2])


def _se2_theta(self, jacobians: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """se2.py:130-137."""
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        J = self.tensor.new_zeros(self.tensor.shape[0], 1, 3)
        J[:, 0, 2] = 1
        jacobians.append(J)
    return torch.atan2(self.tensor[:, 3], self.tensor[:, 2]).unsqueeze(1)


def _se2_xy(self, jacobians: Optional[List[torch.Tensor]] = None) -> Point2:
    """se2.py:143-152: d xy / d tangent = [R, 0]."""
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        c, s = self.tensor[:, 2], self.tensor[:, 3]
        z = torch.zeros_like(c)
        jacobians.append(torch.stack((torch.stack((c, -s, z), -1), torch.stack((s, c, z), -1)), -2))
    return Point2(tensor=self.tensor[:, :2])


def _se2_transform_to(self, point: Union[Point2, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point2:
    """se2.py:399-432: R^T (p - t);  d/dg = [[-1, 0, ret_y], [0, -1, -ret_x]], d/dp = R^T."""
    p = _point_tensor(point, 2, "SE2.transform_to")
    c, s = self.tensor[:, 2], self.tensor[:, 3]
    ret = _rot2(p - self.tensor[:, :2], c, -s)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        o, z = torch.ones_like(ret[:, 0]), torch.zeros_like(ret[:, 0])
        Jg = torch.stack((torch.stack((-o, z, ret[:, 1]), -1), torch.stack((z, -o, -ret[:, 0]), -1)), -2)
        cb, sb = c.expand_as(o), s.expand_as(o)
        jacobians.extend([Jg, torch.stack((torch.stack((cb, sb), -1), torch.stack((-sb, cb), -1)), -2)])
    return Point2(tensor=ret)


def _se2_transform_from(self, point: Union[Point2, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point2:
    """se2.py:434-468: R p + t;  d/dg = [R, R (-p_y, p_x)^T], d/dp = R."""
    p = _point_tensor(point, 2, "SE2.transform_from")
    c, s = self.tensor[:, 2], self.tensor[:, 3]
    ret = _rot2(p, c, s) + self.tensor[:, :2]
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        o = torch.ones_like(ret[:, 0])
        cb, sb = c * o, s * o
        px, py = p[:, 0] * o, p[:, 1] * o
        Jg = torch.stack((torch.stack((cb, -sb, -sb * px - cb * py), -1), torch.stack((sb, cb, cb * px - sb * py), -1)), -2)
        jacobians.extend([Jg, torch.stack((torch.stack((cb, -sb), -1), torch.stack((sb, cb), -1)), -2)])
    return Point2(tensor=ret)


def _se2_to_matrix(self) -> torch.Tensor:
    g = self.tensor
    c, s, x, y = g[:, 2], g[:, 3], g[:, 0], g[:, 1]
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack((torch.stack((c, -s, x), -1), torch.stack((s, c, y), -1), torch.stack((z, z, o), -1)), -2)


def _se2_hat(tangent_vector: torch.Tensor) -> torch.Tensor:
    """se2.py:375-383: [[0, -theta, u1], [theta, 0, u2], [0, 0, 0]]."""
    u1, u2, th_ = tangent_vector[:, 0], tangent_vector[:, 1], tangent_vector[:, 2]
    z = torch.zeros_like(th_)
    return torch.stack((torch.stack((z, -th_, u1), -1), torch.stack((th_, z, u2), -1), torch.stack((z, z, z), -1)), -2)


def _se2_vee(matrix: torch.Tensor) -> torch.Tensor:
    return torch.stack((matrix[:, 0, 2], matrix[:, 1, 2], 0.5 * (matrix[:, 1, 0] - matrix[:, 0, 1])), dim=1)


def _se2_update_from_x_y_theta(self, x_y_theta: torch.Tensor):
    self.update(torch.cat([x_y_theta[:, :2], x_y_theta[:, 2:3].cos(), x_y_theta[:, 2:3].sin()], dim=1))


def _so2_theta(self) -> torch.Tensor:
    return torch.atan2(self.tensor[:, 1], self.tensor[:, 0]).unsqueeze(1)


def _so2_to_cos_sin(self):
    return self.tensor[:, 0], self.tensor[:, 1]


def _so2_to_matrix(self) -> torch.Tensor:
    c, s = self.tensor[:, 0], self.tensor[:, 1]
    return torch.stack((torch.stack((c, -s), -1), torch.stack((s, c), -1)), -2)


def _so2_rotate(self, point: Union[Point2, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point2:
    """so2.py:278-291: R p;  d/dtheta = (-ret_y, ret_x)^T, d/dp = R."""
    p = _point_tensor(point, 2, "SO2.rotate")
    c, s = self.tensor[:, 0], self.tensor[:, 1]
    ret = _rot2(p, c, s)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        o = torch.ones_like(ret[:, 0])
        jacobians.extend([torch.stack((-ret[:, 1], ret[:, 0]), -1).unsqueeze(-1),
                          torch.stack((torch.stack((c * o, -s * o), -1), torch.stack((s * o, c * o), -1)), -2)])
    return Point2(tensor=ret)


def _so2_unrotate(self, point: Union[Point2, torch.Tensor], jacobians: Optional[List[torch.Tensor]] = None) -> Point2:
    """so2.py:293-306: R^T p;  d/dtheta = (ret_y, -ret_x)^T, d/dp = R^T."""
    p = _point_tensor(point, 2, "SO2.unrotate")
    c, s = self.tensor[:, 0], self.tensor[:, 1]
    ret = _rot2(p, c, -s)
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        o = torch.ones_like(ret[:, 0])
        jacobians.extend([torch.stack((ret[:, 1], -ret[:, 0]), -1).unsqueeze(-1),
                          torch.stack((torch.stack((c * o, s * o), -1), torch.stack((-s * o, c * o), -1)), -2)])
    return Point2(tensor=ret)


def _so2_hat(tangent_vector: torch.Tensor) -> torch.Tensor:
    th_ = tangent_vector.view(-1)
    z = torch.zeros_like(th_)
    return torch.stack((torch.stack((z, -th_), -1), torch.stack((th_, z), -1)), -2)


def _so2_vee(matrix: torch.Tensor) -> torch.Tensor:
    return (0.5 * (matrix[:, 1, 0] - matrix[:, 0, 1])).unsqueeze(1)


def _so2_update_from_angle(self, theta: torch.Tensor):
    theta = theta.view(-1, 1)
    self.update(torch.cat([theta.cos(), theta.sin()], dim=1))


# ---------------------------------------------------------------------------------------------------------------- Vector
def _vec_like(self, t: torch.Tensor):
    cls = type(self) if t.ndim == 2 and t.shape[1] == self.tensor.shape[1] else Vector
    return cls(tensor=t)


def _install_vector_api():
    """geometry/vector.py:60-230: the additive group on R^n and the element-wise helpers."""
    V = Vector
    V.__add__ = lambda self, o: _vec_like(self, self.tensor + _t(o))
    V.__sub__ = lambda self, o: _vec_like(self, self.tensor - _t(o))
    V.__neg__ = lambda self: _vec_like(self, -self.tensor)
    V.__mul__ = lambda self, o: _vec_like(self, self.tensor * _t(o))
    V.__rmul__ = V.__mul__
    V.__truediv__ = lambda self, o: _vec_like(self, self.tensor / _t(o))
    V.__matmul__ = lambda self, o: self.tensor @ _t(o)
    V.dot = lambda self, o: (self.tensor * _t(o)).sum(dim=1)
    V.inner = V.dot
    V.outer = lambda self, o: self.tensor.unsqueeze(2) * _t(o).unsqueeze(1)
    V.abs = lambda self: _vec_like(self, self.tensor.abs())
    V.norm = lambda self, *args, **kwargs: torch.norm(self.tensor, *args, **kwargs)
    V.cat = lambda self, vecs: Vector(tensor=torch.cat([self.tensor] + [_t(v) for v in (vecs if isinstance(vecs, (list, tuple)) else [vecs])], dim=1))
    V.allclose = lambda self, o, *args, **kwargs: torch.allclose(self.tensor, _t(o), *args, **kwargs)
    V.compose = lambda self, o: _vec_like(self, self.tensor + _t(o))
    V.inverse = lambda self: _vec_like(self, -self.tensor)
    V.between = lambda self, o: _vec_like(self, _t(o) - self.tensor)
    V.log_map = lambda self, jacobians=None: _vector_log(self, jacobians)
    V.exp_map = staticmethod(lambda tangent_vector, jacobians=None: _vector_exp(tangent_vector, jacobians))
    V.adjoint = lambda self: torch.eye(self.dof(), dtype=self.dtype, device=self.device).repeat(self.tensor.shape[0], 1, 1)
    V.to_matrix = lambda self: self.tensor.clone()
    V.project = lambda self, euclidean_grad, is_sparse=False: euclidean_grad.clone()
    Point2.x = lambda self: self.tensor[:, 0]
    Point2.y = lambda self: self.tensor[:, 1]
    Point3.x = lambda self: self.tensor[:, 0]
    Point3.y = lambda self: self.tensor[:, 1]
    Point3.z = lambda self: self.tensor[:, 2]


def _vector_log(self, jacobians):
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        jacobians.append(torch.eye(self.dof(), dtype=self.dtype, device=self.device).repeat(self.tensor.shape[0], 1, 1))
    return self.tensor.clone()


def _vector_exp(tangent_vector, jacobians):
    if jacobians is not None:
        LieGroup._check_jacobians_list(jacobians)
        jacobians.append(torch.eye(tangent_vector.shape[1], dtype=tangent_vector.dtype, device=tangent_vector.device).repeat(tangent_vector.shape[0], 1, 1))
    return Vector(tensor=tangent_vector.clone())


def _check_jacobians_list(jacobians: List[torch.Tensor]):
    """lie_group.py:43-46."""
    if len(jacobians) != 0:
        raise ValueError("jacobians list to be populated must be empty.")


def _project(self, euclidean_grad: torch.Tensor, is_sparse: bool = False) -> torch.Tensor:
    """Manifold.project (manifold.py:110-113): Euclidean gradient w.r.t. the storage -> tangent space; is_sparse=True is the Jacobian
    form [B, dim, *storage] (what AutoDiffCostFunction feeds it), the default a gradient of the storage's own shape."""
    if is_sparse:
        return type(self).project_tensor(self.tensor, euclidean_grad)
    return type(self).project_tensor(self.tensor, euclidean_grad.unsqueeze(1)).squeeze(1)


def _install_group_jacobians(cls):
    """lie_group.py:125-195: the optional Jacobian outputs of compose / inverse / between / local, assembled from adjoints and dlog exactly
    as the reference does (d compose = [Ad(g2^-1), I], d inverse = -Ad(g), d between = [Ad(g2^-1) (-Ad(g1)), I],
    d local = [-Ad(diff^-1) dlog, dlog])."""
    base_compose, base_inverse = cls.compose, cls.inverse

    def _eye(g):
        return torch.eye(g.dof(), dtype=g.dtype, device=g.device).repeat(g.tensor.shape[0], 1, 1)

    def compose(self, variable2, jacobians: Optional[List[torch.Tensor]] = None):
        if type(self) is not type(variable2):
            raise ValueError("Lie groups for compose must be of the same type.")
        out = base_compose(self, variable2)
        if jacobians is not None:
            LieGroup._check_jacobians_list(jacobians)
            jacobians.extend([base_inverse(variable2).adjoint(), _eye(variable2)])
        return out

    def inverse(self, jacobian: Optional[List[torch.Tensor]] = None):
        out = base_inverse(self)
        if jacobian is not None:
            LieGroup._check_jacobians_list(jacobian)
            jacobian.append(-self.adjoint())
        return out

    def between(self, variable2, jacobians: Optional[List[torch.Tensor]] = None):
        v1_inverse = base_inverse(self)
        out = base_compose(v1_inverse, variable2)
        if jacobians is not None:
            LieGroup._check_jacobians_list(jacobians)
            jacobians.extend([base_inverse(variable2).adjoint() @ (-self.adjoint()), _eye(variable2)])
        return out

    def local(self, variable2, jacobians: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        diff = between(self, variable2)
        if jacobians is None:
            return diff.log_map()
        LieGroup._check_jacobians_list(jacobians)
        dlog: List[torch.Tensor] = []
        ret = diff.log_map(dlog)
        jacobians.extend([-(base_inverse(diff).adjoint() @ dlog[0]), dlog[0]])
        return ret

    def retract(self, delta: torch.Tensor):
        return base_compose(self, cls.exp_map(delta))

    cls.compose, cls.inverse, cls.between, cls.local, cls.retract = compose, inverse, between, local, retract


def typed_view(var: Variable, tensor: torch.Tensor) -> Variable:
    """A variable of `var`'s class and name around another tensor (what the reference's AutoDiffCostFunction hands to err_fn:
    copies of the variables holding the traced tensors, cost_function.py:283-316) -- built without the constructor: no checks, no
    bump of the update counters the engine watches."""
    o = object.__new__(type(var))
    o._tensor, o.name, o._id, o._num_updates = tensor, var.name, var._id, 0
    return o


def install():
    from . import functional
    LieGroup._check_jacobians_list = staticmethod(_check_jacobians_list)
    for cls in (SE3, SO3, SE2, SO2):
        _install_tape_route(cls)
        _install_group_jacobians(cls)
        cls.project = _project
    SE3.rotation, SE3.translation = _se3_rotation, _se3_translation
    SE3.transform_from, SE3.transform_to, SE3.to_matrix = _se3_transform_from, _se3_transform_to, _se3_to_matrix
    SE3.hat, SE3.vee = staticmethod(_se3_hat), staticmethod(_se3_vee)
    SE3.x_y_z_unit_quaternion_to_SE3 = staticmethod(_x_y_z_unit_quaternion_to_se3)
    SE3.to_x_y_z_quaternion = _se3_to_x_y_z_quaternion
    SE3.update_from_x_y_z_quaternion = lambda self, q: self.update(_x_y_z_unit_quaternion_to_se3(q).tensor)
    SE3.update_from_rot_and_trans = lambda self, rotation, translation: self.update(torch.cat((rotation.tensor, translation.tensor.unsqueeze(-1)), dim=-1))
    SO3.rotate, SO3.unrotate, SO3.to_matrix = _so3_rotate, _so3_unrotate, (lambda self: self.tensor.clone())
    SO3.hat, SO3.vee = staticmethod(_so3_hat), staticmethod(_so3_vee)
    SO3.to_quaternion, SO3.unit_quaternion_to_SO3 = _so3_to_quaternion, staticmethod(_unit_quaternion_to_so3)
    SO3.update_from_unit_quaternion = lambda self, q: self.update(_unit_quaternion_to_so3(q).tensor)
    SE2.rotation, SE2.translation = property(_se2_rotation), property(_se2_translation)      # properties in se2.py:125-141
    SE2.theta, SE2.xy = _se2_theta, _se2_xy
    SE2.transform_to, SE2.transform_from, SE2.to_matrix = _se2_transform_to, _se2_transform_from, _se2_to_matrix
    SE2.hat, SE2.vee = staticmethod(_se2_hat), staticmethod(_se2_vee)
    SE2.update_from_x_y_theta = _se2_update_from_x_y_theta
    SE2.update_from_rot_and_trans = lambda self, rotation, translation: self.update(torch.cat((translation.tensor, rotation.tensor), dim=1))
    SO2.theta, SO2.to_cos_sin, SO2.to_matrix = _so2_theta, _so2_to_cos_sin, _so2_to_matrix
    SO2.rotate, SO2.unrotate = _so2_rotate, _so2_unrotate
    SO2.hat, SO2.vee = staticmethod(_so2_hat), staticmethod(_so2_vee)
    SO2.update_from_angle = _so2_update_from_angle
    _install_vector_api()
    for cls, key in ((SE3, "se3"), (SO3, "so3"), (SE2, "se2"), (SO2, "so2"), (Vector, "vector"), (Point2, "point2"), (Point3, "point3")):
        cls.rand = staticmethod(getattr(functional, f"rand_{key}"))
        cls.randn = staticmethod(getattr(functional, f"randn_{key}"))
