"""Module-level conveniences of the reference's package namespace (theseus/geometry/__init__.py: adjoint, between, compose, exp_map,
inverse, local, log_map, retract; rand_* / randn_* generators).  The operators forward to the group classes (CUDA kernels); the random
generators build inputs with the torch restatements, so they also work on the CPU (input fabrication is not the compute path)."""
import math
from typing import Optional

import torch

from . import lie_torch
from .geometry import SE2, SE3, SO2, SO3, LieGroup, Point2, Point3, Vector


def adjoint(group: LieGroup) -> torch.Tensor:
    return group.adjoint()


def between(a: LieGroup, b: LieGroup) -> LieGroup:
    return a.between(b)


def compose(a: LieGroup, b: LieGroup) -> LieGroup:
    return a.compose(b)


def inverse(group: LieGroup) -> LieGroup:
    return group.inverse()


def local(a, b) -> torch.Tensor:
    return a.local(b)


def log_map(group: LieGroup) -> torch.Tensor:
    return group.log_map()


def exp_map(group_like: LieGroup, tangent_vector: torch.Tensor) -> LieGroup:
    return type(group_like).exp_map(tangent_vector)


def retract(group, delta: torch.Tensor):
    return group.retract(delta)


def _gen(generator, *size, dtype, device, normal):
    f = torch.randn if normal else torch.rand
    return f(*size, generator=generator, dtype=dtype or torch.get_default_dtype(), device=device)


def _rand_so3_tangent(B, generator, dtype, device, normal):
    if normal:
        return _gen(generator, B, 3, dtype=dtype, device=device, normal=True)
    # uniform direction, angle in [0, pi): a simple (not Haar-exact) input generator
    v = _gen(generator, B, 3, dtype=dtype, device=device, normal=True)
    v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return v * (math.pi * _gen(generator, B, 1, dtype=dtype, device=device, normal=False))


def _make(cls_name):
    def rand(*size, generator=None, dtype: Optional[torch.dtype] = None, device=None, requires_grad: bool = False, _normal=False):
        B = int(size[0]) if size else 1
        if cls_name in ("vector", "point2", "point3"):
            d = {"point2": 2, "point3": 3}.get(cls_name, int(size[1]) if len(size) > 1 else 1)
            t = _gen(generator, B, d, dtype=dtype, device=device, normal=_normal)
            out = {"vector": Vector, "point2": Point2, "point3": Point3}[cls_name](tensor=t)
        elif cls_name == "so3":
            out = SO3(tensor=lie_torch._so3_exp_parts(_rand_so3_tangent(B, generator, dtype, device, _normal))[0])
        elif cls_name == "se3":
            R = lie_torch._so3_exp_parts(_rand_so3_tangent(B, generator, dtype, device, _normal))[0]
            out = SE3(tensor=torch.cat((R, _gen(generator, B, 3, 1, dtype=dtype, device=device, normal=_normal)), dim=2))
        elif cls_name == "so2":   # so2.py:50-94: uniform angle in [-pi, pi) / normal angle
            ang = _gen(generator, B, 1, dtype=dtype, device=device, normal=_normal)
            out = SO2(theta=ang if _normal else 2 * math.pi * ang - math.pi)
        else:  # se2
            th_ = _gen(generator, B, 1, dtype=dtype, device=device, normal=_normal) * (1.0 if _normal else 2 * math.pi) - (0.0 if _normal else math.pi)
            out = SE2(x_y_theta=torch.cat((_gen(generator, B, 2, dtype=dtype, device=device, normal=_normal), th_), dim=1))
        if requires_grad:
            out.tensor.requires_grad_(True)
        return out

    def randn(*size, **kw):
        return rand(*size, _normal=True, **kw)
    return rand, randn


rand_vector, randn_vector = _make("vector")
rand_point2, randn_point2 = _make("point2")
rand_point3, randn_point3 = _make("point3")
rand_so3, randn_so3 = _make("so3")
rand_se3, randn_se3 = _make("se3")
rand_se2, randn_se2 = _make("se2")
rand_so2, randn_so2 = _make("so2")
