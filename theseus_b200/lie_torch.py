"""Differentiable torch restatements of the retraction X <- X * exp(delta), used ONLY on the autograd tape of the backward modes
(optimizer._optimize_loop_differentiable); the forward-only path retracts with the fused kernel (thb_retract).

Closed forms and near-zero branches follow the reference: torchlie so3_impl.py:220-261 (_exp_impl_helper), se3_impl.py:178-216
(_exp_impl_helper), :703-708 (_compose_impl), theseus/geometry/se2.py:239-300 (exp_map), :318-332 (_compose_impl),
theseus/geometry/lie_group.py:197-198 (_retract_impl).  eps tables: torchlie/global_params.py:44-58, theseus/global_params.py:46-59.
"""
import torch

_NEAR_ZERO = {torch.float32: 1e-2, torch.float64: 5e-3}
_SE2_NEAR_ZERO = {torch.float32: 3e-2, torch.float64: 1e-6}


def _so3_exp_parts(w: torch.Tensor):
    theta = w.norm(dim=-1, keepdim=True)
    theta2 = theta * theta
    nz = theta < _NEAR_ZERO[w.dtype]
    one = torch.ones((), dtype=w.dtype, device=w.device)
    theta_nz = torch.where(nz, one, theta)
    theta2_nz = torch.where(nz, one, theta2)
    cosine = torch.where(nz, 8 / (4 + theta2) - 1, theta.cos())
    sine = theta.sin()
    sine_by_theta = torch.where(nz, 0.5 * cosine + 0.5, sine / theta_nz)
    omc = torch.where(nz, 0.5 * sine_by_theta, (1 - cosine) / theta2_nz)
    sa = sine_by_theta * w
    z = torch.zeros_like(sa[..., 0])
    skew = torch.stack((torch.stack((z, -sa[..., 2], sa[..., 1]), -1), torch.stack((sa[..., 2], z, -sa[..., 0]), -1),
                        torch.stack((-sa[..., 1], sa[..., 0], z), -1)), -2)
    R = omc[..., None] * (w[..., :, None] * w[..., None, :]) + cosine[..., None] * torch.eye(3, dtype=w.dtype, device=w.device) + skew
    return R, (theta, theta2, theta_nz, theta2_nz, sine, sine_by_theta, omc, nz)


def so3_retract(R: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    return R @ _so3_exp_parts(delta)[0]


def se3_exp(xi: torch.Tensor) -> torch.Tensor:
    v, w = xi[..., :3], xi[..., 3:]
    R, (theta, theta2, theta_nz, theta2_nz, sine, sbt, omc, nz) = _so3_exp_parts(w)
    tms = torch.where(nz, 1.0 / 6 - theta2 / 120, (theta - sine) / (theta_nz * theta2_nz))
    t = sbt * v + omc * torch.linalg.cross(w, v) + tms * (w * (w * v).sum(-1, keepdim=True))
    return torch.cat((R, t[..., None]), dim=-1)


def se3_retract(T: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    E = se3_exp(delta)
    return torch.cat((T[..., :3] @ E[..., :3], T[..., :3] @ E[..., 3:] + T[..., 3:]), dim=-1)


def se2_exp(xi: torch.Tensor) -> torch.Tensor:
    u, theta = xi[..., :2], xi[..., 2]
    cosine, sine = theta.cos(), theta.sin()
    small = theta.abs() < _SE2_NEAR_ZERO[xi.dtype]
    theta_nz = torch.where(small, torch.ones((), dtype=xi.dtype, device=xi.device), theta)
    sbt = torch.where(small, 1 - theta ** 2 / 6, sine / theta_nz)
    cmo = torch.where(small, -theta / 2 + theta ** 3 / 24, (cosine - 1) / theta_nz)
    return torch.stack((sbt * u[..., 0] + cmo * u[..., 1], sbt * u[..., 1] - cmo * u[..., 0], cosine, sine), dim=-1)


def se2_retract(T: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    E = se2_exp(delta)
    c0, s0, c1, s1 = T[..., 2], T[..., 3], E[..., 2], E[..., 3]
    return torch.stack((c0 * E[..., 0] - s0 * E[..., 1] + T[..., 0], s0 * E[..., 0] + c0 * E[..., 1] + T[..., 1],
                        c0 * c1 - s0 * s1, s0 * c1 + c0 * s1), dim=-1)


def retract(kind: int, X: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    """kind = Manifold.KIND (thb_var_kind): 0 SE3, 1 SO3, 2 Vector, 3 SE2."""
    if kind == 2:
        return X + delta.view(delta.shape[0], *X.shape[1:])
    if kind == 0:
        return se3_retract(X, delta)
    if kind == 1:
        return so3_retract(X, delta)
    if kind == 3:
        return se2_retract(X, delta)
    raise NotImplementedError(f"differentiable retraction for variable kind {kind}")
