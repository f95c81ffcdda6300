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


def so2_exp(theta: torch.Tensor) -> torch.Tensor:
    """theseus/geometry/so2.py:99-100, :167-186: tangent [B,1] -> storage [cos, sin]."""
    return torch.cat((theta.cos(), theta.sin()), dim=-1)


def so2_log(X: torch.Tensor) -> torch.Tensor:
    """so2.py:206-222: atan2(sin, cos) -> [B,1]."""
    return torch.atan2(X[..., 1], X[..., 0]).unsqueeze(-1)


def so2_inverse(X: torch.Tensor) -> torch.Tensor:
    return torch.stack((X[..., 0], -X[..., 1]), dim=-1)   # so2.py:232-234


def so2_compose(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    return torch.stack((A[..., 0] * B[..., 0] - A[..., 1] * B[..., 1], A[..., 1] * B[..., 0] + A[..., 0] * B[..., 1]), dim=-1)  # so2.py:224-230


def retract(kind: int, X: torch.Tensor, delta: torch.Tensor) -> torch.Tensor:
    """kind = Manifold.KIND (thb_var_kind): 0 SE3, 1 SO3, 2 Vector, 3 SE2, 4 SO2."""
    if kind == 4:
        return so2_compose(X, so2_exp(delta))
    if kind == 2:
        return X + delta.view(delta.shape[0], *X.shape[1:])
    if kind == 0:
        return se3_retract(X, delta)
    if kind == 1:
        return so3_retract(X, delta)
    if kind == 3:
        return se2_retract(X, delta)
    raise NotImplementedError(f"differentiable retraction for variable kind {kind}")


# ---------------------------------------------------------------------------------------------------------------------------------
# inverse / compose / log: torch restatements used to put the fused-kernel cost functions (Between, Difference, Reprojection) on the
# autograd tape of the backward modes.  torchlie so3_impl.py:390-433 (_log_impl_helper), se3_impl.py:354-396, :578-581, :703-708;
# theseus/geometry/se2.py:165-228, :318-339.
_NEAR_PI = {torch.float32: 1e-2, torch.float64: 1e-7}


def so3_log(R: torch.Tensor):
    sa = 0.5 * torch.stack((R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]), dim=-1)
    cosine = 0.5 * (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1)
    sine = sa.norm(dim=-1)
    theta = torch.atan2(sine, cosine)
    nz = theta < _NEAR_ZERO[R.dtype]
    npi = 1 + cosine <= _NEAR_PI[R.dtype]
    nzp = nz | npi
    one = torch.ones((), dtype=R.dtype, device=R.device)
    sine_nz = torch.where(nzp, one, sine)
    scale = torch.where(nzp, 1 + sine ** 2 / 6, theta / sine_nz)
    ret = sa * scale[..., None]
    # near pi: axis from the dominant column of (R + R^T)/2 - cos I
    d = torch.diagonal(R, dim1=-2, dim2=-1)
    major = ((d[..., 1] > d[..., 0]) & (d[..., 1] > d[..., 2])).long() + 2 * ((d[..., 2] > d[..., 0]) & (d[..., 2] > d[..., 1])).long()
    onehot = (major[..., None] == torch.arange(3, device=R.device)).to(R.dtype)  # (one_hot has no vmap batching rule)
    row = (R * onehot[..., :, None]).sum(dim=-2)
    col = (R * onehot[..., None, :]).sum(dim=-1)
    sel = 0.5 * (row + col) - onehot * cosine[..., None]
    axis = sel / torch.where(nz, one, sel.norm(dim=-1))[..., None]
    sgn_t = torch.sign((sa * onehot).sum(dim=-1))
    sgn = torch.where(sgn_t != 0, sgn_t, one)
    out = torch.where(npi[..., None], axis * (theta * sgn)[..., None], ret)
    return out, (theta, sine, cosine)


def se3_log(T: torch.Tensor) -> torch.Tensor:
    w, (theta, sine, cosine) = so3_log(T[..., :3])
    nz = theta < _NEAR_ZERO[T.dtype]
    one = torch.ones((), dtype=T.dtype, device=T.device)
    theta2 = theta ** 2
    st = sine * theta
    tcm2 = 2 * cosine - 2
    tcm2_nz = torch.where(nz, one, tcm2)
    theta2_nz = torch.where(nz, one, theta2)
    a = torch.where(nz, 1 - theta2 / 12, -st / tcm2_nz)
    b = torch.where(nz, 1.0 / 12 + theta2 / 720, (st + tcm2) / (theta2_nz * tcm2_nz))
    t = T[..., 3]
    lin = a[..., None] * t - 0.5 * torch.linalg.cross(w, t) + b[..., None] * (w * (w * t).sum(-1, keepdim=True))
    return torch.cat((lin, w), dim=-1)


def se3_inverse(T: torch.Tensor) -> torch.Tensor:
    Rt = T[..., :3].transpose(-1, -2)
    return torch.cat((Rt, -(Rt @ T[..., 3:])), dim=-1)


def se3_compose(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    return torch.cat((A[..., :3] @ B[..., :3], A[..., :3] @ B[..., 3:] + A[..., 3:]), dim=-1)


def se2_log(T: torch.Tensor) -> torch.Tensor:
    cosine, sine = T[..., 2], T[..., 3]
    theta = torch.atan2(sine, cosine)
    small = theta.abs() < _SE2_NEAR_ZERO[T.dtype]
    sine_nz = torch.where(small, torch.ones((), dtype=T.dtype, device=T.device), sine)
    a = 0.5 * (1 + cosine) * torch.where(small, 1 + sine ** 2 / 6, theta / sine_nz)
    b = 0.5 * theta
    return torch.stack((a * T[..., 0] + b * T[..., 1], a * T[..., 1] - b * T[..., 0], theta), dim=-1)


def se2_inverse(T: torch.Tensor) -> torch.Tensor:
    c, s = T[..., 2], T[..., 3]
    return torch.stack((-(c * T[..., 0] + s * T[..., 1]), -(-s * T[..., 0] + c * T[..., 1]), c, -s), dim=-1)


def se2_compose(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    c0, s0, c1, s1 = A[..., 2], A[..., 3], B[..., 2], B[..., 3]
    return torch.stack((c0 * B[..., 0] - s0 * B[..., 1] + A[..., 0], s0 * B[..., 0] + c0 * B[..., 1] + A[..., 1],
                        c0 * c1 - s0 * s1, s0 * c1 + c0 * s1), dim=-1)


def local(kind: int, X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """log(X^-1 Y): `X.local(Y)` of the reference (lie_group.py:163-170); Vector: Y - X."""
    if kind == 2:
        return Y - X
    if kind == 0:
        return se3_log(se3_compose(se3_inverse(X), Y))
    if kind == 1:
        return so3_log(X.transpose(-1, -2) @ Y)[0]
    if kind == 3:
        return se2_log(se2_compose(se2_inverse(X), Y))
    if kind == 4:
        return so2_log(so2_compose(so2_inverse(X), Y))
    raise NotImplementedError(f"local() for variable kind {kind}")


def between(kind: int, X: torch.Tensor, Y: torch.Tensor) -> torch.Tensor:
    """X^-1 Y (lie_group.py:150-161)."""
    if kind == 0:
        return se3_compose(se3_inverse(X), Y)
    if kind == 1:
        return X.transpose(-1, -2) @ Y
    if kind == 3:
        return se2_compose(se2_inverse(X), Y)
    if kind == 4:
        return so2_compose(so2_inverse(X), Y)
    if kind == 2:
        return Y - X            # Vector: the additive group (geometry/vector.py: inverse = -x, compose = x + y)
    raise NotImplementedError(f"between() for variable kind {kind}")


def velocity_to_tangent(kind: int, D: torch.Tensor, dD: torch.Tensor) -> torch.Tensor:
    """Tangent coordinates xi of velocities dD = D hat(xi) of a group element D: xi = vee(D^-1 dD).
    D [B, *g], dD [B, K, *g] (K velocities per element) -> [B, K, dof]."""
    if kind == 0:   # SE3: D^-1 dD = [R^T dR | R^T dt]
        M = D[:, None, :, :3].transpose(-1, -2) @ dD
        w = 0.5 * torch.stack((M[..., 2, 1] - M[..., 1, 2], M[..., 0, 2] - M[..., 2, 0], M[..., 1, 0] - M[..., 0, 1]), dim=-1)
        return torch.cat((M[..., 3], w), dim=-1)
    if kind == 1:
        M = D[:, None].transpose(-1, -2) @ dD
        return 0.5 * torch.stack((M[..., 2, 1] - M[..., 1, 2], M[..., 0, 2] - M[..., 2, 0], M[..., 1, 0] - M[..., 0, 1]), dim=-1)
    if kind == 3:   # SE2 storage [x, y, cos, sin]
        c, s_ = D[:, None, 2], D[:, None, 3]
        vx = c * dD[..., 0] + s_ * dD[..., 1]
        vy = -s_ * dD[..., 0] + c * dD[..., 1]
        return torch.stack((vx, vy, c * dD[..., 3] - s_ * dD[..., 2]), dim=-1)
    if kind == 4:   # SO2 storage [cos, sin]: theta' = cos dsin - sin dcos
        return (D[:, None, 0] * dD[..., 1] - D[:, None, 1] * dD[..., 0]).unsqueeze(-1)
    raise NotImplementedError(f"velocity_to_tangent for variable kind {kind}")

