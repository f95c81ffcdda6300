"""Oracle: SO3 / SE3 / SE2 closed forms (numpy restatement; test infrastructure only).

Follows torchlie/torchlie/functional/so3_impl.py and se3_impl.py of the
reference (line numbers cited per function).  All functions take arrays with a
leading batch shape ``[..., k]`` and compute in the dtype of their input
(float32 or float64), using the reference's per-dtype eps table
(torchlie/torchlie/global_params.py:44-58).

Storage conventions (reference): SO3 ``[...,3,3]``; SE3 ``[...,3,4] = [R|t]``;
SE3 tangent = ``[v(3), w(3)]`` (translation first, se3_impl.py:195-196).
"""
import numpy as np

# torchlie/torchlie/global_params.py:44-58
_EPS = {
    np.dtype("float32"): dict(near_pi=1e-2, near_zero=1e-2, d_near_zero=2e-1),
    np.dtype("float64"): dict(near_pi=1e-7, near_zero=5e-3, d_near_zero=1e-2),
}
# theseus/global_params.py:46-59 (SE2 / SO2)
_EPS_TH = {
    np.dtype("float32"): dict(so2_norm=1e-12, so2_matrix=1e-5, se2_near_zero=3e-2, se2_d_near_zero=1e-1),
    np.dtype("float64"): dict(so2_norm=1e-12, so2_matrix=4e-7, se2_near_zero=1e-6, se2_d_near_zero=1e-3),
}
_NON_ZERO = 1.0  # torchlie/functional/constants.py:17


def eps(kind, dtype):
    return _EPS[np.dtype(dtype)][kind]


def hat3(w):
    """so3_impl.py:587-599 (_hat_impl)."""
    out = np.zeros(w.shape[:-1] + (3, 3), dtype=w.dtype)
    out[..., 0, 1] = -w[..., 2]
    out[..., 0, 2] = w[..., 1]
    out[..., 1, 0] = w[..., 2]
    out[..., 1, 2] = -w[..., 0]
    out[..., 2, 0] = -w[..., 1]
    out[..., 2, 1] = w[..., 0]
    return out


def _outer(a, b):
    return a[..., :, None] * b[..., None, :]


# ----------------------------------------------------------------------------- SO3
def so3_exp_helper(w):
    """so3_impl.py:220-261 (_exp_impl_helper)."""
    dt = w.dtype
    theta = np.linalg.norm(w, axis=-1)[..., None, None].astype(dt)
    theta2 = theta**2
    nz = theta < eps("near_zero", dt)
    theta_nz = np.where(nz, dt.type(_NON_ZERO), theta)
    theta2_nz = np.where(nz, dt.type(_NON_ZERO), theta2)
    cosine = np.where(nz, 8 / (4 + theta2) - 1, np.cos(theta)).astype(dt)
    sine = np.sin(theta).astype(dt)
    sine_by_theta = np.where(nz, 0.5 * cosine + 0.5, sine / theta_nz).astype(dt)
    omc = np.where(nz, 0.5 * sine_by_theta, (1 - cosine) / theta2_nz).astype(dt)
    R = omc * _outer(w, w)
    c = cosine[..., 0, 0]
    R[..., 0, 0] += c
    R[..., 1, 1] += c
    R[..., 2, 2] += c
    sa = sine_by_theta[..., 0] * w
    R[..., 0, 1] -= sa[..., 2]
    R[..., 1, 0] += sa[..., 2]
    R[..., 0, 2] += sa[..., 1]
    R[..., 2, 0] -= sa[..., 1]
    R[..., 1, 2] -= sa[..., 0]
    R[..., 2, 1] += sa[..., 0]
    return R.astype(dt), (theta, theta2, theta_nz, theta2_nz, sine, cosine, sine_by_theta, omc)


def so3_exp(w):
    return so3_exp_helper(w)[0]


def so3_jexp(w):
    """so3_impl.py:270-320 (_jexp_impl): right Jacobian of exp."""
    dt = w.dtype
    _, (theta, theta2, theta_nz, theta2_nz, sine, _, sbt, omc) = so3_exp_helper(w)
    nz = theta < eps("near_zero", dt)
    theta3_nz = theta_nz * theta2_nz
    tms = np.where(nz, dt.type(0), (theta - sine) / theta3_nz).astype(dt)
    J = tms * _outer(w, w)
    for i in range(3):
        J[..., i, i] += sbt[..., 0, 0]
    t = omc[..., 0] * w
    J[..., 0, 1] += t[..., 2]
    J[..., 1, 0] -= t[..., 2]
    J[..., 0, 2] -= t[..., 1]
    J[..., 2, 0] += t[..., 1]
    J[..., 1, 2] += t[..., 0]
    J[..., 2, 1] -= t[..., 0]
    return J.astype(dt)


def so3_log_helper(R):
    """so3_impl.py:390-433 (_log_impl_helper)."""
    dt = R.dtype
    sa = np.zeros(R.shape[:-2] + (3,), dtype=dt)
    sa[..., 0] = 0.5 * (R[..., 2, 1] - R[..., 1, 2])
    sa[..., 1] = 0.5 * (R[..., 0, 2] - R[..., 2, 0])
    sa[..., 2] = 0.5 * (R[..., 1, 0] - R[..., 0, 1])
    cosine = (0.5 * (R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2] - 1)).astype(dt)
    sine = np.linalg.norm(sa, axis=-1).astype(dt)
    theta = np.arctan2(sine, cosine).astype(dt)
    nz = theta < eps("near_zero", dt)
    npi = 1 + cosine <= eps("near_pi", dt)
    nzp = nz | npi
    sine_nz = np.where(nzp, dt.type(_NON_ZERO), sine)
    scale = np.where(nzp, 1 + sine**2 / 6, theta / sine_nz).astype(dt)
    ret = sa * scale[..., None]
    # near pi
    d0, d1, d2 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    major = ((d1 > d0) & (d1 > d2)).astype(np.int64) + 2 * ((d2 > d0) & (d2 > d1)).astype(np.int64)
    m = major[..., None, None]
    row = np.take_along_axis(R, np.broadcast_to(m, R.shape[:-2] + (1, 3)), axis=-2)[..., 0, :]
    col = np.take_along_axis(R, np.broadcast_to(m, R.shape[:-2] + (3, 1)), axis=-1)[..., :, 0]
    sel = (0.5 * (row + col)).astype(dt)
    onehot = (np.arange(3) == major[..., None])
    sel = sel - onehot * cosine[..., None]
    nrm = np.linalg.norm(sel, axis=-1).astype(dt)
    axis = sel / np.where(nz, dt.type(_NON_ZERO), nrm)[..., None]
    sgn_t = np.sign(np.sum(sa * onehot, axis=-1))
    sgn = np.where(sgn_t != 0, sgn_t, 1.0).astype(dt)
    out = np.where(npi[..., None], axis * (theta * sgn)[..., None], ret)
    return out.astype(dt), (theta, sine, cosine)


def so3_log(R):
    return so3_log_helper(R)[0]


def so3_jlog_helper(w, theta, sine, cosine):
    """so3_impl.py:442-479 (_jlog_impl_helper)."""
    dt = w.dtype
    dnz = theta < eps("d_near_zero", dt)
    theta2 = theta**2
    st = sine * theta
    tcm2 = 2 * cosine - 2
    tcm2_nz = np.where(dnz, dt.type(_NON_ZERO), tcm2)
    theta2_nz = np.where(dnz, dt.type(_NON_ZERO), theta2)
    a = np.where(dnz, 1 - theta2 / 12, -st / tcm2_nz).astype(dt)
    b = np.where(dnz, 1.0 / 12 + theta2 / 720, (st + tcm2) / (theta2_nz * tcm2_nz)).astype(dt)
    bw = b[..., None] * w
    J = _outer(bw, w)
    h = 0.5 * w
    J[..., 0, 1] -= h[..., 2]
    J[..., 1, 0] += h[..., 2]
    J[..., 0, 2] += h[..., 1]
    J[..., 2, 0] -= h[..., 1]
    J[..., 1, 2] -= h[..., 0]
    J[..., 2, 1] += h[..., 0]
    for i in range(3):
        J[..., i, i] += a
    return J.astype(dt), bw


def so3_jlog(R):
    w, (theta, sine, cosine) = so3_log_helper(R)
    return so3_jlog_helper(w, theta, sine, cosine)[0], w


def so3_inverse(R):
    """so3_impl.py:561-563."""
    return np.swapaxes(R, -1, -2).copy()


def so3_compose(R0, R1):
    """so3_impl.py:669-672."""
    return R0 @ R1


def so3_adjoint(R):
    return R.copy()


# ----------------------------------------------------------------------------- SE3
def se3_exp(xi):
    """se3_impl.py:178-216 (_exp_impl_helper)."""
    dt = xi.dtype
    v, w = xi[..., :3], xi[..., 3:]
    R, (theta, theta2, theta_nz, theta2_nz, sine, _, sbt, omc) = so3_exp_helper(w)
    nz = theta < eps("near_zero", dt)
    theta3_nz = theta_nz * theta2_nz
    tms = np.where(nz, 1.0 / 6 - theta2 / 120, (theta - sine) / theta3_nz).astype(dt)
    t = sbt[..., 0] * v
    t = t + omc[..., 0] * np.cross(w, v)
    t = t + tms[..., 0] * (w * np.sum(w * v, axis=-1, keepdims=True))
    out = np.concatenate([R, t[..., None]], axis=-1)
    return out.astype(dt)


def se3_jexp(xi):
    """se3_impl.py:225-330 (_jexp_impl_helper / _jexp_impl): right Jacobian of exp."""
    dt = xi.dtype
    v, w = xi[..., :3], xi[..., 3:]
    R, (theta, theta2, theta_nz, theta2_nz, sine, _, sbt, omc) = so3_exp_helper(w)
    nz = theta < eps("near_zero", dt)
    theta3_nz = theta_nz * theta2_nz
    tms_t = np.where(nz, 1.0 / 6 - theta2 / 120, (theta - sine) / theta3_nz).astype(dt)
    tms_rot = np.where(nz, dt.type(0), tms_t)
    J = np.zeros(xi.shape[:-1] + (6, 6), dtype=dt)
    Jr = tms_rot * _outer(w, w)
    for i in range(3):
        Jr[..., i, i] += sbt[..., 0, 0]
    t = omc[..., 0] * w
    Jr[..., 0, 1] += t[..., 2]
    Jr[..., 1, 0] -= t[..., 2]
    Jr[..., 0, 2] -= t[..., 1]
    Jr[..., 2, 0] += t[..., 1]
    Jr[..., 1, 2] += t[..., 0]
    Jr[..., 2, 1] -= t[..., 0]
    J[..., :3, :3] = Jr
    J[..., 3:, 3:] = Jr
    d_omc = np.where(nz, dt.type(-1 / 12.0), (sbt - 2 * omc) / theta2_nz).astype(dt)[..., 0]
    d_tms = np.where(nz, dt.type(-1 / 60.0), (omc - 3 * tms_t) / theta2_nz).astype(dt)[..., 0]
    wv = np.cross(w, v)
    wwv = np.cross(w, wv)
    sw = tms_t[..., 0] * w
    Jt = _outer(d_omc * wv + d_tms * wwv, w)
    Jt = Jt - _outer(v, sw)
    tv = -omc[..., 0] * v - tms_t[..., 0] * wv
    Jt = Jt + hat3(tv)
    sv = np.sum(sw * v, axis=-1)
    for i in range(3):
        Jt[..., i, i] += sv
    J[..., :3, 3:] = np.swapaxes(R, -1, -2) @ Jt
    return J.astype(dt)


def se3_log_helper(T):
    """se3_impl.py:354-396 (_log_impl_helper)."""
    dt = T.dtype
    w, (theta, sine, cosine) = so3_log_helper(T[..., :3])
    nz = theta < eps("near_zero", dt)
    theta2 = theta**2
    st = sine * theta
    tcm2 = 2 * cosine - 2
    tcm2_nz = np.where(nz, dt.type(_NON_ZERO), tcm2)
    theta2_nz = np.where(nz, dt.type(_NON_ZERO), theta2)
    a = np.where(nz, 1 - theta2 / 12, -st / tcm2_nz).astype(dt)
    b = np.where(nz, 1.0 / 12 + theta2 / 720, (st + tcm2) / (theta2_nz * tcm2_nz)).astype(dt)
    t = T[..., 3]
    lin = a[..., None] * t
    lin = lin - 0.5 * np.cross(w, t)
    lin = lin + b[..., None] * (w * np.sum(w * t, axis=-1, keepdims=True))
    xi = np.concatenate([lin, w], axis=-1).astype(dt)
    return xi, (theta, theta2, theta2_nz, sine, cosine, tcm2_nz, a, b)


def se3_log(T):
    return se3_log_helper(T)[0]


def se3_jlog(T):
    """se3_impl.py:405-483 (_jlog_impl_helper / _jlog_impl). Returns (J[...,6,6], xi)."""
    dt = T.dtype
    xi, (theta, theta2, theta2_nz, sine, cosine, tcm2_nz, a, b) = se3_log_helper(T)
    lin, ang = xi[..., :3], xi[..., 3:]
    dnz = theta < eps("d_near_zero", dt)
    J = np.zeros(xi.shape[:-1] + (6, 6), dtype=dt)
    Jr, b_ang = so3_jlog_helper(ang, theta, sine, cosine)
    J[..., :3, :3] = Jr
    J[..., 3:, 3:] = Jr
    theta_nz = np.where(dnz, dt.type(_NON_ZERO), theta)
    theta4_nz = theta2_nz**2
    c = np.where(dnz, -1 / 360.0 - theta2 / 7560.0,
                 -(2 * tcm2_nz + theta * sine + theta2) / (theta4_nz * tcm2_nz)).astype(dt)
    d = np.where(dnz, -1 / 6.0 - theta2 / 180.0, (theta - sine) / (theta_nz * tcm2_nz)).astype(dt)
    e = np.sum(ang * lin, axis=-1)
    ce_ang = (c * e)[..., None] * ang
    Q = _outer(ce_ang, ang)
    Q = Q + _outer(b_ang, lin) + _outer(lin, b_ang)
    for i in range(3):
        Q[..., i, i] += e * d
    J[..., :3, 3:] = Q
    h = 0.5 * lin
    J[..., 0, 4] -= h[..., 2]
    J[..., 1, 3] += h[..., 2]
    J[..., 0, 5] += h[..., 1]
    J[..., 2, 3] -= h[..., 1]
    J[..., 1, 5] -= h[..., 0]
    J[..., 2, 4] += h[..., 0]
    return J.astype(dt), xi


def se3_adjoint(T):
    """se3_impl.py:531-538."""
    out = np.zeros(T.shape[:-2] + (6, 6), dtype=T.dtype)
    R = T[..., :3]
    out[..., :3, :3] = R
    out[..., 3:, 3:] = R
    out[..., :3, 3:] = hat3(T[..., 3]) @ R
    return out


def se3_inverse(T):
    """se3_impl.py:578-581."""
    Rt = np.swapaxes(T[..., :3], -1, -2)
    t = -(Rt @ T[..., 3:])
    return np.concatenate([Rt, t], axis=-1)


def se3_compose(T0, T1):
    """se3_impl.py:703-708."""
    R = T0[..., :3] @ T1[..., :3]
    t = T0[..., :3] @ T1[..., 3:] + T0[..., 3:]
    return np.concatenate([R, t], axis=-1)


def se3_transform(T, p):
    """se3_impl.py:757-761 (transform: R p + t)."""
    return (T[..., :3] @ p[..., None])[..., 0] + T[..., 3]


def se3_retract(T, delta):
    """theseus/geometry/lie_group.py:197-198 (_retract_impl = compose(exp_map(delta)))."""
    return se3_compose(T, se3_exp(delta))


def so3_retract(R, delta):
    return so3_compose(R, so3_exp(delta))


# ----------------------------------------------------------------------------- SE2 / SO2
def so2_exp(theta):
    """theseus/geometry/so2.py:167-181: [cos, sin]."""
    return np.stack([np.cos(theta), np.sin(theta)], axis=-1).astype(theta.dtype)


def se2_exp(xi):
    """theseus/geometry/se2.py:239-300 (exp_map). Storage [x, y, cos, sin]; tangent [ux, uy, theta]."""
    dt = xi.dtype
    u = xi[..., :2]
    theta = xi[..., 2]
    cosine, sine = np.cos(theta), np.sin(theta)
    small = np.abs(theta) < _EPS_TH[np.dtype(dt)]["se2_near_zero"]
    non_zero = np.ones((), dtype=dt)
    theta_nz = np.where(small, non_zero, theta)
    sine_by_theta = np.where(small, 1 - theta**2 / 6, sine / theta_nz)
    cosine_minus_one_by_theta = np.where(small, -theta / 2 + theta**3 / 24, (cosine - 1) / theta_nz)
    x = sine_by_theta * u[..., 0] + cosine_minus_one_by_theta * u[..., 1]
    y = sine_by_theta * u[..., 1] - cosine_minus_one_by_theta * u[..., 0]
    return np.stack([x, y, cosine, sine], axis=-1).astype(dt)


def se2_log(T):
    """theseus/geometry/se2.py:165-228 (_log_map_impl)."""
    dt = T.dtype
    cosine, sine = T[..., 2], T[..., 3]
    theta = np.arctan2(sine, cosine)
    small = np.abs(theta) < _EPS_TH[np.dtype(dt)]["se2_near_zero"]
    non_zero = np.ones((), dtype=dt)
    sine_nz = np.where(small, non_zero, sine)
    a = 0.5 * (1 + cosine) * np.where(small, 1 + sine**2 / 6, theta / sine_nz)
    b = 0.5 * theta
    ux = a * T[..., 0] + b * T[..., 1]
    uy = a * T[..., 1] - b * T[..., 0]
    return np.stack([ux, uy, theta], axis=-1).astype(dt)


def se2_compose(T0, T1):
    """theseus/geometry/se2.py:318-332."""
    c0, s0 = T0[..., 2], T0[..., 3]
    c1, s1 = T1[..., 2], T1[..., 3]
    x = c0 * T1[..., 0] - s0 * T1[..., 1] + T0[..., 0]
    y = s0 * T1[..., 0] + c0 * T1[..., 1] + T0[..., 1]
    return np.stack([x, y, c0 * c1 - s0 * s1, s0 * c1 + c0 * s1], axis=-1)


def se2_inverse(T):
    """theseus/geometry/se2.py:334-339."""
    c, s = T[..., 2], T[..., 3]
    x = -(c * T[..., 0] + s * T[..., 1])
    y = -(-s * T[..., 0] + c * T[..., 1])
    return np.stack([x, y, c, -s], axis=-1)


def se2_jlog(T):
    """theseus/geometry/se2.py:165-228 (_log_map_impl with jacobians).  Returns (J[...,3,3], xi)."""
    dt = T.dtype
    xi = se2_log(T)
    ux, uy, theta = xi[..., 0], xi[..., 1], xi[..., 2]
    cosine, sine = T[..., 2], T[..., 3]
    d_small = np.abs(theta) < _EPS_TH[np.dtype(dt)]["se2_d_near_zero"]
    one = np.ones((), dtype=dt)
    theta_nz = np.where(d_small, one, theta)
    omc_nz = np.where(d_small, one, 1 - cosine)
    half = 0.5 * theta
    a = np.where(d_small, 1 - theta**2 / 12.0, half * sine / omc_nz)
    coeff = np.where(d_small, theta / 12.0 + theta**3 / 720.0, 1.0 / theta_nz - 0.5 * sine / omc_nz)
    J = np.zeros(T.shape[:-1] + (3, 3), dtype=dt)
    J[..., 0, 0] = a
    J[..., 1, 1] = a
    J[..., 0, 1] = -half
    J[..., 1, 0] = half
    J[..., 0, 2] = coeff * ux + 0.5 * uy
    J[..., 1, 2] = coeff * uy - 0.5 * ux
    J[..., 2, 2] = 1
    return J, xi


def se2_adjoint(T):
    """theseus/geometry/se2.py:309-316."""
    out = np.zeros(T.shape[:-1] + (3, 3), dtype=T.dtype)
    c, s = T[..., 2], T[..., 3]
    out[..., 0, 0], out[..., 0, 1], out[..., 1, 0], out[..., 1, 1] = c, -s, s, c
    out[..., 0, 2] = T[..., 1]
    out[..., 1, 2] = -T[..., 0]
    out[..., 2, 2] = 1
    return out


def se2_retract(T, delta):
    return se2_compose(T, se2_exp(delta))


# ---- SO2: storage [cos, sin], tangent [theta] (theseus/geometry/so2.py) ----
def so2_exp(theta):
    """so2.py:99-100, :167-186."""
    return np.concatenate([np.cos(theta), np.sin(theta)], axis=-1)


def so2_log(X):
    """so2.py:206-222: atan2(sin, cos)."""
    return np.arctan2(X[..., 1], X[..., 0])[..., None]


def so2_compose(A, B):
    """so2.py:224-230."""
    return np.stack([A[..., 0] * B[..., 0] - A[..., 1] * B[..., 1], A[..., 1] * B[..., 0] + A[..., 0] * B[..., 1]], axis=-1)


def so2_inverse(X):
    """so2.py:232-234."""
    return np.stack([X[..., 0], -X[..., 1]], axis=-1)


def so2_jlog(X):
    """so2.py:209-219: the log Jacobian is 1.  Returns (J[...,1,1], theta)."""
    return np.ones(X.shape[:-1] + (1, 1), dtype=X.dtype), so2_log(X)


def so2_adjoint(X):
    """so2.py:116-117."""
    return np.ones(X.shape[:-1] + (1, 1), dtype=X.dtype)


def so2_retract(X, delta):
    return so2_compose(X, so2_exp(delta))
