"""Synthetic problem generators (host-side INPUT generation only -- not part of the compute path).

`pose_graph_synthetic_3d` follows the recipe of the reference's
theseus/utils/examples/pose_graph/dataset.py:238-365 (PoseGraphDataset.generate_synthetic_3D): a random walk
of relative poses ~ exp(U[-1/2,1/2]^3 x U[-1,1]^3), odometry edges with uniform noise, random loop closures
with probability `loop_closure_ratio` (1..max_num_loop_closures edges to earlier poses), initial poses = ground
truth perturbed by the same noise model.  The graph structure is shared by the batch; values differ per item.
Plain torch on the CPU (it runs once, before the timed region); the SE3 exponential here is a textbook Rodrigues
formula used only to fabricate inputs.
"""
from typing import List, Tuple

import numpy as np
import torch


def _exp_se3(xi: torch.Tensor) -> torch.Tensor:
    v, w = xi[:, :3], xi[:, 3:]
    th = w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = w / th
    K = torch.zeros(xi.shape[0], 3, 3, dtype=xi.dtype)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    s, c = torch.sin(th).unsqueeze(2), torch.cos(th).unsqueeze(2)
    eye = torch.eye(3, dtype=xi.dtype).expand_as(K)
    KK = K @ K
    R = eye + s * K + (1 - c) * KK
    t3 = th.unsqueeze(2)
    V = eye + ((1 - c) / t3) * K + ((t3 - s) / t3) * KK
    t = (V @ v.unsqueeze(2))
    return torch.cat([R, t], dim=2)


def _compose(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    R = a[:, :, :3] @ b[:, :, :3]
    t = a[:, :, :3] @ b[:, :, 3:] + a[:, :, 3:]
    return torch.cat([R, t], dim=2)


def _inverse(a: torch.Tensor) -> torch.Tensor:
    Rt = a[:, :, :3].transpose(1, 2)
    return torch.cat([Rt, -(Rt @ a[:, :, 3:])], dim=2)


def pose_graph_synthetic_3d(num_poses: int, batch_size: int, translation_noise: float = 0.05, rotation_noise: float = 0.02,
                            loop_closure_ratio: float = 0.2, max_num_loop_closures: int = 10, seed: int = 0,
                            dtype: torch.dtype = torch.float64):
    """Returns dict(poses [N,B,3,4], gt_poses [N,B,3,4], edges [(i,j)], meas [E,B,3,4], info [6])."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    B = batch_size

    def urand(n, scale_t, scale_r):
        u = 2.0 * torch.rand(n, 6, generator=gen, dtype=dtype) - 1.0
        return torch.cat([u[:, :3] * scale_t, u[:, 3:] * scale_r], dim=1)

    eye = torch.eye(3, 4, dtype=dtype).repeat(B, 1, 1)
    gt: List[torch.Tensor] = [eye.clone()]
    poses: List[torch.Tensor] = [eye.clone()]
    edges: List[Tuple[int, int]] = []
    meas: List[torch.Tensor] = []
    for n in range(1, num_poses):
        gt_rel = _exp_se3(urand(B, 0.5, 1.0))
        rel = _compose(gt_rel, _exp_se3(urand(B, translation_noise, rotation_noise)))
        gt.append(_compose(gt[-1], gt_rel))
        poses.append(_compose(poses[-1], rel))
        edges.append((n - 1, n))
        meas.append(rel)
        if rng.random() <= loop_closure_ratio and n - 1 > 0:
            k = int(rng.integers(max_num_loop_closures)) + 1
            for i in sorted(set(int(x) for x in rng.integers(0, n - 1, k))):
                gt_rel = _compose(_inverse(gt[i]), gt[n])
                rel = _compose(gt_rel, _exp_se3(urand(1, translation_noise, rotation_noise)).expand(B, 3, 4))
                edges.append((i, n))
                meas.append(rel)
    for i in range(num_poses):
        poses[i] = _compose(gt[i], _exp_se3(urand(1, translation_noise, rotation_noise)).expand(B, 3, 4))
    info = torch.tensor([1 / translation_noise] * 3 + [1 / rotation_noise] * 3, dtype=dtype)
    return dict(poses=torch.stack(poses, 0).contiguous(), gt_poses=torch.stack(gt, 0).contiguous(), edges=edges,
                meas=torch.stack(meas, 0).contiguous(), info=info)


def build_pose_graph_objective(th, data, device, prior_weight: float = 1e-3, batch_slice=None):
    """Objective exactly as examples/pose_graph/pose_graph_cube.py:56-83: one Between per edge with a (batch-1)
    DiagonalCostWeight, plus a Difference prior on pose 0 with ScaleCostWeight(reg_w).  Returns (objective, poses)."""
    dtype = data["poses"].dtype
    sl = batch_slice if batch_slice is not None else slice(None)
    P, M = data["poses"][:, sl], data["meas"][:, sl]
    poses = [th.SE3(tensor=P[i].to(device), name=f"VERTEX_SE3__{i}") for i in range(P.shape[0])]
    objective = th.Objective(dtype=dtype)
    w = data["info"].view(1, 6).to(device)
    for e, (i, j) in enumerate(data["edges"]):
        z = th.SE3(tensor=M[e].to(device), name=f"EDGE_SE3__{i}_{j}__{e}")
        objective.add(th.Between(poses[i], poses[j], z, th.DiagonalCostWeight(th.Variable(w.clone(), name=f"EDGE_WEIGHT__{i}_{j}__{e}")),
                                 name=f"between__{e}"))
    prior = th.Difference(poses[0], th.SE3(tensor=P[0].clone().to(device), name="VERTEX_SE3__0__PRIOR"),
                          th.ScaleCostWeight(th.Variable(torch.tensor([[prior_weight]], dtype=dtype, device=device), name="PRIOR_WEIGHT")),
                          name="pose_prior")
    objective.add(prior)
    objective.to(device)
    return objective, poses


def _exp_se3_nd(xi: torch.Tensor) -> torch.Tensor:
    """_exp_se3 over any leading dimensions, on xi's device."""
    shp = xi.shape[:-1]
    out = _exp_se3_dev(xi.reshape(-1, 6))
    return out.view(*shp, 3, 4)


def _exp_se3_dev(xi: torch.Tensor) -> torch.Tensor:
    v, w = xi[:, :3], xi[:, 3:]
    th = w.norm(dim=1, keepdim=True).clamp_min(1e-12)
    k = w / th
    K = torch.zeros(xi.shape[0], 3, 3, dtype=xi.dtype, device=xi.device)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 2], k[:, 1], k[:, 2], -k[:, 0], -k[:, 1], k[:, 0]
    s, c = torch.sin(th).unsqueeze(2), torch.cos(th).unsqueeze(2)
    eye = torch.eye(3, dtype=xi.dtype, device=xi.device).expand_as(K)
    KK = K @ K
    R = eye + s * K + (1 - c) * KK
    t3 = th.unsqueeze(2)
    V = eye + ((1 - c) / t3) * K + ((t3 - s) / t3) * KK
    return torch.cat([R, V @ v.unsqueeze(2)], dim=2)


def _compose_nd(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    R = a[..., :3] @ b[..., :3]
    t = a[..., :3] @ b[..., 3:] + a[..., 3:]
    return torch.cat([R, t], dim=-1)


def _inverse_nd(a: torch.Tensor) -> torch.Tensor:
    Rt = a[..., :3].transpose(-1, -2)
    return torch.cat([Rt, -(Rt @ a[..., 3:])], dim=-1)


def pose_graph_sphere(rings: int, per_ring: int, batch_size: int, translation_noise: float = 0.05, rotation_noise: float = 0.02,
                      radius: float = 10.0, seed: int = 0, dtype: torch.dtype = torch.float64, device="cpu"):
    """sphere2500-like topology: `rings` x `per_ring` poses on a sphere, odometry chain
    i -> i+1 plus ring-to-ring edges i -> i+per_ring: E = N-1 + N-per_ring (2500 nodes -> 4949 edges).
    Measurements = ground-truth relative pose composed with uniform noise (the noise model of generate_synthetic_3D);
    initial poses = ground truth composed with the same noise.  Vectorised over poses / edges; `device` only says where the
    fabrication runs (the tensors are returned on the CPU: inputs start in host memory)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    N, B = rings * per_ring, batch_size

    def urand(shape, scale_t, scale_r):
        u = 2.0 * torch.rand(*shape, 6, generator=gen, dtype=dtype, device=dev) - 1.0
        return torch.cat([u[..., :3] * scale_t, u[..., 3:] * scale_r], dim=-1)

    idx = torch.arange(N, device=dev)
    ring, k = idx // per_ring, idx % per_ring
    phi = (ring.to(dtype) + 0.5) / rings * np.pi           # polar angle
    theta = k.to(dtype) / per_ring * 2 * np.pi
    pos = radius * torch.stack([torch.sin(phi) * torch.cos(theta), torch.sin(phi) * torch.sin(theta), torch.cos(phi)], dim=1)
    R = torch.zeros(N, 3, 3, dtype=dtype, device=dev)     # orientation: yaw along the ring direction
    c, s = torch.cos(theta), torch.sin(theta)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, -s, s, c, 1.0
    gt1 = torch.cat([R, pos.unsqueeze(2)], dim=2)                               # [N,3,4]
    # per-batch-item ground truth = shared shape perturbed a little per item
    gt = _compose_nd(gt1.unsqueeze(1).expand(N, B, 3, 4), _exp_se3_nd(urand((N, B), 0.2, 0.05)))
    edges = [(i, i + 1) for i in range(N - 1)] + [(i, i + per_ring) for i in range(N - per_ring)]
    ei = torch.tensor([e[0] for e in edges], device=dev)
    ej = torch.tensor([e[1] for e in edges], device=dev)
    meas = torch.empty(len(edges), B, 3, 4, dtype=dtype, device=dev)
    CH = 512                                                                     # chunks of edges bound the temporaries
    for c0 in range(0, len(edges), CH):
        a, b_ = ei[c0:c0 + CH], ej[c0:c0 + CH]
        rel = _compose_nd(_inverse_nd(gt[a]), gt[b_])
        meas[c0:c0 + CH] = _compose_nd(rel, _exp_se3_nd(urand((a.shape[0], B), translation_noise, rotation_noise)))
    poses = _compose_nd(gt, _exp_se3_nd(urand((N, B), translation_noise, rotation_noise)))
    info = torch.tensor([1 / translation_noise] * 3 + [1 / rotation_noise] * 3, dtype=dtype)
    return dict(poses=poses.contiguous().cpu(), gt_poses=gt.contiguous().cpu(), edges=edges, meas=meas.contiguous().cpu(), info=info)
