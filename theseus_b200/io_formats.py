"""On-disk problem formats either side of the NLS path, host side only:

* g2o pose graphs  -- theseus/utils/examples/pose_graph/dataset.py:35-104 (read_3D_g2o_file), :110-172 (read_2D_g2o_file)
* BAL bundle adjustment files -- theseus/utils/examples/bundle_adjustment/data.py:166-207 (load_bal_dataset), :209-240 (save),
  camera parameter block [rodrigues(3), t(3), f, k1, k2] (data.py:43-60)

The readers return the same objects as the reference's (`PoseGraphEdge` with `.i .j .relative_pose .weight`, `Camera`,
`Observation`), built from this package's geometry classes on the CPU; `objective.to("cuda")` moves them afterwards.
"""
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import lie_torch
from .core import DiagonalCostWeight
from .geometry import Point2, Point3, SE2, SE3, Variable, Vector


def quaternion_to_rotation(q: torch.Tensor) -> torch.Tensor:
    """[..., 4] = (w, x, y, z) -> [..., 3, 3]  (torchlie so3_impl.py:821-852)."""
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    q00, q01, q02, q03 = w * w, w * x, w * y, w * z
    q11, q12, q13, q22, q23, q33 = x * x, x * y, x * z, y * y, y * z, z * z
    return torch.stack((torch.stack((q00 + q11 - q22 - q33, 2 * (q12 - q03), 2 * (q13 + q02)), -1),
                        torch.stack((2 * (q12 + q03), q00 - q11 + q22 - q33, 2 * (q23 - q01)), -1),
                        torch.stack((2 * (q13 - q02), 2 * (q23 + q01), q00 - q11 - q22 + q33), -1)), -2)


def x_y_z_unit_quaternion_to_SE3(x_y_z_quaternion: torch.Tensor, name: Optional[str] = None) -> SE3:
    """[B,7] = (x, y, z, qw, qx, qy, qz) -> SE3  (theseus/geometry/se3.py:128-144)."""
    if x_y_z_quaternion.ndim == 1:
        x_y_z_quaternion = x_y_z_quaternion.unsqueeze(0)
    if x_y_z_quaternion.ndim != 2 or x_y_z_quaternion.shape[1] != 7:
        raise ValueError("x_y_z_quaternion can only be 7-D vectors.")
    R = quaternion_to_rotation(x_y_z_quaternion[:, 3:])
    return SE3(tensor=torch.cat((R, x_y_z_quaternion[:, :3, None]), dim=2), name=name)


class PoseGraphEdge:
    """theseus/utils/examples/pose_graph/dataset.py:14-30."""

    def __init__(self, i: int, j: int, relative_pose, weight: Optional[DiagonalCostWeight] = None):
        self.i, self.j, self.relative_pose, self.weight = i, j, relative_pose, weight

    def to(self, *args, **kwargs):
        self.weight.to(*args, **kwargs)
        self.relative_pose.to(*args, **kwargs)


def read_3D_g2o_file(path: str, dtype: Optional[torch.dtype] = None) -> Tuple[int, List[SE3], List[PoseGraphEdge]]:
    """VERTEX_SE3:QUAT id x y z qx qy qz qw ; EDGE_SE3:QUAT i j x y z qx qy qz qw + upper-triangular 6x6 information (21 values),
    whose diagonal's square root becomes the DiagonalCostWeight (dataset.py:35-104)."""
    dtype = dtype or torch.get_default_dtype()
    num_vertices, verts, edges = 0, {}, []

    def pose(tokens):
        v = torch.from_numpy(np.array([tokens], dtype=np.float64)).to(dtype)
        v[:, 3:] /= torch.norm(v[:, 3:], dim=1)
        v[:, 3:] = v[:, [6, 3, 4, 5]]  # (qx,qy,qz,qw) on disk -> (qw,qx,qy,qz)
        return v

    with open(path, "r") as f:
        for line in f:
            tokens = line.split()
            if not tokens:
                continue
            if tokens[0] == "EDGE_SE3:QUAT":
                i, j, n = int(tokens[1]), int(tokens[2]), len(edges)
                rel = x_y_z_unit_quaternion_to_SE3(pose(tokens[3:10]), name=f"EDGE_SE3__{n}")
                w = torch.from_numpy(np.array(tokens[10:], dtype=np.float64)[[0, 6, 11, 15, 18, 20]]).to(dtype).sqrt().view(1, -1)
                edges.append(PoseGraphEdge(i, j, rel, DiagonalCostWeight(Variable(w), name=f"EDGE_WEIGHT__{n}")))
                num_vertices = max(num_vertices, i, j)
            elif tokens[0] == "VERTEX_SE3:QUAT":
                i = int(tokens[1])
                verts[i] = pose(tokens[2:])
                num_vertices = max(num_vertices, i)
    vertices = [x_y_z_unit_quaternion_to_SE3(q, name=f"VERTEX_SE3__{i}") for i, q in sorted(verts.items())]
    return num_vertices + 1, vertices, edges


def read_2D_g2o_file(path: str, dtype: Optional[torch.dtype] = None) -> Tuple[int, List[SE2], List[PoseGraphEdge]]:
    """VERTEX_SE2 id x y theta ; EDGE_SE2 i j x y theta + upper-triangular 3x3 information (dataset.py:110-172; the reference's
    `np.array(1, tokens[6:], ...)` there is a typo for `np.array(tokens[6:], ...)`, which is what this does)."""
    dtype = dtype or torch.get_default_dtype()
    num_vertices, verts, edges = 0, {}, []
    with open(path, "r") as f:
        for line in f:
            tokens = line.split()
            if not tokens:
                continue
            if tokens[0] == "EDGE_SE2":
                i, j, n = int(tokens[1]), int(tokens[2]), len(edges)
                xyt = torch.from_numpy(np.array([tokens[3:6]], dtype=np.float64)).to(dtype)
                w = torch.from_numpy(np.array(tokens[6:], dtype=np.float64)[[0, 3, 5]]).to(dtype).sqrt().view(1, -1)
                edges.append(PoseGraphEdge(i, j, SE2(x_y_theta=xyt, name=f"EDGE_SE2__{n}"), DiagonalCostWeight(Variable(w), name=f"EDGE_WEIGHT__{n}")))
                num_vertices = max(num_vertices, i, j)
            elif tokens[0] == "VERTEX_SE2":
                i = int(tokens[1])
                verts[i] = torch.from_numpy(np.array([tokens[2:]], dtype=np.float64)).to(dtype)
                num_vertices = max(num_vertices, i)
    vertices = [SE2(x_y_theta=v, name=f"VERTEX_SE2__{i}") for i, v in sorted(verts.items())]
    return num_vertices + 1, vertices, edges


class Camera:
    """theseus/utils/examples/bundle_adjustment/data.py:14-60: pose + focal length + two radial distortion coefficients."""

    def __init__(self, pose: SE3, focal_length: Vector, calib_k1: Vector, calib_k2: Vector):
        self.pose, self.focal_length, self.calib_k1, self.calib_k2 = pose, focal_length, calib_k1, calib_k2

    @staticmethod
    def from_params(params: List[float], name: str = "Cam") -> "Camera":
        r = lie_torch._so3_exp_parts(torch.tensor(params[:3], dtype=torch.float64).unsqueeze(0))[0]
        t = torch.tensor([params[3:6]], dtype=torch.float64).unsqueeze(2)
        return Camera(SE3(tensor=torch.cat([r, t], dim=2), name=name + "_pose"),
                      Vector(tensor=torch.tensor([params[6:7]], dtype=torch.float64), name=name + "_focal_length"),
                      Vector(tensor=torch.tensor([params[7:8]], dtype=torch.float64), name=name + "_calib_k1"),
                      Vector(tensor=torch.tensor([params[8:9]], dtype=torch.float64), name=name + "_calib_k2"))


class Observation:
    """data.py:122-131."""

    def __init__(self, camera_index: int, point_index: int, image_feature_point: Point2):
        self.camera_index, self.point_index, self.image_feature_point = camera_index, point_index, image_feature_point


def load_bal_dataset(path: str):
    """BAL text format: header `num_cameras num_points num_observations`, then the observations `cam pt u v`, then 9 lines per camera,
    then 3 lines per point (data.py:166-207)."""
    observations, cameras, points = [], [], []
    with open(path, "rt") as f:
        num_cameras, num_points, num_observations = [int(x) for x in f.readline().rstrip().split()]
        for i in range(num_observations):
            fields = f.readline().rstrip().split()
            feat = Point2(tensor=torch.tensor([float(fields[2]), float(fields[3])], dtype=torch.float64).unsqueeze(0), name=f"Feat{i}")
            observations.append(Observation(int(fields[0]), int(fields[1]), feat))
        for i in range(num_cameras):
            cameras.append(Camera.from_params([float(f.readline().rstrip()) for _ in range(9)], name=f"Cam{i}"))
        for i in range(num_points):
            points.append(Point3(tensor=torch.tensor([float(f.readline().rstrip()) for _ in range(3)], dtype=torch.float64).unsqueeze(0),
                                 name=f"Pt{i}"))
    return cameras, points, observations
