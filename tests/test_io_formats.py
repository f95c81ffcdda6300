"""g2o / BAL readers (host side) against what the reference's own readers return for the same files
(tests/golden/io_small.g2o, io_small_bal.txt -> io_kat.npz, written by make_golden.py io)."""
import os

import numpy as np
import torch

from theseus_b200 import io_formats as io
from helpers import load

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_read_3d_g2o_matches_reference():
    g = load("io_kat")
    n, verts, edges = io.read_3D_g2o_file(os.path.join(HERE, "io_small.g2o"), dtype=torch.float64)
    assert n == int(g["g2o_n"]) and len(verts) == g["g2o_verts"].shape[0] and len(edges) == g["g2o_edge_ij"].shape[0]
    np.testing.assert_allclose(np.concatenate([v.tensor.numpy() for v in verts], 0), g["g2o_verts"], rtol=0, atol=1e-15)
    assert np.array_equal(np.array([[e.i, e.j] for e in edges]), g["g2o_edge_ij"])
    np.testing.assert_allclose(np.concatenate([e.relative_pose.tensor.numpy() for e in edges], 0), g["g2o_edge_pose"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(np.concatenate([e.weight.diagonal.tensor.numpy() for e in edges], 0), g["g2o_edge_w"], rtol=1e-15)
    assert [v.name for v in verts] == [f"VERTEX_SE3__{i}" for i in range(n)] and edges[2].relative_pose.name == "EDGE_SE3__2"


def test_load_bal_matches_reference():
    g = load("io_kat")
    cams, pts, obs = io.load_bal_dataset(os.path.join(HERE, "io_small_bal.txt"))
    np.testing.assert_allclose(np.concatenate([c.pose.tensor.numpy() for c in cams], 0), g["bal_cam_pose"], rtol=0, atol=1e-15)
    for k, attr in (("bal_cam_f", "focal_length"), ("bal_cam_k1", "calib_k1"), ("bal_cam_k2", "calib_k2")):
        assert np.array_equal(np.concatenate([getattr(c, attr).tensor.numpy() for c in cams], 0), g[k])
    assert np.array_equal(np.concatenate([p.tensor.numpy() for p in pts], 0), g["bal_pts"])
    assert np.array_equal(np.array([[o.camera_index, o.point_index] for o in obs]), g["bal_obs"])
    assert np.array_equal(np.concatenate([o.image_feature_point.tensor.numpy() for o in obs], 0), g["bal_feat"])


def test_read_2d_g2o(tmp_path):
    p = tmp_path / "t.g2o"
    p.write_text("VERTEX_SE2 0 0 0 0\nVERTEX_SE2 1 1.0 0.1 0.2\nEDGE_SE2 0 1 1.0 0.0 0.2 100 0 0 100 0 400\n")
    n, verts, edges = io.read_2D_g2o_file(str(p), dtype=torch.float64)
    assert n == 2 and len(verts) == 2 and len(edges) == 1
    np.testing.assert_allclose(verts[1].tensor.numpy(), [[1.0, 0.1, np.cos(0.2), np.sin(0.2)]], atol=1e-15)
    np.testing.assert_allclose(edges[0].weight.diagonal.tensor.numpy(), [[10.0, 10.0, 20.0]])
