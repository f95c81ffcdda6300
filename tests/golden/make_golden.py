"""Generate golden fixtures by running the REFERENCE ITSELF (facebookresearch/theseus v0.2.3).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes small .npz files next to this script.  The fixtures pin oracle/ (tests/test_oracle_*.py,
CPU) and the CUDA path (tests/test_gpu_*.py).  At test time only the pure problem builders backward_problem / backward_inputs /
BACKWARD_CASES are loaded from this file (so generator and test build the same problem); the reference is imported under __main__ only.
    python tests/golden/make_golden.py backward     # regenerates backward_kat.npz alone
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    # torchkin's vendored URDF parser imports lxml at import time; stub it.
    lxml = types.ModuleType("lxml")
    etree = types.ModuleType("lxml.etree")
    import xml.etree.ElementTree as ET
    for k in dir(ET):
        setattr(etree, k, getattr(ET, k))
    lxml.etree = etree
    sys.modules["lxml"] = lxml
    sys.modules["lxml.etree"] = etree
    for p in (REF, REF + "/torchlie", REF + "/torchkin"):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import theseus as th
    import torchlie.functional as lieF
    return th, lieF


def _angle_sweep_tangents(rng, dtype, dof):
    """Random tangents plus the reference's own angle sweep (tests/theseus_tests/geometry/test_se3.py:56-129)."""
    out = []
    out.append(rng.standard_normal((20, dof)))
    for ang in [0.0, 1e-9, 1e-5, 3e-3, 4.9e-3, 5.1e-3, 9e-3, 1.1e-2, 0.19, 0.21, 1.0,
                np.pi - 1e-3, np.pi - 1e-7, np.pi - 1e-11, np.pi, 2 * np.pi - 1e-11, 2 * np.pi - 1e-3]:
        for _ in range(3):
            ax = rng.standard_normal(3)
            ax /= np.linalg.norm(ax)
            t = np.zeros(dof)
            t[-3:] = ax * ang
            if dof == 6:
                t[:3] = rng.standard_normal(3)
            out.append(t[None])
    # axis-aligned near-pi cases to exercise the major-diagonal selection
    for k in range(3):
        for sgn in (+1, -1):
            t = np.zeros(dof)
            t[-3 + k] = sgn * (np.pi - 1e-9)
            if dof == 6:
                t[:3] = rng.standard_normal(3)
            out.append(t[None])
    return np.concatenate(out, 0).astype(dtype)


def make_lie(th, lieF):
    import torch
    rng = np.random.default_rng(0)
    out = {}
    for dtname, dt, tdt in (("f64", np.float64, torch.float64), ("f32", np.float32, torch.float32)):
        from torchlie.functional import so3_impl, se3_impl
        for gname, dof, F in (("so3", 3, so3_impl), ("se3", 6, se3_impl)):
            tang = _angle_sweep_tangents(rng, dt, dof)
            t = torch.from_numpy(tang)
            G = F._exp_impl(t)
            (jexp,), _ = F._jexp_impl(t)
            (jlog,), logG = F._jlog_impl(G)
            G2 = F._exp_impl(torch.from_numpy(rng.standard_normal(tang.shape).astype(dt)))
            pre = f"{gname}_{dtname}_"
            out[pre + "tangent"] = tang
            out[pre + "exp"] = G.numpy()
            out[pre + "jexp"] = jexp.numpy()
            out[pre + "log"] = logG.numpy()
            out[pre + "jlog"] = jlog.numpy()
            out[pre + "adj"] = F._adjoint_impl(G).numpy()
            out[pre + "inv"] = F._inverse_impl(G).numpy()
            out[pre + "other"] = G2.numpy()
            out[pre + "compose"] = F._compose_impl(G, G2).numpy()
    np.savez_compressed(os.path.join(HERE, "lie_kat.npz"), **out)
    print("lie_kat.npz", len(out), "arrays")


def make_costs(th):
    import torch
    torch.manual_seed(1)
    out = {}
    for dtname, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        B = 16
        X0 = th.SE3.rand(B, dtype=tdt)
        X1 = th.SE3.rand(B, dtype=tdt)
        Z = th.SE3.rand(B, dtype=tdt)
        w = torch.rand(1, 6, dtype=tdt) + 0.5
        cf = th.Between(X0, X1, Z, th.DiagonalCostWeight(w))
        (J0, J1), e = cf.weighted_jacobians_error()
        cl = th.Difference(X0, Z, th.ScaleCostWeight(torch.tensor(0.37, dtype=tdt)))
        (Jl,), el = cl.weighted_jacobians_error()
        pre = f"{dtname}_"
        out.update({pre + "X0": X0.tensor.numpy(), pre + "X1": X1.tensor.numpy(), pre + "Z": Z.tensor.numpy(),
                    pre + "w": w.numpy(), pre + "between_J0": J0.numpy(), pre + "between_J1": J1.numpy(),
                    pre + "between_e": e.numpy(), pre + "local_J": Jl.numpy(), pre + "local_e": el.numpy()})
    np.savez_compressed(os.path.join(HERE, "costs_kat.npz"), **out)
    print("costs_kat.npz", len(out), "arrays")


def _pgo_objective(th, num_poses, B, seed, dtype, loop_closure_ratio=0.2, init_perturb=0.0, robust=None, outlier_ratio=0.0):
    import torch
    from theseus.utils.examples.pose_graph.dataset import PoseGraphDataset
    torch.manual_seed(seed)
    np.random.seed(seed)
    rng = torch.Generator().manual_seed(seed)
    pg, _ = PoseGraphDataset.generate_synthetic_3D(
        num_poses=num_poses, translation_noise=0.05, rotation_noise=0.02, loop_closure_ratio=loop_closure_ratio,
        loop_closure_outlier_ratio=outlier_ratio, dataset_size=B, batch_size=B, generator=rng, dtype=dtype)
    if init_perturb > 0:
        # harder start (so that LM needs all its iterations and really rejects steps): extra random right-perturbation
        for p in pg.poses[1:]:
            p.tensor = p.compose(th.SE3.exp_map(init_perturb * (2 * torch.rand(B, 6, dtype=dtype) - 1))).tensor
    # objective exactly as examples/pose_graph/pose_graph_cube.py:56-83
    objective = th.Objective(dtype=dtype)
    log_loss_radius = th.Vector(tensor=torch.tensor([[0.5]], dtype=dtype), name="log_loss_radius")
    for edge in pg.edges:
        cf = th.Between(pg.poses[edge.i], pg.poses[edge.j], edge.relative_pose, edge.weight)
        if robust == "welsch":  # as examples/pose_graph/pose_graph_synthetic.py builds its robust relative-pose costs
            cf = th.RobustCostFunction(cf, th.WelschLoss, log_loss_radius, name=f"robust_{cf.name}")
        elif robust == "geman":  # graduated non-convexity wrapper, control value 3
            if not objective.has_aux_var("gnc_mu"):
                gnc_mu = th.Vector(tensor=torch.tensor([[3.0]], dtype=dtype), name="gnc_mu")
            cf = th.GNCRobustCostFunction(cf, th.GemanMcClureLoss, log_loss_radius, gnc_mu, name=f"robust_{cf.name}")
        objective.add(cf)
    prior = th.Difference(var=pg.poses[0], cost_weight=th.ScaleCostWeight(torch.tensor(1e-3, dtype=dtype)),
                          target=pg.poses[0].copy(new_name=pg.poses[0].name + "__PRIOR"))
    objective.add(prior)
    return pg, objective


def make_pgo(th, name, num_poses, B, seed, iters, lm_kwargs, method="lm", full_trace=True, loop_closure_ratio=0.2,
             init_perturb=0.0, robust=None, outlier_ratio=0.0):
    import torch
    dtype = torch.float64
    pg, objective = _pgo_objective(th, num_poses, B, seed, dtype, loop_closure_ratio, init_perturb, robust, outlier_ratio)
    out_robust = robust
    out = {}
    out["poses0"] = np.stack([p.tensor.numpy() for p in pg.poses], 0)            # [N,B,3,4]
    out["edges"] = np.array([[e.i, e.j] for e in pg.edges], dtype=np.int64)        # [E,2]
    out["meas"] = np.stack([e.relative_pose.tensor.numpy() for e in pg.edges], 0)  # [E,B,3,4]
    out["edge_w"] = np.stack([e.weight.diagonal.tensor.numpy() for e in pg.edges], 0)  # [E,1,6]
    out["prior_w"] = np.array(1e-3)
    out["robust"] = np.array(robust or "")
    out["log_loss_radius"] = np.array([[0.5]])
    if robust == "geman":
        out["gnc_mu"] = np.array([[3.0]])
    cls = {"lm": th.LevenbergMarquardt, "gn": th.GaussNewton, "dogleg": th.Dogleg}[method]
    opt = cls(objective, linear_solver_cls=th.CholeskyDenseSolver, vectorize=True, max_iterations=iters,
              step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0)
    ordering = [v.name for v in opt.linear_solver.linearization.ordering]
    assert ordering == [p.name for p in pg.poses], "default ordering must be pose order"
    sp = th.SparseLinearization(objective)
    out["A_row_ptr"] = sp.A_row_ptr.astype(np.int64)
    out["A_col_ind"] = sp.A_col_ind.astype(np.int64)
    out["var_start_cols"] = np.array(sp.var_start_cols, dtype=np.int64)
    tr = dict(delta=[], lam=[], Atb=[], AtA_diag=[], A_val=[], b=[], AtA=[], err=[])

    def cb(optimizer, info, delta, it):
        lin = optimizer.linear_solver.linearization
        tr["delta"].append(delta.detach().numpy().copy())
        tr["err"].append(info.last_err.numpy().copy())  # fp64 (err_history is stored in fp32 by the reference)
        tr["Atb"].append(lin.Atb.squeeze(2).numpy().copy())
        tr["AtA_diag"].append(lin.AtA.diagonal(dim1=1, dim2=2).numpy().copy())
        if method == "lm":
            tr["lam"].append(np.array(optimizer._damping, dtype=np.float64) * np.ones(B))
        if method == "dogleg":
            tr["lam"].append(optimizer._trust_region.view(-1).numpy().copy())   # trace_lam = trust-region radius after the step
        if full_trace and it == 0:
            tr["AtA"].append(lin.AtA.numpy().copy())
            tr["b"].append(lin.b.numpy().copy())

    if full_trace:
        # sparse linearization at the initial point (A_val in the reference CSR layout)
        objective.update()
        sp.linearize()
        out["A_val0"] = sp.A_val.numpy().copy()
        out["b0"] = sp.b.numpy().copy()
        out["Atb0_sparse"] = sp.Atb.squeeze(2).numpy().copy()
    with torch.no_grad():
        info = opt.optimize(track_err_history=True, end_iter_callback=cb, **lm_kwargs)
    out["err_history"] = info.err_history.numpy()
    out["poses_final"] = np.stack([objective.optim_vars[n].tensor.numpy() for n in ordering], 0)
    for k, v in tr.items():
        if v:
            out["trace_" + k] = np.stack(v, 0)
    out["kwargs_json"] = np.array(repr(dict(method=method, iters=iters, **lm_kwargs)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "err0", out["err_history"][:, 0], "-> errK", out["err_history"][:, -1])


def make_dense_solver(th):
    """Random SPD systems in the style of tests/theseus_tests/optimizer/linear/test_dense_solver.py:12-77."""
    import torch
    torch.manual_seed(0)
    out = {}
    for n in (1, 6, 10, 48, 130):
        B = 5
        A = torch.randn(B, n + 3, n, dtype=torch.float64)
        AtA = A.transpose(1, 2) @ A + 0.1 * torch.eye(n, dtype=torch.float64)
        Atb = torch.randn(B, n, 1, dtype=torch.float64)
        lam = torch.rand(B, dtype=torch.float64) + 0.01
        for ell in (True, False):
            D = th.CholeskyDenseSolver._apply_damping(AtA, lam, ellipsoidal=ell, eps=1e-8)
            L = torch.linalg.cholesky(D)
            x = torch.cholesky_solve(Atb, L).squeeze(2)
            out[f"n{n}_x_{'ell' if ell else 'sph'}"] = x.numpy()
        out[f"n{n}_AtA"] = AtA.numpy()
        out[f"n{n}_Atb"] = Atb.numpy()
        out[f"n{n}_lam"] = lam.numpy()
    np.savez_compressed(os.path.join(HERE, "dense_solver_kat.npz"), **out)
    print("dense_solver_kat.npz")


def make_ba(th, name, num_cameras, num_points, B, seed, iters, robust=False, track_length=4):
    """Bundle adjustment as examples/bundle_adjustment.py:106-164 (Reprojection + reg priors + known-camera priors), without
    the Huber wrapper (robust losses are a 'next' row), batch built by re-perturbing the scene per item."""
    import random
    import torch
    import theseus.utils.examples as theg
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    ba = theg.BundleAdjustmentDataset.generate_synthetic(num_cameras=num_cameras, num_points=num_points, average_track_length=track_length,
                                                        track_locality=0.3, feat_random=1.5, prob_feat_is_outlier=0.0)
    dtype = torch.float64
    # batch: item 0 = the dataset, items 1.. = re-perturbed copies (cameras: Camera.perturbed; points: U[-0.2,0.2])
    cam_pose = [torch.cat([c.pose.tensor] + [gc.perturbed().pose.tensor for _ in range(B - 1)], 0) for c, gc in zip(ba.cameras, ba.gt_cameras)]
    pts = [torch.cat([p.tensor] + [gp.tensor + (torch.rand(1, 3, dtype=dtype) * 2 - 1) * 0.2 for _ in range(B - 1)], 0)
           for p, gp in zip(ba.points, ba.gt_points)]
    cams = [th.SE3(tensor=cam_pose[i], name=f"Cam{i}_pose") for i in range(num_cameras)]
    points = [th.Point3(tensor=pts[i], name=f"Pt{i}") for i in range(num_points)]
    objective = th.Objective(dtype=dtype)
    weight = th.ScaleCostWeight(torch.tensor(1.0, dtype=dtype))
    log_loss_radius = th.Vector(1, name="log_loss_radius", dtype=dtype)
    obs_ci, obs_pi, feats = [], [], []
    for o, obs in enumerate(ba.observations):
        cam = ba.cameras[obs.camera_index]
        cf = th.eb.Reprojection(camera_pose=cams[obs.camera_index], world_point=points[int(obs.point_index)],
                                focal_length=cam.focal_length, calib_k1=cam.calib_k1, calib_k2=cam.calib_k2,
                                image_feature_point=obs.image_feature_point, weight=weight, name=f"reproj_{o}")
        if robust:  # examples/bundle_adjustment.py:122-128 (HuberLoss, log_loss_radius = 0)
            cf = th.RobustCostFunction(cf, th.HuberLoss, log_loss_radius, name=f"robust_{cf.name}")
        objective.add(cf)
        obs_ci.append(obs.camera_index); obs_pi.append(int(obs.point_index)); feats.append(obs.image_feature_point.tensor.numpy())
    zero_point3 = th.Point3(dtype=dtype, name="zero_point")
    identity_se3 = th.SE3(dtype=dtype, name="zero_se3")
    w = np.sqrt(1e-4)
    damping_weight = th.ScaleCostWeight(w * torch.ones(1, dtype=dtype))
    reg_order = list(objective.optim_vars.keys())
    for vname in reg_order:
        var = objective.optim_vars[vname]
        objective.add(th.Difference(var, identity_se3 if isinstance(var, th.SE3) else zero_point3, damping_weight, name=f"reg_{vname}"))
    camera_weight = th.ScaleCostWeight(100 * torch.ones(1, dtype=dtype))
    known = [0, num_cameras - 1]
    for i in known:
        objective.add(th.Difference(cams[i], th.SE3(tensor=ba.gt_cameras[i].pose.tensor.clone(), name=f"Cam{i}_gt_pose"), camera_weight,
                                    name=f"camera_diff_{i}"))
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, vectorize=True, max_iterations=iters, step_size=1.0,
                                abs_err_tolerance=0, rel_err_tolerance=0)
    order = [v.name for v in opt.linear_solver.linearization.ordering]
    sp = th.SparseLinearization(objective)
    objective.update(); sp.linearize()
    out = dict(cam_pose0=np.stack([c.numpy() for c in cam_pose], 0), pts0=np.stack([p.numpy() for p in pts], 0),
               obs_cam=np.array(obs_ci), obs_pt=np.array(obs_pi), feats=np.stack(feats, 0),
               focal=np.stack([c.focal_length.tensor.numpy() for c in ba.cameras], 0),
               k1=np.stack([c.calib_k1.tensor.numpy() for c in ba.cameras], 0), k2=np.stack([c.calib_k2.tensor.numpy() for c in ba.cameras], 0),
               known=np.array(known), known_pose=np.stack([ba.gt_cameras[i].pose.tensor.numpy() for i in known], 0),
               order=np.array(order), reg_order=np.array(reg_order),
               A_row_ptr=sp.A_row_ptr.astype(np.int64), A_col_ind=sp.A_col_ind.astype(np.int64),
               A_val0=sp.A_val.numpy().copy(), b0=sp.b.numpy().copy(), robust=np.array("huber" if robust else ""))
    tr = dict(delta=[], err=[], lam=[])

    def cb(optimizer, info, delta, it):
        tr["delta"].append(delta.numpy().copy()); tr["err"].append(info.last_err.numpy().copy())
        tr["lam"].append(optimizer._damping.numpy().copy())
    with torch.no_grad():
        info = opt.optimize(track_err_history=True, end_iter_callback=cb, damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
    out["err_history"] = info.err_history.numpy()
    out["final"] = np.concatenate([objective.optim_vars[n].tensor.numpy().reshape(B, -1) for n in order], 1)
    for k, v in tr.items():
        out["trace_" + k] = np.stack(v, 0)
    out["kwargs_json"] = np.array(repr(dict(method="lm", iters=iters, damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "err", out["err_history"][:, 0], "->", out["trace_err"][-1], "n_obs", len(obs_ci))


def make_simple_example(th):
    """examples/simple_example.py:17-45 (config C1): y = v exp(x), one AutoDiff cost of dim 20, GaussNewton 10 iterations,
    default dtype float32, batch 1; forward pass only."""
    import torch
    torch.manual_seed(0)
    x_true = torch.linspace(-1, 1, 20).view(1, -1)
    y_true = 0.5 * torch.exp(x_true)
    x = th.Variable(torch.randn_like(x_true), name="x")
    y = th.Variable(y_true, name="y")
    v = th.Vector(1, name="v")

    def error_fn(optim_vars, aux_vars):
        xx, yy = aux_vars
        return yy.tensor - optim_vars[0].tensor * torch.exp(xx.tensor)

    objective = th.Objective()
    objective.add(th.AutoDiffCostFunction([v], error_fn, 20, aux_vars=[x, y], cost_weight=th.ScaleCostWeight(1.0)))
    opt = th.GaussNewton(objective, max_iterations=10)
    layer = th.TheseusLayer(opt)
    phi = x_true + 0.1 * torch.ones_like(x_true)
    with torch.no_grad():
        sol, info = layer.forward(input_tensors={"x": phi.clone(), "v": torch.ones(1, 1)},
                                  optimizer_kwargs=dict(track_err_history=True, track_state_history=True))
    np.savez_compressed(os.path.join(HERE, "simple_example.npz"), x=phi.numpy(), y=y_true.numpy(), v_final=sol["v"].numpy(),
                        err_history=info.err_history.numpy(), v_history=info.state_history["v"].numpy(),
                        converged_iter=info.converged_iter.numpy(), status=np.array([s.value for s in info.status]))
    print("simple_example v ->", sol["v"].numpy().ravel(), "converged_iter", info.converged_iter.numpy(), info.err_history.numpy()[0, :4])


def make_se2(th):
    """SE2 (config C4's group): exp/log/compose/inverse/adjoint KATs, Between/Difference Jacobians and a small 2-D pose-graph
    LM trace, all from the reference classes (theseus/geometry/se2.py, embodied/measurements/between.py)."""
    import torch
    torch.manual_seed(11)
    dt = torch.float64
    out = {}
    B = 24
    xi = torch.randn(B, 3, dtype=dt)
    xi[:6, 2] *= 1e-7   # near-zero branches (se2 eps 1e-6 / 1e-3)
    xi[6:10, 2] *= 1e-4
    G = th.SE2.exp_map(xi)
    H = th.SE2.exp_map(torch.randn(B, 3, dtype=dt))
    jl = []
    out.update(tangent=xi.numpy(), exp=G.tensor.numpy(), log=G.log_map(jacobians=jl).numpy(), jlog=jl[0].numpy(),
               adj=G.adjoint().numpy(), inv=G.inverse().tensor.numpy(), other=H.tensor.numpy(), compose=G.compose(H).tensor.numpy())
    Z = th.SE2.exp_map(torch.randn(B, 3, dtype=dt))
    w = torch.rand(1, 3, dtype=dt) + 0.5
    (J0, J1), e = th.Between(G, H, Z, th.DiagonalCostWeight(w)).weighted_jacobians_error()
    (Jl,), el = th.Difference(G, Z, th.ScaleCostWeight(torch.tensor(0.7, dtype=dt))).weighted_jacobians_error()
    out.update(Z=Z.tensor.numpy(), w=w.numpy(), between_J0=J0.numpy(), between_J1=J1.numpy(), between_e=e.numpy(),
               local_J=Jl.numpy(), local_e=el.numpy())
    # small 2-D pose graph: ring of N poses with odometry + a few chords, batch 3
    N, Bp = 7, 3
    gt = [th.SE2.exp_map(torch.zeros(Bp, 3, dtype=dt))]
    for i in range(1, N):
        gt.append(gt[-1].compose(th.SE2.exp_map(torch.tensor([[1.0, 0.1, 0.6]], dtype=dt).repeat(Bp, 1) + 0.1 * torch.randn(Bp, 3, dtype=dt))))
    edges = [(i, i + 1) for i in range(N - 1)] + [(0, 3), (2, 6), (1, 5)]
    poses = [th.SE2(tensor=gt[i].compose(th.SE2.exp_map(0.2 * torch.randn(Bp, 3, dtype=dt))).tensor, name=f"P{i}") for i in range(N)]
    objective = th.Objective(dtype=dt)
    meas = []
    for k, (i, j) in enumerate(edges):
        z = gt[i].inverse().compose(gt[j]).compose(th.SE2.exp_map(0.02 * torch.randn(Bp, 3, dtype=dt)))
        z = th.SE2(tensor=z.tensor, name=f"Z{k}")
        meas.append(z.tensor.numpy())
        objective.add(th.Between(poses[i], poses[j], z, th.DiagonalCostWeight(th.Variable(torch.tensor([[10.0, 10.0, 20.0]], dtype=dt), name=f"W{k}"))))
    objective.add(th.Difference(poses[0], th.SE2(tensor=gt[0].tensor.clone(), name="P0_prior"), th.ScaleCostWeight(torch.tensor(5.0, dtype=dt))))
    poses0 = np.stack([p.tensor.numpy().copy() for p in poses], 0)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, vectorize=True, max_iterations=6, abs_err_tolerance=0, rel_err_tolerance=0)
    tr = dict(delta=[], err=[])

    def cb(optimizer, info, delta, it):
        tr["delta"].append(delta.numpy().copy()); tr["err"].append(info.last_err.numpy().copy())
    objective.update()
    with torch.no_grad():
        info = opt.optimize(track_err_history=True, end_iter_callback=cb, damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
    out.update(pg_poses0=poses0, pg_edges=np.array(edges), pg_meas=np.stack(meas, 0), pg_prior=gt[0].tensor.numpy(),
               pg_err0=info.err_history[:, 0].numpy(), pg_trace_err=np.stack(tr["err"], 0), pg_trace_delta=np.stack(tr["delta"], 0),
               pg_final=np.stack([p.tensor.numpy() for p in poses], 0))
    np.savez_compressed(os.path.join(HERE, "se2_kat.npz"), **out)
    print("se2_kat.npz pose-graph err", out["pg_err0"], "->", out["pg_trace_err"][-1])


def make_so3(th):
    """SO3 Between / Difference Jacobians + a retract KAT from the reference (theseus/geometry/so3.py, torchlie so3_impl.py)."""
    import torch
    torch.manual_seed(21)
    dt = torch.float64
    B = 20
    X0, X1, Z = th.SO3.rand(B, dtype=dt), th.SO3.rand(B, dtype=dt), th.SO3.rand(B, dtype=dt)
    w = torch.rand(1, 3, dtype=dt) + 0.5
    (J0, J1), e = th.Between(X0, X1, Z, th.DiagonalCostWeight(w)).weighted_jacobians_error()
    (Jl,), el = th.Difference(X0, Z, th.ScaleCostWeight(torch.tensor(1.3, dtype=dt))).weighted_jacobians_error()
    delta = 0.3 * torch.randn(B, 3, dtype=dt)
    R = X0.retract(delta)
    np.savez_compressed(os.path.join(HERE, "so3_kat.npz"), X0=X0.tensor.numpy(), X1=X1.tensor.numpy(), Z=Z.tensor.numpy(), w=w.numpy(),
                        between_J0=J0.numpy(), between_J1=J1.numpy(), between_e=e.numpy(), local_J=Jl.numpy(), local_e=el.numpy(),
                        delta=delta.numpy(), retract=R.tensor.numpy())
    print("so3_kat.npz")


def backward_problem(th, torch, dtype=None):
    """Shared by the generator (th = reference) and tests/test_gpu_backward.py (th = theseus_b200): exponential curve fit
    y = a exp(b x) + c with optimisation variables ab (Vector 2), c (Vector 1), a data cost (AutoDiff, dim 12) and a prior on ab."""
    dtype = dtype or torch.float64
    N = 12
    ab = th.Vector(2, name="ab", dtype=dtype)
    c = th.Vector(1, name="c", dtype=dtype)
    x = th.Variable(torch.zeros(1, N, dtype=dtype), name="x")
    y = th.Variable(torch.zeros(1, N, dtype=dtype), name="y")
    t = th.Variable(torch.zeros(1, 2, dtype=dtype), name="t")

    def data_err(optim_vars, aux_vars):
        p, cc = optim_vars
        xx, yy = aux_vars
        return yy.tensor - (p.tensor[:, 0:1] * torch.exp(p.tensor[:, 1:2] * xx.tensor) + cc.tensor)

    def prior_err(optim_vars, aux_vars):
        return optim_vars[0].tensor - aux_vars[0].tensor

    objective = th.Objective(dtype=dtype)
    objective.add(th.AutoDiffCostFunction([ab, c], data_err, N, aux_vars=[x, y], cost_weight=th.ScaleCostWeight(torch.tensor(1.0, dtype=dtype)), name="data"))
    objective.add(th.AutoDiffCostFunction([ab], prior_err, 2, aux_vars=[t], cost_weight=th.ScaleCostWeight(torch.tensor(0.3, dtype=dtype)), name="prior"))
    return objective


def backward_inputs(torch, B=3, seed=21):
    g = torch.Generator().manual_seed(seed)
    N = 12
    x = torch.linspace(-1, 1, N, dtype=torch.float64).repeat(B, 1) + 0.05 * torch.randn(B, N, generator=g, dtype=torch.float64)
    a = 1.0 + 0.3 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    b = 0.5 + 0.3 * torch.rand(B, 1, generator=g, dtype=torch.float64)
    y = a * torch.exp(b * x) + 0.2 + 0.02 * torch.randn(B, N, generator=g, dtype=torch.float64)
    t = torch.cat([a, b], 1) + 0.1 * torch.randn(B, 2, generator=g, dtype=torch.float64)
    ab0 = torch.cat([a, b], 1) + 0.2 * torch.randn(B, 2, generator=g, dtype=torch.float64)
    c0 = torch.zeros(B, 1, dtype=torch.float64)
    w = torch.randn(B, 3, generator=g, dtype=torch.float64)
    return dict(x=x, y=y, t=t, ab=ab0, c=c0), w


BACKWARD_CASES = {
    "implicit_gn": ("gn", 12, dict(backward_mode="implicit")),
    "unroll_gn": ("gn", 5, dict(backward_mode="unroll")),
    # LM with adaptive damping on the tape: every taped iteration must be a DECISIVE one (clear error reduction), otherwise the
    # accept/reject of a converged step is rounding noise and so is which steps carry gradient -> strong damping, few iterations
    "truncated_lm": ("lm", 5, dict(backward_mode="truncated", backward_num_iterations=2, damping=2.0, adaptive_damping=True, ellipsoidal_damping=True)),
    "unroll_lm": ("lm", 4, dict(backward_mode="unroll", damping=2.0, adaptive_damping=True, ellipsoidal_damping=True)),
    "implicit_lm": ("lm", 12, dict(backward_mode="implicit", damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)),
}


def autodiff_lie_problem(th, torch, inputs):
    """Point-set alignment with AutoDiffCostFunctions over Lie-group variables (err_fn works on the raw storage tensors):
    SE3 pose T3: e = R p + t - q (K=6 points), SE2 pose T2: e = R(cos,sin) p2 + t2 - q2 (K=5 points).  Shared by generator and tests."""
    dtype = torch.float64
    T3 = th.SE3(tensor=inputs["T3"].clone(), name="T3")
    T2 = th.SE2(tensor=inputs["T2"].clone(), name="T2")
    p, q = th.Variable(inputs["p"].clone(), name="p"), th.Variable(inputs["q"].clone(), name="q")
    p2, q2 = th.Variable(inputs["p2"].clone(), name="p2"), th.Variable(inputs["q2"].clone(), name="q2")

    def err3(optim_vars, aux_vars):
        g = optim_vars[0].tensor
        pp, qq = aux_vars[0].tensor, aux_vars[1].tensor          # [B,K,3]
        r = torch.einsum("bij,bkj->bki", g[:, :, :3], pp) + g[:, None, :, 3] - qq
        return r.reshape(r.shape[0], -1)

    def err2(optim_vars, aux_vars):
        g = optim_vars[0].tensor                                  # [B,4] = x, y, cos, sin
        pp, qq = aux_vars[0].tensor, aux_vars[1].tensor          # [B,K,2]
        c, s_ = g[:, None, 2], g[:, None, 3]
        rx = c * pp[..., 0] - s_ * pp[..., 1] + g[:, None, 0] - qq[..., 0]
        ry = s_ * pp[..., 0] + c * pp[..., 1] + g[:, None, 1] - qq[..., 1]
        return torch.stack((rx, ry), dim=-1).reshape(g.shape[0], -1)

    objective = th.Objective(dtype=dtype)
    objective.add(th.AutoDiffCostFunction([T3], err3, 18, aux_vars=[p, q], cost_weight=th.ScaleCostWeight(torch.tensor(1.5, dtype=dtype)), name="align3"))
    objective.add(th.AutoDiffCostFunction([T2], err2, 10, aux_vars=[p2, q2], cost_weight=th.ScaleCostWeight(torch.tensor(0.7, dtype=dtype)), name="align2"))
    return objective


def make_autodiff_lie(th):
    """AutoDiffCostFunction over SE3 / SE2 variables: vmap(jacrev) Euclidean Jacobians + project (cost_function.py:343-393)."""
    import torch
    g = torch.Generator().manual_seed(33)
    B = 3
    d = torch.float64
    Tgt3 = th.SE3.exp_map(0.6 * torch.randn(B, 6, generator=g, dtype=d))
    Tgt2 = th.SE2.exp_map(0.8 * torch.randn(B, 3, generator=g, dtype=d))
    p = torch.randn(B, 6, 3, generator=g, dtype=d)
    q = torch.einsum("bij,bkj->bki", Tgt3.tensor[:, :, :3], p) + Tgt3.tensor[:, None, :, 3] + 0.01 * torch.randn(B, 6, 3, generator=g, dtype=d)
    p2 = torch.randn(B, 5, 2, generator=g, dtype=d)
    c, s_ = Tgt2.tensor[:, None, 2], Tgt2.tensor[:, None, 3]
    q2 = torch.stack((c * p2[..., 0] - s_ * p2[..., 1] + Tgt2.tensor[:, None, 0], s_ * p2[..., 0] + c * p2[..., 1] + Tgt2.tensor[:, None, 1]), -1)
    q2 = q2 + 0.01 * torch.randn(B, 5, 2, generator=g, dtype=d)
    T3_0 = th.SE3.exp_map(0.3 * torch.randn(B, 6, generator=g, dtype=d)).compose(Tgt3)
    T2_0 = th.SE2.exp_map(0.3 * torch.randn(B, 3, generator=g, dtype=d)).compose(Tgt2)
    inputs = dict(T3=T3_0.tensor, T2=T2_0.tensor, p=p, q=q, p2=p2, q2=q2)
    objective = autodiff_lie_problem(th, torch, inputs)
    iters = 6
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=iters, step_size=1.0,
                                abs_err_tolerance=0, rel_err_tolerance=0)
    lin = th.SparseLinearization(objective)
    objective.update()
    lin.linearize()
    out = {k: v.numpy() for k, v in inputs.items()}
    out["A_val0"], out["b0"] = lin.A_val.numpy().copy(), lin.b.numpy().copy()
    deltas, errs = [], []

    def cb(optimizer, info, delta, it):
        deltas.append(delta.detach().numpy().copy()); errs.append(info.last_err.detach().numpy().copy())

    with torch.no_grad():
        info = opt.optimize(end_iter_callback=cb, damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
    out["trace_delta"], out["trace_err"] = np.stack(deltas, 0), np.stack(errs, 0)
    out["final_T3"], out["final_T2"] = objective.get_optim_var("T3").tensor.numpy(), objective.get_optim_var("T2").tensor.numpy()
    np.savez_compressed(os.path.join(HERE, "autodiff_lie.npz"), **out)
    print("autodiff_lie err trace", np.stack(errs, 0)[:, 0])


BACKWARD_LIE_CASES = {
    "implicit_gn": ("gn", 10, dict(backward_mode="implicit")),
    "unroll_gn": ("gn", 4, dict(backward_mode="unroll")),
    "unroll_lm": ("lm", 3, dict(backward_mode="unroll", damping=1.0, adaptive_damping=True, ellipsoidal_damping=True)),
}


def make_backward_lie(th):
    """Backward modes with Lie-group variables: the SE3 + SE2 point-alignment objective of make_autodiff_lie, gradients of an outer
    loss on the optimised poses w.r.t. the observed points q, q2 (the tactile-style use: learn the measurement model end to end)."""
    import torch
    base = dict(np.load(os.path.join(HERE, "autodiff_lie.npz")))
    out = {}
    for name, (method, iters, kw) in BACKWARD_LIE_CASES.items():
        inputs = {k: torch.from_numpy(base[k]) for k in ("T3", "T2", "p", "q", "p2", "q2")}
        objective = autodiff_lie_problem(th, torch, inputs)
        cls = th.GaussNewton if method == "gn" else th.LevenbergMarquardt
        opt = cls(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0)
        layer = th.TheseusLayer(opt)
        leaves = {k: inputs[k].clone().requires_grad_(True) for k in ("q", "q2", "p")}
        sol, info = layer.forward({**leaves, "T3": inputs["T3"].clone(), "T2": inputs["T2"].clone()}, optimizer_kwargs=dict(kw, track_err_history=True))
        eh = info.err_history.numpy()
        print("   rel. error reduction per iteration:", np.array2string(((eh[:, :-1] - eh[:, 1:]) / eh[:, :-1]).min(axis=0), precision=2))
        wgt3 = torch.linspace(0.5, 1.5, 12, dtype=torch.float64).view(1, 3, 4)
        wgt2 = torch.linspace(-1.0, 1.0, 4, dtype=torch.float64).view(1, 4)
        loss = (sol["T3"] * wgt3).sum() + (sol["T2"] * wgt2).sum()
        loss.backward()
        out[name + "_T3"], out[name + "_T2"] = sol["T3"].detach().numpy(), sol["T2"].detach().numpy()
        for k in leaves:
            out[name + "_grad_" + k] = leaves[k].grad.numpy()
        print("backward_lie", name, "loss", loss.item(), "|grad_q|", float(leaves["q"].grad.abs().max()), "|grad_q2|", float(leaves["q2"].grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "backward_lie_kat.npz"), **out)


G2O_3D = """VERTEX_SE3:QUAT 0 0 0 0 0 0 0 1
VERTEX_SE3:QUAT 1 1.02 0.03 -0.01 0.01 0.02 0.10 0.994
VERTEX_SE3:QUAT 2 2.05 0.10 0.02 -0.02 0.03 0.25 0.967
VERTEX_SE3:QUAT 3 2.90 0.95 0.05 0.05 -0.01 0.60 0.798
EDGE_SE3:QUAT 0 1 1.0 0.0 0.0 0.0 0.0 0.0998 0.995 100 0 0 0 0 0 100 0 0 0 0 100 0 0 0 400 0 0 400 0 400
EDGE_SE3:QUAT 1 2 1.0 0.1 0.0 0.0 0.0 0.149 0.9888 90 1 0 0 0 0 110 0 2 0 0 95 0 0 0 380 0 0 420 3 410
EDGE_SE3:QUAT 2 3 1.0 0.5 0.0 0.02 0.0 0.38 0.9248 100 0 0 0 0 0 100 0 0 0 0 100 0 0 0 400 0 0 400 0 400
EDGE_SE3:QUAT 0 3 2.9 1.0 0.0 0.0 0.0 0.64 0.7684 50 0 0 0 0 0 50 0 0 0 0 50 0 0 0 200 0 0 200 0 200
"""


def make_io(th):
    """On-disk formats: the g2o reader (pose_graph/dataset.py:35-104) and the BAL loader (bundle_adjustment/data.py:166-207) of the
    reference on small files committed next to this script."""
    import torch
    import theseus.utils.examples as theg
    from theseus.utils.examples.pose_graph.dataset import read_3D_g2o_file
    g2o = os.path.join(HERE, "io_small.g2o")
    with open(g2o, "w") as f:
        f.write(G2O_3D)
    n, verts, edges = read_3D_g2o_file(g2o, dtype=torch.float64)
    out = dict(g2o_n=np.array(n), g2o_verts=np.concatenate([v.tensor.numpy() for v in verts], 0),
               g2o_edge_ij=np.array([[e.i, e.j] for e in edges]), g2o_edge_pose=np.concatenate([e.relative_pose.tensor.numpy() for e in edges], 0),
               g2o_edge_w=np.concatenate([e.weight.diagonal.tensor.numpy() for e in edges], 0))
    torch.manual_seed(5); np.random.seed(5)
    import random
    random.seed(5)
    ba = theg.BundleAdjustmentDataset.generate_synthetic(num_cameras=3, num_points=8, average_track_length=2, track_locality=0.3)
    bal = os.path.join(HERE, "io_small_bal.txt")
    ba.save_to_file(bal)
    cams, pts, obs = theg.BundleAdjustmentDataset.load_bal_dataset(bal)
    out.update(bal_cam_pose=np.concatenate([c.pose.tensor.numpy() for c in cams], 0),
               bal_cam_f=np.concatenate([c.focal_length.tensor.numpy() for c in cams], 0),
               bal_cam_k1=np.concatenate([c.calib_k1.tensor.numpy() for c in cams], 0),
               bal_cam_k2=np.concatenate([c.calib_k2.tensor.numpy() for c in cams], 0),
               bal_pts=np.concatenate([p.tensor.numpy() for p in pts], 0),
               bal_obs=np.array([[o.camera_index, o.point_index] for o in obs]),
               bal_feat=np.concatenate([o.image_feature_point.tensor.numpy() for o in obs], 0))
    np.savez_compressed(os.path.join(HERE, "io_kat.npz"), **out)
    print("io_kat: g2o", n, "vertices", len(edges), "edges; bal", len(cams), "cameras", len(pts), "points", len(obs), "observations")


BACKWARD_PGO_CASES = {
    # (fixture, method, iters, optimizer kwargs)
    "plain_implicit_gn": ("pgo_small_lm_hard", "gn", 8, dict(backward_mode="implicit")),
    "plain_unroll_lm": ("pgo_small_lm_hard", "lm", 3, dict(backward_mode="unroll", damping=1.0, adaptive_damping=True, ellipsoidal_damping=True)),
    "welsch_implicit_lm": ("pgo_small_welsch", "lm", 8, dict(backward_mode="implicit", damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)),
    "welsch_unroll_lm": ("pgo_small_welsch", "lm", 3, dict(backward_mode="unroll", damping=1.0, adaptive_damping=True, ellipsoidal_damping=True)),
}


def backward_pgo_run(th, torch, case, device="cpu", solver_kwargs=None):
    """Shared by the generator (th = reference) and tests/test_gpu_backward.py (th = theseus_b200): pose graph of a committed fixture
    with Between (+ Welsch) costs, leaves = every edge's DiagonalCostWeight, the prior's scale and the robust log-radius; outer loss =
    fixed random weights dotted with the optimised poses.  Returns (poses [N,B,3,4], grads dict)."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    from helpers import load, pgo_objective
    fixture, method, iters, kw = BACKWARD_PGO_CASES[case]
    g = load(fixture)
    objective, poses = pgo_objective(th, g, device=device)
    # gauge: the fixture's prior has weight 1e-3 (cond(AtA) ~ 4e8), which makes an UN-damped GN step -- what the implicit mode ends
    # with -- reproducible only to ~1e-6 between two correct implementations; a firm prior on pose 0 makes the comparison sharp
    gauge_target = th.SE3(tensor=poses[0].tensor.detach().clone(), name="GAUGE_TARGET")
    objective.add(th.Difference(poses[0], gauge_target, th.ScaleCostWeight(torch.tensor(10.0, dtype=torch.float64).to(poses[0].tensor.device)), name="gauge"))
    leaves = {}
    for name, cf in objective.cost_functions.items():
        inner = getattr(cf, "cost_function", cf)
        wt = inner.weight
        leaves["w_" + inner.name] = (wt.diagonal if hasattr(wt, "diagonal") else wt.scale).tensor
        if hasattr(cf, "log_loss_radius"):
            leaves["log_loss_radius"] = cf.log_loss_radius.tensor
    for t in leaves.values():
        t.requires_grad_(True)
    cls = th.GaussNewton if method == "gn" else th.LevenbergMarquardt
    opt = cls(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **(solver_kwargs or dict(linear_solver_cls=th.CholeskyDenseSolver)))
    layer = th.TheseusLayer(opt)
    sol, info = layer.forward({p.name: p.tensor.clone() for p in poses}, optimizer_kwargs=dict(kw))
    gen = torch.Generator().manual_seed(77)
    P = torch.stack([sol[p.name] for p in poses], 0)
    Wfix = torch.randn(P.shape, generator=gen, dtype=torch.float64).to(P.device)
    (P * Wfix).sum().backward()
    return P.detach(), {k: v.grad.detach() for k, v in leaves.items()}


def make_backward_pgo(th):
    """Backward modes through the reference's own Between / RobustCostFunction (torchlie analytic Jacobians, autograd through them)."""
    import torch
    out = {}
    for case in BACKWARD_PGO_CASES:
        P, grads = backward_pgo_run(th, torch, case)
        out[case + "_poses"] = P.numpy()
        for k, v in grads.items():
            out[case + "_grad_" + k] = v.numpy()
        print("backward_pgo", case, "max |grad w|", max(float(v.abs().max()) for k, v in grads.items() if k.startswith("w_")),
              "grad log_radius", grads.get("log_loss_radius"))
    np.savez_compressed(os.path.join(HERE, "backward_pgo_kat.npz"), **out)


def make_moving_frame(th):
    """MovingFrameBetween (embodied/measurements/moving_frame_between.py): weighted error + analytic Jacobians for SE2 and SE3."""
    import torch
    torch.manual_seed(3)
    d = torch.float64
    out = {}
    for name, cls in (("se2", th.SE2), ("se3", th.SE3)):
        B = 6
        vs = [cls.rand(B, dtype=d) for _ in range(5)]
        w = torch.rand(1, vs[0].dof(), dtype=d) + 0.5
        cf = th.eb.MovingFrameBetween(vs[0], vs[1], vs[2], vs[3], vs[4], th.DiagonalCostWeight(w))
        jacs, e = cf.weighted_jacobians_error()
        out.update({f"{name}_in": np.stack([v.tensor.numpy() for v in vs], 0), f"{name}_w": w.numpy(), f"{name}_e": e.numpy(),
                    f"{name}_J": np.stack([j.numpy() for j in jacs], 0)})
    # quasi-static pushing (embodied/motionmodel/quasi_static_pushing_planar.py), SE2
    B = 6
    vs = [th.SE2.rand(B, dtype=d) for _ in range(4)]
    c2 = torch.rand(B, 1, dtype=d) + 0.1
    w = torch.rand(1, 3, dtype=d) + 0.5
    cf = th.eb.QuasiStaticPushingPlanar(vs[0], vs[1], vs[2], vs[3], th.Variable(c2), th.DiagonalCostWeight(w))
    jacs, e = cf.weighted_jacobians_error()
    out.update(qsp_in=np.stack([v.tensor.numpy() for v in vs], 0), qsp_c2=c2.numpy(), qsp_w=w.numpy(), qsp_e=e.numpy(),
               qsp_J=np.stack([j.numpy() for j in jacs], 0))
    # effector-object contact (embodied/collision/eff_obj_contact.py): random smooth SDF grid, effector positions inside and outside
    B = 8
    rows, cols = 12, 15
    yy, xx = torch.meshgrid(torch.arange(rows, dtype=d), torch.arange(cols, dtype=d), indexing="ij")
    sdf = (((xx - 7.0) * 0.1) ** 2 + ((yy - 5.5) * 0.1) ** 2).sqrt().unsqueeze(0) - 0.35 + 0.01 * torch.randn(1, rows, cols, dtype=d)
    origin = torch.tensor([[-0.7, -0.55]], dtype=d)
    obj = th.SE2.rand(B, dtype=d)
    eff_xy_obj = torch.rand(B, 2, dtype=d) * torch.tensor([1.3, 1.0], dtype=d) + origin   # inside the grid (object frame) ...
    eff_xy_obj[-1] = torch.tensor([5.0, 5.0], dtype=d)                                       # ... except the last one
    eff = th.SE2(x_y_theta=torch.cat([obj.transform_from(eff_xy_obj).tensor, torch.rand(B, 1, dtype=d)], 1))
    w = torch.rand(1, 1, dtype=d) + 0.5
    cf = th.eb.EffectorObjectContactPlanar(obj, eff, origin, sdf, 0.1, torch.tensor(0.05, dtype=d), th.ScaleCostWeight(w))
    jacs, e = cf.weighted_jacobians_error()
    out.update(eoc_obj=obj.tensor.numpy(), eoc_eff=eff.tensor.numpy(), eoc_sdf=sdf.numpy(), eoc_origin=origin.numpy(), eoc_w=w.numpy(),
               eoc_e=e.numpy(), eoc_J=np.stack([j.numpy() for j in jacs], 0))
    np.savez_compressed(os.path.join(HERE, "moving_frame_kat.npz"), **out)
    print("moving_frame_kat", {k: v.shape for k, v in out.items()})


def make_robust(th):
    """RobustCostFunction / GNCRobustCostFunction (core/robust_cost_function.py, core/robust_loss.py) around Between(SE3) and a Vector
    Difference: weighted Jacobians + error (linearisation rescale) and weighted_error (the loss value spread over `dim` entries), for every
    loss of the reference, with and without flatten_dims.  Squared norms straddle the radius (both branches of Huber / Hinge)."""
    import torch
    torch.manual_seed(17)
    d = torch.float64
    B = 8
    X0, X1 = th.SE3.rand(B, dtype=d), th.SE3.rand(B, dtype=d)
    # measurements close to / far from the current relative pose: small and large residuals
    scale = torch.tensor([1e-3, 0.05, 0.2, 0.5, 1.0, 2.0, 0.3, 0.7], dtype=d).view(B, 1)
    Z = th.SE3(tensor=X0.between(X1).compose(th.SE3.exp_map(scale * (2 * torch.rand(B, 6, dtype=d) - 1))).tensor)
    w = torch.rand(1, 6, dtype=d) + 0.5
    log_radius = torch.tensor([[-0.7]], dtype=d)
    mu = torch.tensor([[3.0]], dtype=d)
    V, Vt = th.Vector(tensor=torch.randn(B, 4, dtype=d) * scale), th.Vector(tensor=torch.zeros(B, 4, dtype=d))
    wv = torch.tensor(1.3, dtype=d)
    out = dict(X0=X0.tensor.numpy(), X1=X1.tensor.numpy(), Z=Z.tensor.numpy(), w=w.numpy(), log_radius=log_radius.numpy(), mu=mu.numpy(),
               V=V.tensor.numpy(), Vt=Vt.tensor.numpy(), wv=np.array(float(wv)))
    losses = dict(welsch=th.WelschLoss, huber=th.HuberLoss, hinge=th.HingeLoss, geman=th.GemanMcClureLoss)
    for name, cls in losses.items():
        for flat in (False, True):
            for cname, inner in (("between", lambda: th.Between(X0, X1, Z, th.DiagonalCostWeight(w))),
                                 ("vecdiff", lambda: th.Difference(V, Vt, th.ScaleCostWeight(wv)))):
                lr = th.Vector(tensor=log_radius.clone(), name=f"lr_{name}_{flat}_{cname}")
                if name == "geman":
                    cf = th.GNCRobustCostFunction(inner(), cls, lr, th.Vector(tensor=mu.clone(), name=f"mu_{flat}_{cname}"), flatten_dims=flat)
                else:
                    cf = th.RobustCostFunction(inner(), cls, lr, flatten_dims=flat)
                jacs, e = cf.weighted_jacobians_error()
                key = f"{name}_{'flat' if flat else 'full'}_{cname}"
                out[key + "_J"] = np.stack([j.numpy() for j in jacs], 0)
                out[key + "_e"] = e.numpy()
                out[key + "_val"] = cf.weighted_error().numpy()
    x = (th.Between(X0, X1, Z, th.DiagonalCostWeight(w)).weighted_error() ** 2).sum(1)
    print("robust_kat: squared norms", x.numpy().round(4), "radius", float(log_radius.exp()))
    np.savez_compressed(os.path.join(HERE, "robust_kat.npz"), **out)
    print("robust_kat", len(out), "arrays")


def so2_problem(th, torch, thetas0, meas, edges, w_edge, w_prior, device="cpu"):
    """Rotation averaging on SO2, shared by generator and tests: N angles, Between edges (i, j) with measured relative rotations, a
    prior on the first.  thetas0 [N,B,1], meas [E,B,1]."""
    d = thetas0.dtype
    vs = [th.SO2(theta=thetas0[i].to(device), name=f"R{i}") for i in range(thetas0.shape[0])]
    objective = th.Objective(dtype=d)
    for e, (i, j) in enumerate(edges):
        objective.add(th.Between(vs[i], vs[j], th.SO2(theta=meas[e].to(device), name=f"Z{e}"),
                                 th.ScaleCostWeight(torch.tensor(float(w_edge[e]), dtype=d, device=device)), name=f"between_{e}"))
    objective.add(th.Difference(vs[0], th.SO2(theta=thetas0[0].to(device), name="R0_prior"),
                                th.ScaleCostWeight(torch.tensor(float(w_prior), dtype=d, device=device)), name="prior"))
    if str(device) != "cpu":
        objective.to(device)
    return objective, vs


def make_so2(th):
    """SO2 (geometry/so2.py): exp / log / compose / inverse / adjoint / project known answers, Between and Difference weighted
    Jacobians, and an LM trace of a small rotation-averaging problem (dense Cholesky)."""
    import torch
    torch.manual_seed(23)
    d = torch.float64
    B = 12
    theta = torch.cat([torch.tensor([0.0, 1e-9, -1e-9, np.pi - 1e-9, -np.pi + 1e-9, np.pi / 2], dtype=d), 6 * torch.rand(B - 6, dtype=d) - 3]).view(B, 1)
    X = th.SO2(theta=theta)
    Y = th.SO2.rand(B, dtype=d)
    out = dict(theta=theta.numpy(), X=X.tensor.numpy(), Y=Y.tensor.numpy(), log_X=X.log_map().numpy(), compose=X.compose(Y).tensor.numpy(),
               inverse=X.inverse().tensor.numpy(), adjoint=X.adjoint().numpy(), local=X.local(Y).numpy(),
               retract=X.retract(theta.flip(0) * 0.3).tensor.numpy(), retract_delta=(theta.flip(0) * 0.3).numpy())
    G = torch.randn(B, 5, 2, dtype=d)
    out["proj_in"], out["proj_out"] = G.numpy(), X.project(G, is_sparse=True).numpy()
    Z = th.SO2.rand(B, dtype=d)
    cf = th.Between(X, Y, Z, th.ScaleCostWeight(torch.tensor(0.7, dtype=d)))
    (J0, J1), e = cf.weighted_jacobians_error()
    cl = th.Difference(X, Z, th.ScaleCostWeight(torch.tensor(1.3, dtype=d)))
    (Jl,), el = cl.weighted_jacobians_error()
    out.update(Z=Z.tensor.numpy(), between_J0=J0.numpy(), between_J1=J1.numpy(), between_e=e.numpy(), local_J=Jl.numpy(), local_e=el.numpy())
    # LM trace
    N, Bp = 7, 4
    gt = 2 * np.pi * torch.rand(N, Bp, 1, dtype=d) - np.pi
    edges = [(i, (i + 1) % N) for i in range(N)] + [(0, 3), (2, 5), (1, 4)]
    meas = torch.stack([(gt[j] - gt[i]) + 0.05 * torch.randn(Bp, 1, dtype=d) for (i, j) in edges], 0)
    thetas0 = gt + 0.4 * torch.randn(N, Bp, 1, dtype=d)
    w_edge = (torch.rand(len(edges), dtype=d) + 0.5).numpy()
    objective, vs = so2_problem(th, torch, thetas0, meas, edges, w_edge, 0.1)
    iters = 6
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, vectorize=True, max_iterations=iters, step_size=1.0,
                                abs_err_tolerance=0, rel_err_tolerance=0)
    assert [v.name for v in opt.linear_solver.linearization.ordering] == [v.name for v in vs]
    tr = dict(err=[], delta=[])

    def cb(optimizer, info, delta, it):
        tr["err"].append(info.last_err.numpy().copy()); tr["delta"].append(delta.detach().numpy().copy())
    with torch.no_grad():
        objective.update()
        err0 = objective.error_metric().numpy().copy()
        opt.optimize(end_iter_callback=cb, damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
    out.update(lm_thetas0=thetas0.numpy(), lm_meas=meas.numpy(), lm_edges=np.array(edges, dtype=np.int64), lm_w_edge=w_edge, lm_w_prior=np.array(0.1),
               lm_err0=err0, lm_trace_err=np.stack(tr["err"], 0), lm_trace_delta=np.stack(tr["delta"], 0),
               lm_final=np.stack([v.tensor.numpy() for v in vs], 0))
    np.savez_compressed(os.path.join(HERE, "so2_kat.npz"), **out)
    print("so2_kat", len(out), "arrays; LM err", err0, "->", tr["err"][-1])


def tactile_problem(th, torch, inputs, device="cpu"):
    """A planar-pushing (tactile-style, config C4's cost set) objective shared by generator and tests: T object poses o_t and T
    effector poses e_t (SE2); per step: EffectorObjectContactPlanar(o_t, e_t), Difference(e_t, measured effector pose);
    between steps: QuasiStaticPushingPlanar(o_t, o_t+1, e_t, e_t+1), MovingFrameBetween(o_t, o_t+1, e_t, e_t+1, measurement);
    a prior on o_0.  inputs: dict of tensors (see make_tactile)."""
    d = torch.float64
    T = inputs["obj"].shape[0]
    objs = [th.SE2(tensor=inputs["obj"][t].clone().to(device), name=f"obj_{t}") for t in range(T)]
    effs = [th.SE2(tensor=inputs["eff"][t].clone().to(device), name=f"eff_{t}") for t in range(T)]
    c2 = th.Variable(inputs["c_square"].clone().to(device), name="c_square")
    radius = th.Variable(inputs["eff_radius"].clone().to(device), name="eff_radius")
    sdf = th.Variable(inputs["sdf"].clone().to(device), name="sdf_data")
    origin = th.Point2(tensor=inputs["sdf_origin"].clone().to(device), name="sdf_origin")
    cell = th.Variable(inputs["sdf_cell"].clone().to(device), name="sdf_cell")
    w = lambda v, n: th.ScaleCostWeight(th.Variable(torch.tensor([[v]], dtype=d).to(device), name=n))
    w_qsp, w_eoc, w_mfb, w_eff, w_prior = w(2.0, "w_qsp"), w(3.0, "w_eoc"), w(1.5, "w_mfb"), w(10.0, "w_eff"), w(10.0, "w_prior")
    objective = th.Objective(dtype=d)
    for t in range(T):
        objective.add(th.eb.EffectorObjectContactPlanar(objs[t], effs[t], origin, sdf, cell, radius, w_eoc, name=f"eoc_{t}"))
        objective.add(th.Difference(effs[t], th.SE2(tensor=inputs["eff_meas"][t].clone().to(device), name=f"eff_meas_{t}"), w_eff, name=f"effprior_{t}"))
    for t in range(T - 1):
        objective.add(th.eb.QuasiStaticPushingPlanar(objs[t], objs[t + 1], effs[t], effs[t + 1], c2, w_qsp, name=f"qsp_{t}"))
        objective.add(th.eb.MovingFrameBetween(objs[t], objs[t + 1], effs[t], effs[t + 1],
                                               th.SE2(tensor=inputs["mfb_meas"][t].clone().to(device), name=f"mfb_meas_{t}"), w_mfb, name=f"mfb_{t}"))
    objective.add(th.Difference(objs[0], th.SE2(tensor=inputs["obj"][0].clone().to(device), name="obj0_prior"), w_prior, name="objprior"))
    if str(device) != "cpu":
        objective.to(device)
    return objective, objs, effs, dict(c_square=c2, eff_radius=radius, w_qsp=w_qsp.scale, w_eoc=w_eoc.scale, w_mfb=w_mfb.scale)


def make_tactile(th, T=5, B=4, name="tactile_kat", seed=17, keep_deltas=True):
    """LM trace + implicit-mode gradients of the planar-pushing objective above, by the reference (dense solver, fp64).
    (T=25, B=512, name="tactile_c4_kat": BASELINE.json's config C4 at its real batch size.)"""
    import torch
    torch.manual_seed(seed)
    d = torch.float64
    rows, cols = 16, 16
    yy, xx = torch.meshgrid(torch.arange(rows, dtype=d), torch.arange(cols, dtype=d), indexing="ij")
    sdf = (((xx - 7.5) * 0.05) ** 2 + ((yy - 7.5) * 0.05) ** 2).sqrt().unsqueeze(0) - 0.2    # disc of radius 0.2 centred in the object frame
    obj_gt = [th.SE2(x_y_theta=torch.cat([0.05 * t + 0.02 * torch.randn(B, 2, dtype=d), 0.1 * t + 0.05 * torch.randn(B, 1, dtype=d)], 1)) for t in range(T)]
    ang = [torch.rand(B, 1, dtype=d) * 6.28 for _ in range(T)]
    eff_gt = [th.SE2(x_y_theta=torch.cat([obj_gt[t].transform_from(0.25 * torch.cat([a.cos(), a.sin()], 1)).tensor, torch.zeros(B, 1, dtype=d)], 1))
              for t, a in enumerate(ang)]
    noise = lambda g, s: g.compose(th.SE2.exp_map(s * torch.randn(B, 3, dtype=d)))
    inputs = dict(obj=torch.stack([noise(g, 0.05).tensor for g in obj_gt], 0), eff=torch.stack([noise(g, 0.02).tensor for g in eff_gt], 0),
                  eff_meas=torch.stack([g.tensor for g in eff_gt], 0),
                  mfb_meas=torch.stack([noise(obj_gt[t].between(eff_gt[t]).between(obj_gt[t + 1].between(eff_gt[t + 1])), 0.01).tensor for t in range(T - 1)], 0),
                  c_square=torch.tensor([[0.3]], dtype=d), eff_radius=torch.tensor([[0.05]], dtype=d), sdf=sdf,
                  sdf_origin=torch.tensor([[-0.375, -0.375]], dtype=d), sdf_cell=torch.tensor([[0.05]], dtype=d))
    out = {k: v.numpy() for k, v in inputs.items()}
    lm = dict(damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
    # forward trace
    objective, objs, effs, leaves = tactile_problem(th, torch, inputs)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=8, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0)
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.detach().numpy().copy()); deltas.append(delta.detach().numpy().copy())
    with torch.no_grad():
        objective.update()
        out["err0"] = objective.error_metric().numpy()
        info = opt.optimize(end_iter_callback=cb, **lm)
    out["trace_err"] = np.stack(errs, 0)
    out["trace_delta"] = np.stack(deltas, 0) if keep_deltas else np.stack(deltas[:1], 0)
    out["final_obj"] = np.stack([o.tensor.numpy() for o in objs], 0)
    # implicit-mode gradients w.r.t. c_square, eff_radius and three cost weights
    objective, objs, effs, leaves = tactile_problem(th, torch, inputs)
    for v in leaves.values():
        v.tensor.requires_grad_(True)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=8, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0)
    sol, info = th.TheseusLayer(opt).forward({v.name: v.tensor.clone() for v in objs + effs}, optimizer_kwargs=dict(lm, backward_mode="implicit"))
    gen = torch.Generator().manual_seed(5)
    P = torch.stack([sol[o.name] for o in objs], 0)
    (P * torch.randn(P.shape, generator=gen, dtype=d)).sum().backward()
    for k, v in leaves.items():
        out["grad_" + k] = v.tensor.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "err", out["err0"][:4], "->", out["trace_err"][-1][:4], "grads", {k: float(np.abs(out["grad_" + k]).max()) for k in leaves})



def geometry_api_values(L, torch, I, tape, device="cpu", group_ops=True):
    """Every public method of the geometry classes beyond the fused-kernel ones, called the same way on the reference (`L` = theseus) and
    on theseus_b200: group operations (tape=True: operands require grad -> the differentiable route), actions on points with their
    Jacobians, accessors, conversions, Vector arithmetic, projections.  `I`: dict of input tensors (make_geom_api).  Shared by the
    generator and tests/test_geometry_api.py."""
    out = {}
    dev = lambda t: t.to(device)
    mk = lambda t: dev(t).clone().requires_grad_(True) if tape else dev(t).clone()
    T = lambda x: x.tensor if hasattr(x, "tensor") else x
    for name, cls, X, Y, tv in (("se3", L.SE3, I["X3"], I["Y3"], I["t6"]), ("so3", L.SO3, I["R3"], I["S3"], I["t3"]),
                                ("se2", L.SE2, I["X2"], I["Y2"], I["t3"]), ("so2", L.SO2, I["R2"], I["S2"], I["t1"])):
        a, b = cls(tensor=mk(X)), cls(tensor=mk(Y))
        if group_ops:
            out[name + "_compose"] = T(a.compose(b)); out[name + "_inverse"] = T(a.inverse()); out[name + "_log"] = a.log_map()
            out[name + "_exp"] = T(cls.exp_map(mk(tv))); out[name + "_adjoint"] = a.adjoint(); out[name + "_between"] = T(a.between(b))
            out[name + "_local"] = a.local(b); out[name + "_retract"] = T(a.retract(mk(tv)))
            if not tape:     # the optional Jacobian outputs (lie_group.py:125-195), plain operands
                J = []; a.compose(b, jacobians=J); out[name + "_compose_J0"], out[name + "_compose_J1"] = J
                J = []; a.inverse(jacobian=J); out[name + "_inverse_J"] = J[0]
                J = []; a.between(b, jacobians=J); out[name + "_between_J0"], out[name + "_between_J1"] = J
                J = []; a.local(b, jacobians=J); out[name + "_local_J0"], out[name + "_local_J1"] = J
                J = []; a.log_map(jacobians=J); out[name + "_log_J"] = J[0]
                J = []; cls.exp_map(dev(tv), jacobians=J); out[name + "_exp_J"] = J[0]
        out[name + "_to_matrix"] = a.to_matrix(); out[name + "_hat"] = cls.hat(dev(tv)); out[name + "_vee"] = cls.vee(cls.hat(dev(tv)))
    a3, a2, r3, r2 = L.SE3(tensor=mk(I["X3"])), L.SE2(tensor=mk(I["X2"])), L.SO3(tensor=mk(I["R3"])), L.SO2(tensor=mk(I["R2"]))
    p3, p2 = dev(I["p3"]), dev(I["p2"])
    for nm, obj, meth, pt in (("se3_tf", a3, "transform_from", p3), ("se3_tt", a3, "transform_to", p3), ("so3_rot", r3, "rotate", p3),
                              ("so3_unrot", r3, "unrotate", p3), ("se2_tf", a2, "transform_from", p2), ("se2_tt", a2, "transform_to", p2),
                              ("so2_rot", r2, "rotate", p2), ("so2_unrot", r2, "unrotate", p2)):
        J = []
        out[nm] = T(getattr(obj, meth)(pt, jacobians=J)); out[nm + "_Jg"], out[nm + "_Jp"] = J
        out[nm + "_pt"] = T(getattr(obj, meth)((L.Point3 if pt.shape[1] == 3 else L.Point2)(tensor=pt)))
    out["se3_rot"], out["se3_trans"] = T(a3.rotation()), T(a3.translation())
    out["se2_rot"], out["se2_trans"], out["se2_theta"], out["se2_xy"] = T(a2.rotation), T(a2.translation), a2.theta(), T(a2.xy())
    J = []; a2.theta(jacobians=J); out["se2_theta_J"] = J[0]
    J = []; a2.xy(jacobians=J); out["se2_xy_J"] = J[0]
    out["so2_theta"] = r2.theta(); out["so2_cs0"], out["so2_cs1"] = r2.to_cos_sin()
    out["so3_quat"] = r3.to_quaternion(); out["so3_from_quat"] = T(L.SO3.unit_quaternion_to_SO3(dev(I["q"])))
    out["se3_xyzq"] = a3.to_x_y_z_quaternion(); out["se3_from_xyzq"] = T(L.SE3.x_y_z_unit_quaternion_to_SE3(torch.cat((p3, dev(I["q"])), 1)))
    v, w = L.Vector(tensor=p3.clone()), L.Vector(tensor=dev(I["t3"]).clone())
    out["v_add"], out["v_sub"], out["v_neg"], out["v_mul"] = T(v + w), T(v - w), T(-v), T(v * w)
    out["v_dot"], out["v_outer"], out["v_abs"], out["v_norm"] = v.dot(w), v.outer(w), T(v.abs()), v.norm()
    out["v_cat"], out["v_between"], out["v_compose"], out["v_inverse"], out["v_log"] = T(v.cat(w)), T(v.between(w)), T(v.compose(w)), T(v.inverse()), v.log_map()
    pt = L.Point3(tensor=p3.clone()); out["p_x"], out["p_y"], out["p_z"] = pt.x(), pt.y(), pt.z()
    out["se3_project"] = L.SE3(tensor=dev(I["X3"]).clone()).project(dev(I["g"]))
    out["se3_project_sparse"] = L.SE3(tensor=dev(I["X3"]).clone()).project(dev(I["g2"]), is_sparse=True)
    out["se2_project_sparse"] = L.SE2(tensor=dev(I["X2"]).clone()).project(dev(I["g2"])[:, :, 0], is_sparse=True)
    return {k: v.detach().cpu() for k, v in out.items()}


def make_geom_api(th):
    """tests/golden/geom_api_kat.npz: inputs + the reference's outputs of geometry_api_values (operands on the tape)."""
    import torch
    d = torch.float64
    gen = torch.Generator().manual_seed(1)
    B = 4
    I = dict(X3=th.SE3.rand(B, generator=gen, dtype=d).tensor, Y3=th.SE3.rand(B, generator=gen, dtype=d).tensor,
             R3=th.SO3.rand(B, generator=gen, dtype=d).tensor, S3=th.SO3.rand(B, generator=gen, dtype=d).tensor,
             X2=th.SE2.rand(B, generator=gen, dtype=d).tensor, Y2=th.SE2.rand(B, generator=gen, dtype=d).tensor,
             R2=th.SO2.rand(B, generator=gen, dtype=d).tensor, S2=th.SO2.rand(B, generator=gen, dtype=d).tensor,
             p3=torch.randn(B, 3, generator=gen, dtype=d), p2=torch.randn(B, 2, generator=gen, dtype=d),
             t6=0.7 * torch.randn(B, 6, generator=gen, dtype=d), t3=0.7 * torch.randn(B, 3, generator=gen, dtype=d),
             t1=torch.randn(B, 1, generator=gen, dtype=d), g=torch.randn(B, 3, 4, generator=gen, dtype=d),
             g2=torch.randn(B, 5, 3, 4, generator=gen, dtype=d))
    q = torch.randn(B, 4, generator=gen, dtype=d)
    I["q"] = q / q.norm(dim=1, keepdim=True)
    out = geometry_api_values(th, torch, I, tape=True)
    out.update(geometry_api_values(th, torch, I, tape=False))      # + the Jacobian outputs of the group operations
    arrays = {"in_" + k: v.numpy() for k, v in I.items()}
    arrays.update({"out_" + k: v.numpy() for k, v in out.items()})
    np.savez_compressed(os.path.join(HERE, "geom_api_kat.npz"), **arrays)
    print("geom_api_kat:", len(out), "outputs")


def make_c5(th):
    """Config C5's problem at full size (2 500 SE3 poses on a sphere, 4 949 edges + prior, n = 15 000), one batch item, 3 LM iterations
    with the reference's CholeskyDenseSolver on the CPU (dense A is 3.6 GB, AtA 1.8 GB: minutes).  Input data from this repository's
    generator (theseus_b200.datasets.pose_graph_sphere: plain torch input fabrication), fixture in the pgo_* format."""
    import time
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(0, os.path.dirname(HERE))
    from theseus_b200.datasets import pose_graph_sphere
    from helpers import pgo_objective
    data = pose_graph_sphere(50, 50, 1, seed=0)
    E = len(data["edges"])
    g = dict(poses0=data["poses"].numpy(), edges=np.array(data["edges"], dtype=np.int64), meas=data["meas"].numpy(),
             edge_w=np.tile(data["info"].numpy().reshape(1, 1, 6), (E, 1, 1)), prior_w=np.array(1e-3), robust=np.array(""),
             log_loss_radius=np.array([[0.5]]))

    class G(dict):
        files = list(g.keys())
    t0 = time.time()
    objective, poses = pgo_objective(th, G(g), device="cpu")
    iters = 3
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, vectorize=True, max_iterations=iters, step_size=1.0,
                                abs_err_tolerance=0, rel_err_tolerance=0)
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.numpy().copy()); deltas.append(delta.numpy().copy())
        print("  c5 iteration", it, "err", info.last_err.numpy(), round(time.time() - t0, 1), "s", flush=True)
    lm = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
    with torch.no_grad():
        objective.update()
        g["err0"] = objective.error_metric().numpy()
        opt.optimize(end_iter_callback=cb, **lm)
    g["trace_err"], g["trace_delta"] = np.stack(errs, 0), np.stack(deltas, 0)
    g["poses_final"] = np.stack([p.tensor.numpy() for p in poses], 0)
    g["kwargs_json"] = np.array(repr(dict(method="lm", iters=iters, **lm)))
    np.savez_compressed(os.path.join(HERE, "pgo_c5_lm.npz"), **g)
    print("pgo_c5_lm err0", g["err0"], "->", g["trace_err"][-1], "took", round(time.time() - t0, 1), "s")


PGO_BENCHMARK_EXPECTED = dict(   # the literal values held by /root/reference/tests/theseus_tests/test_pgo_benchmark.py:34-39, 66-71
    dense_or_lu=[-0.29886279606812166, -0.3054215856589109, -0.27485602196709225, -0.3005231105990632],
    baspacho=[-0.2988627960682926, -0.30542158565900696, -0.27485602196705955, -0.3005231105991407])
PGO_BENCHMARK_CFG = dict(seed=1, num_poses=64, batch_size=16, dataset_size=256, max_num_batches=4, translation_noise=0.05, rotation_noise=0.02,
                         loop_closure_ratio=0.2, loop_closure_outlier_ratio=0.25, max_iters=10, step_size=0.75, reg_w=1e-3,
                         ratio_known_poses=0.1, lr=0.1, optimizer_kwargs={"backward_mode": "implicit", "track_err_history": True, "adaptive_damping": True,
                                                           "__keep_final_step_size__": True, "verbose": False})


def pgo_benchmark_run(th, torch, g, device="cpu", solver_kwargs=None):
    """examples/pose_graph/pose_graph_synthetic.py:87-286 (`run`) on the fixture's dataset, shared by generator (th = the reference) and
    tests (th = theseus_b200): Welsch-robust Between per edge with a learnable log radius, regularising prior on pose 0, priors on the
    known poses, LevenbergMarquardt(max 10 iterations, step 0.75) in implicit backward mode, one Adam step on the radius per batch.
    Returns the list of per-batch losses (what test_pgo_benchmark.py compares at rel 1e-10)."""
    cfg = PGO_BENCHMARK_CFG
    d = torch.float64
    Bsz, nb = cfg["batch_size"], cfg["max_num_batches"]
    P, GT, M = (torch.from_numpy(g[k]).to(device) for k in ("poses", "gt_poses", "meas"))      # [N, nb*B, 3, 4], ..., [E, nb*B, 3, 4]
    edges = [tuple(int(x) for x in e) for e in g["edges"]]
    info = torch.from_numpy(g["info"]).to(device).view(1, -1)
    N = P.shape[0]
    known = [int(i) for i in g["known_poses"]]

    def pose_loss(pose_tensors, gt_tensors):
        a = th.SE3(tensor=torch.cat(list(pose_tensors)))
        b_ = th.SE3(tensor=torch.cat(list(gt_tensors)))
        return th.local(a, b_).norm(dim=1).sum().view(1)

    sl0 = slice(0, Bsz)
    poses = [th.SE3(tensor=P[i, sl0].clone(), name=f"VERTEX_SE3__{i}") for i in range(N)]
    gts = [th.SE3(tensor=GT[i, sl0].clone(), name=f"VERTEX_SE3_GT__{i}") for i in range(N)]
    log_loss_radius = th.Vector(1, name="log_loss_radius", dtype=d)
    objective = th.Objective(dtype=d)
    rel_vars = []
    for e, (i, j) in enumerate(edges):
        z = th.SE3(tensor=M[e, sl0].clone(), name=f"EDGE_SE3__{i}_{j}")
        rel_vars.append(z)
        w = th.DiagonalCostWeight(th.Variable(info.clone()), name=f"EDGE_WEIGHT__{i}_{j}")
        objective.add(th.RobustCostFunction(cost_function=th.Between(poses[i], poses[j], z, w), loss_cls=th.WelschLoss,
                                            log_loss_radius=log_loss_radius))
    prior_t = th.SE3(tensor=P[0, sl0].clone(), name="VERTEX_SE3__0__PRIOR")
    objective.add(th.Difference(var=poses[0], target=prior_t, cost_weight=th.ScaleCostWeight(torch.tensor(cfg["reg_w"], dtype=d, device=device))))
    w100 = th.ScaleCostWeight(100 * torch.ones(1, dtype=d, device=device))
    for i in known:
        objective.add(th.Difference(poses[i], gts[i], w100, name=f"pose_diff_{i}"))
    optimizer = th.LevenbergMarquardt(objective.to(device), max_iterations=cfg["max_iters"], step_size=cfg["step_size"],
                                      **(solver_kwargs or dict(linear_solver_cls=th.CholeskyDenseSolver)))
    layer = th.TheseusLayer(optimizer)
    layer.to(device=device)
    radius = torch.nn.Parameter(torch.tensor([[3.0]], device=device, dtype=d))
    adam = torch.optim.Adam([radius], lr=cfg["lr"])
    losses = []
    for bi in range(nb):
        sl = slice(bi * Bsz, (bi + 1) * Bsz)
        inputs = {poses[i].name: P[i, sl] for i in range(N)}
        inputs[poses[0].name + "__PRIOR"] = P[0, sl].clone()
        inputs.update({gts[i].name: GT[i, sl] for i in known})
        inputs.update({rel_vars[e].name: M[e, sl] for e in range(len(edges))})
        inputs["log_loss_radius"] = radius.clone()
        with torch.no_grad():
            ref = pose_loss([P[i, sl] for i in range(N)], [GT[i, sl] for i in range(N)])
        out, _ = layer.forward(input_tensors=inputs, optimizer_kwargs=dict(cfg["optimizer_kwargs"]))
        adam.zero_grad()
        loss = (pose_loss([out[poses[i].name] for i in range(N)], [GT[i, sl] for i in range(N)]) - ref) / ref
        loss.backward()
        adam.step()
        losses.append(float(loss.item()))
    return losses


def make_pgo_benchmark(th):
    """The dataset of the reference's own end-to-end KAT (tests/theseus_tests/test_pgo_benchmark.py: 64 poses, batch 16, 4 batches, seed
    1): generated by PoseGraphDataset.generate_synthetic_3D exactly like examples/pose_graph/pose_graph_synthetic.py:89-106, checked here
    (reference on the CPU, CholeskyDenseSolver) against the literal losses of the test, and stored with the known-pose draw."""
    import random
    import torch
    import theseus.utils.examples as theg
    cfg = PGO_BENCHMARK_CFG
    torch.manual_seed(cfg["seed"]); np.random.seed(cfg["seed"]); random.seed(cfg["seed"])
    rng = torch.Generator(); rng.manual_seed(0)
    pg, _ = theg.PoseGraphDataset.generate_synthetic_3D(
        num_poses=cfg["num_poses"], translation_noise=cfg["translation_noise"], rotation_noise=cfg["rotation_noise"],
        loop_closure_ratio=cfg["loop_closure_ratio"], loop_closure_outlier_ratio=cfg["loop_closure_outlier_ratio"],
        batch_size=cfg["batch_size"], dataset_size=cfg["dataset_size"], generator=rng, dtype=torch.float64)
    known = [i for i in range(cfg["num_poses"]) if not (np.random.rand() > cfg["ratio_known_poses"])]   # same draw as the example's loop
    n_items = cfg["batch_size"] * cfg["max_num_batches"]
    g = dict(poses=np.stack([p.tensor[:n_items].numpy() for p in pg.poses], 0), gt_poses=np.stack([p.tensor[:n_items].numpy() for p in pg.gt_poses], 0),
             meas=np.stack([e.relative_pose.tensor[:n_items].numpy() if e.relative_pose.tensor.shape[0] > 1 else
                            e.relative_pose.tensor.expand(n_items, 3, 4).numpy() for e in pg.edges], 0),
             edges=np.array([(e.i, e.j) for e in pg.edges], dtype=np.int64), info=pg.edges[0].weight.diagonal.tensor.numpy().reshape(-1),
             known_poses=np.array(known, dtype=np.int64))
    losses = pgo_benchmark_run(th, torch, g)
    exp = PGO_BENCHMARK_EXPECTED["dense_or_lu"]
    rel = max(abs(a - b) / abs(b) for a, b in zip(losses, exp))
    print("pgo_benchmark: reference losses here", losses, "max rel diff vs the literal KAT", rel)
    assert rel < 1e-9
    g["losses_reference_here"] = np.array(losses)
    np.savez_compressed(os.path.join(HERE, "pgo_benchmark_kat.npz"), **g)
    print("pgo_benchmark_kat", {k: v.shape for k, v in g.items()})



def make_backward(th):
    """End-to-end gradients through TheseusLayer (theseus_layer.py:45-97) in the reference's backward modes, dense solver, fp64."""
    import torch
    out = {}
    for name, (method, iters, kw) in BACKWARD_CASES.items():
        objective = backward_problem(th, torch)
        cls = th.GaussNewton if method == "gn" else th.LevenbergMarquardt
        opt = cls(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0)
        layer = th.TheseusLayer(opt)
        inp, w = backward_inputs(torch)
        leaves = {k: inp[k].clone().requires_grad_(True) for k in ("x", "y", "t")}
        sol, info = layer.forward({**leaves, "ab": inp["ab"].clone(), "c": inp["c"].clone()}, optimizer_kwargs=dict(kw, track_err_history=True))
        eh = info.err_history.numpy()
        out[name + "_err_history"] = eh
        print("   rel. error reduction per iteration:", np.array2string(((eh[:, :-1] - eh[:, 1:]) / eh[:, :-1]).min(axis=0), precision=2))
        loss = (torch.cat([sol["ab"], sol["c"]], 1) ** 2 * w).sum()
        loss.backward()
        out[name + "_ab"] = sol["ab"].detach().numpy(); out[name + "_c"] = sol["c"].detach().numpy()
        for k in leaves:
            out[name + "_grad_" + k] = leaves[k].grad.numpy()
        out[name + "_loss"] = np.array(loss.item())
        print("backward", name, "loss", loss.item(), "|grad_y|", float(leaves["y"].grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "backward_kat.npz"), **out)


if __name__ == "__main__":
    th, lieF = _import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "tactile_c4":
        make_tactile(th, T=25, B=512, name="tactile_c4_kat", seed=18, keep_deltas=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pgo_benchmark":
        make_pgo_benchmark(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "backward":
        make_backward(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "autodiff_lie":
        make_autodiff_lie(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ba_c3":   # config C3 at full size (50 cameras x 1000 points x 8 observations per point, Huber)
        make_ba(th, "ba_c3_huber", num_cameras=50, num_points=1000, B=2, seed=11, iters=5, robust=True, track_length=8)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dogleg":
        make_pgo(th, "pgo_small_dogleg", num_poses=8, B=4, seed=12, iters=8, lm_kwargs=dict(trust_region_init=0.3), method="dogleg",
                 loop_closure_ratio=0.5, init_perturb=0.6, full_trace=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "backward_pgo":
        make_backward_pgo(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c5":
        make_c5(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "geom_api":
        make_geom_api(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tactile":
        make_tactile(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "moving_frame":
        make_moving_frame(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "io":
        make_io(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "robust":
        make_robust(th)
        make_pgo(th, "pgo_small_geman", num_poses=10, B=3, seed=13, iters=8, lm_kwargs=dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True),
                 loop_closure_ratio=0.6, robust="geman", outlier_ratio=0.3, init_perturb=0.0)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "so2":
        make_so2(th)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "backward_lie":
        make_backward_lie(th)
        sys.exit(0)
    make_lie(th, lieF)
    make_costs(th)
    make_dense_solver(th)
    lm = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
    make_pgo(th, "pgo_small_lm", num_poses=8, B=3, seed=0, iters=6, lm_kwargs=lm, loop_closure_ratio=0.5)
    make_pgo(th, "pgo_small_gn", num_poses=8, B=3, seed=1, iters=4, lm_kwargs={}, method="gn", loop_closure_ratio=0.5)
    make_pgo(th, "pgo_small_lm_sph", num_poses=8, B=3, seed=2, iters=3,
             lm_kwargs=dict(damping=0.1, adaptive_damping=True, ellipsoidal_damping=False), loop_closure_ratio=0.5)
    make_pgo(th, "pgo64_lm", num_poses=64, B=2, seed=3, iters=10, lm_kwargs=lm, full_trace=False)
    make_pgo(th, "pgo_small_lm_hard", num_poses=8, B=4, seed=4, iters=8, lm_kwargs=lm, loop_closure_ratio=0.5, init_perturb=0.6)
    make_pgo(th, "pgo32_lm_hard", num_poses=32, B=3, seed=5, iters=10, lm_kwargs=lm, full_trace=False, init_perturb=0.5)
    make_ba(th, "ba_small_lm", num_cameras=6, num_points=40, B=3, seed=7, iters=8)
    make_simple_example(th)
    make_se2(th)
    make_so3(th)
    make_ba(th, "ba_small_huber", num_cameras=6, num_points=40, B=3, seed=8, iters=8, robust=True)
    make_pgo(th, "pgo_small_welsch", num_poses=10, B=3, seed=9, iters=8, lm_kwargs=lm, loop_closure_ratio=0.6, robust="welsch",
             outlier_ratio=0.3, init_perturb=0.0)
    make_backward(th)
    make_autodiff_lie(th)
    make_backward_lie(th)
    make_ba(th, "ba_c3_huber", num_cameras=50, num_points=1000, B=2, seed=11, iters=5, robust=True, track_length=8)
    make_io(th)
    make_pgo(th, "pgo_small_dogleg", num_poses=8, B=4, seed=12, iters=8, lm_kwargs=dict(trust_region_init=0.3), method="dogleg",
             loop_closure_ratio=0.5, init_perturb=0.6, full_trace=False)
    make_backward_pgo(th)
    make_moving_frame(th)
    make_tactile(th)
    make_so2(th)
    make_robust(th)
    make_pgo(th, "pgo_small_geman", num_poses=10, B=3, seed=13, iters=8, lm_kwargs=lm, loop_closure_ratio=0.6, robust="geman",
             outlier_ratio=0.3, init_perturb=0.0)
