"""Shared test helpers: load golden fixtures, build the same pose-graph problem for the oracle (numpy spec)
and for the product (theseus_b200 objective)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Fixture(dict):
    """A golden fixture read ONCE into memory (an NpzFile decompresses the whole array on every g[key]: 20 s for the 8000-observation
    bundle-adjustment fixture when accessed inside loops).  Keeps the `.files` attribute the helpers use."""

    @property
    def files(self):
        return list(self.keys())


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return Fixture({k: z[k] for k in z.files})


def pgo_spec(g, dtype=np.float64):
    """Oracle problem description from a pgo_*.npz fixture (layout of oracle/nls.py)."""
    poses0, edges, meas, edge_w = g["poses0"], g["edges"], g["meas"], g["edge_w"]
    spec = dict(dtype=np.dtype(dtype), vars=[], costs=[])
    for i in range(poses0.shape[0]):
        spec["vars"].append(dict(kind="SE3", dof=6, value=poses0[i].astype(dtype)))
    robust = str(g["robust"]) if "robust" in g.files else ""
    for e in range(edges.shape[0]):
        c = dict(kind="between", group="SE3", vars=(int(edges[e, 0]), int(edges[e, 1])),
                 aux=meas[e].astype(dtype), weight=("diag", edge_w[e].astype(dtype)))
        if robust:
            c["robust"] = (robust, g["log_loss_radius"].astype(dtype)) + ((g["gnc_mu"].astype(dtype),) if robust == "geman" else ())
        spec["costs"].append(c)
    spec["costs"].append(dict(kind="local", group="SE3", vars=(0,), aux=poses0[0].astype(dtype),
                              weight=("scale", np.full((1, 1), float(g["prior_w"]), dtype=dtype))))
    return spec


def user_local_cost_cls(th):
    class UserLocal(th.CostFunction):
        """A user-written twin of th.Difference (embodied/misc/local_cost_fn.py:15-70) on the reference's plugin contract: the error
        and its Jacobian come from the library's own group methods, e = log(T^-1 X), J = dlog."""

        def __init__(self, var, target, cost_weight, name=None):
            super().__init__(cost_weight, name=name)
            self.var, self.target = var, target
            self.register_optim_var("var")
            self.register_aux_var("target")

        def dim(self):
            return self.var.dof()

        def error(self):
            return self.target.local(self.var)

        def jacobians(self):
            J = []
            err = self.target.between(self.var).log_map(jacobians=J)
            return [J[0]], err
    return UserLocal


def pgo_objective(th, g, device="cuda", dtype=None, user_prior=False):
    """The product objective built exactly like examples/pose_graph/pose_graph_cube.py:56-83.  user_prior=True: the gauge prior is a
    user-defined cost function (user_local_cost_cls) instead of th.Difference -- same numbers, generic route next to the fused groups."""
    import torch
    dtype = dtype or torch.float64
    poses0, edges, meas, edge_w = g["poses0"], g["edges"], g["meas"], g["edge_w"]
    poses = [th.SE3(tensor=torch.from_numpy(poses0[i]).to(dtype), name=f"VERTEX_SE3__{i}") for i in range(poses0.shape[0])]
    objective = th.Objective(dtype=dtype)
    robust = str(g["robust"]) if "robust" in g.files else ""
    llr = th.Vector(tensor=torch.from_numpy(g["log_loss_radius"]).to(dtype), name="log_loss_radius") if robust else None
    mu = th.Vector(tensor=torch.from_numpy(g["gnc_mu"]).to(dtype), name="gnc_mu") if robust == "geman" else None
    for e in range(edges.shape[0]):
        i, j = int(edges[e, 0]), int(edges[e, 1])
        z = th.SE3(tensor=torch.from_numpy(meas[e]).to(dtype), name=f"EDGE_SE3__{e}_{i}_{j}")
        w = th.DiagonalCostWeight(th.Variable(torch.from_numpy(edge_w[e]).to(dtype), name=f"EDGE_WEIGHT__{e}"))
        cf = th.Between(poses[i], poses[j], z, w, name=f"between_{e}")
        if robust == "geman":
            cf = th.GNCRobustCostFunction(cf, th.GemanMcClureLoss, llr, mu, name=f"robust_between_{e}")
        elif robust:
            cf = th.RobustCostFunction(cf, dict(welsch=th.WelschLoss, huber=th.HuberLoss, hinge=th.HingeLoss)[robust], llr,
                                       name=f"robust_between_{e}")
        objective.add(cf)
    prior_cls = user_local_cost_cls(th) if user_prior else th.Difference
    prior = prior_cls(poses[0], th.SE3(tensor=torch.from_numpy(poses0[0]).to(dtype), name="VERTEX_SE3__0__PRIOR"),
                      th.ScaleCostWeight(torch.tensor(float(g["prior_w"]), dtype=dtype)), name="prior")
    objective.add(prior)
    objective.to(device)
    return objective, poses


def lm_kwargs_of(g):
    kw = eval(str(g["kwargs_json"]))  # written by tests/golden/make_golden.py (repr of a plain dict)
    method = kw.pop("method")
    iters = kw.pop("iters")
    return method, iters, kw


def decisive_iterations(err0, trace_err, tol=1e-7):
    """Number of leading LM iterations whose accept/reject decision is numerically well defined.

    rho = (err_prev - err_new) / den is 0/0 noise once the problem has converged to rounding level (the reference's
    own decisions become arbitrary there: pose-graph LM reaches its floor in 3-5 iterations).  An iteration is
    'decisive' if the reference reduced the error by more than `tol` relative, for every batch item."""
    prev = err0
    k = 0
    for it in range(trace_err.shape[0]):
        red = (prev - trace_err[it]) / prev
        if not np.all(red > tol):
            break
        prev = trace_err[it]
        k += 1
    return k


def ba_spec(g, dtype=np.float64):
    """Oracle problem description of the bundle-adjustment fixture (tests/golden/ba_small_lm.npz)."""
    order = [str(x) for x in g["order"]]
    idx = {n: i for i, n in enumerate(order)}
    spec = dict(dtype=np.dtype(dtype), vars=[], costs=[])
    for n in order:
        if n.startswith("Cam"):
            spec["vars"].append(dict(kind="SE3", dof=6, value=g["cam_pose0"][int(n[3:].split("_")[0])].astype(dtype)))
        else:
            spec["vars"].append(dict(kind="Vector", dof=3, value=g["pts0"][int(n[2:])].astype(dtype)))
    one = np.ones((1, 1), dtype=dtype)
    for o in range(g["obs_cam"].shape[0]):
        c, p = int(g["obs_cam"][o]), int(g["obs_pt"][o])
        cd = dict(kind="reproj", vars=(idx[f"Cam{c}_pose"], idx[f"Pt{p}"]),
                  aux=dict(f=g["focal"][c], z=g["feats"][o], k1=g["k1"][c], k2=g["k2"][c]), weight=("scale", one))
        if "robust" in g.files and str(g["robust"]):
            cd["robust"] = (str(g["robust"]), np.zeros((1, 1), dtype=dtype))
        spec["costs"].append(cd)
    w = np.full((1, 1), np.sqrt(1e-4), dtype=dtype)
    eye = np.eye(3, 4, dtype=dtype)[None]
    for n in [str(x) for x in g["reg_order"]]:
        if n.startswith("Cam"):
            spec["costs"].append(dict(kind="local", group="SE3", vars=(idx[n],), aux=eye, weight=("scale", w)))
        else:
            spec["costs"].append(dict(kind="local", group="Vector", vars=(idx[n],), aux=np.zeros((1, 3), dtype=dtype), weight=("scale", w)))
    for q, i in enumerate(g["known"]):
        spec["costs"].append(dict(kind="local", group="SE3", vars=(idx[f"Cam{int(i)}_pose"],), aux=g["known_pose"][q].astype(dtype),
                                  weight=("scale", np.full((1, 1), 100.0, dtype=dtype))))
    return spec


def ba_objective(th, g, device="cuda"):
    """The product objective built in the same order as examples/bundle_adjustment.py:106-164."""
    import torch
    dtype = torch.float64
    Nc, Np = g["cam_pose0"].shape[0], g["pts0"].shape[0]
    cams = [th.SE3(tensor=torch.from_numpy(g["cam_pose0"][i]), name=f"Cam{i}_pose") for i in range(Nc)]
    pts = [th.Point3(tensor=torch.from_numpy(g["pts0"][i]), name=f"Pt{i}") for i in range(Np)]
    focal = [th.Vector(tensor=torch.from_numpy(g["focal"][i]), name=f"Cam{i}_focal_length") for i in range(Nc)]
    k1 = [th.Vector(tensor=torch.from_numpy(g["k1"][i]), name=f"Cam{i}_calib_k1") for i in range(Nc)]
    k2 = [th.Vector(tensor=torch.from_numpy(g["k2"][i]), name=f"Cam{i}_calib_k2") for i in range(Nc)]
    objective = th.Objective(dtype=dtype)
    weight = th.ScaleCostWeight(th.Variable(torch.ones(1, 1, dtype=dtype), name="reproj_weight"))
    for o in range(g["obs_cam"].shape[0]):
        c, p = int(g["obs_cam"][o]), int(g["obs_pt"][o])
        cf = th.eb.Reprojection(camera_pose=cams[c], world_point=pts[p], focal_length=focal[c], calib_k1=k1[c], calib_k2=k2[c],
                                image_feature_point=th.Point2(tensor=torch.from_numpy(g["feats"][o]), name=f"Feat{o}"),
                                weight=weight, name=f"reproj_{o}")
        if "robust" in g.files and str(g["robust"]):
            if o == 0:
                huber_radius = th.Vector(tensor=torch.zeros(1, 1, dtype=dtype), name="log_loss_radius")
            cf = th.RobustCostFunction(cf, th.HuberLoss, huber_radius, name=f"robust_reproj_{o}")
        objective.add(cf)
    zero_point3 = th.Point3(tensor=torch.zeros(1, 3, dtype=dtype), name="zero_point")
    identity_se3 = th.SE3(tensor=torch.eye(3, 4, dtype=dtype).view(1, 3, 4), name="zero_se3")
    damping_weight = th.ScaleCostWeight(th.Variable(torch.full((1, 1), float(np.sqrt(1e-4)), dtype=dtype), name="reg_weight"))
    for name in list(objective.optim_vars.keys()):
        var = objective.optim_vars[name]
        objective.add(th.Difference(var, identity_se3 if isinstance(var, th.SE3) else zero_point3, damping_weight, name=f"reg_{name}"))
    camera_weight = th.ScaleCostWeight(th.Variable(torch.full((1, 1), 100.0, dtype=dtype), name="camera_weight"))
    for q, i in enumerate(g["known"]):
        objective.add(th.Difference(cams[int(i)], th.SE3(tensor=torch.from_numpy(g["known_pose"][q]), name=f"Cam{int(i)}_gt_pose"),
                                    camera_weight, name=f"camera_diff_{int(i)}"))
    objective.to(device)
    return objective, cams, pts


def se2_pg_spec(g, dtype=np.float64):
    spec = dict(dtype=np.dtype(dtype), vars=[], costs=[])
    for i in range(g["pg_poses0"].shape[0]):
        spec["vars"].append(dict(kind="SE2", dof=3, value=g["pg_poses0"][i].astype(dtype)))
    w = np.array([[10.0, 10.0, 20.0]], dtype=dtype)
    for k, (i, j) in enumerate(g["pg_edges"]):
        spec["costs"].append(dict(kind="between", group="SE2", vars=(int(i), int(j)), aux=g["pg_meas"][k].astype(dtype), weight=("diag", w)))
    spec["costs"].append(dict(kind="local", group="SE2", vars=(0,), aux=g["pg_prior"].astype(dtype), weight=("scale", np.full((1, 1), 5.0, dtype=dtype))))
    return spec


def se2_pg_objective(th, g, device="cuda"):
    import torch
    dt = torch.float64
    poses = [th.SE2(tensor=torch.from_numpy(g["pg_poses0"][i]), name=f"P{i}") for i in range(g["pg_poses0"].shape[0])]
    objective = th.Objective(dtype=dt)
    for k, (i, j) in enumerate(g["pg_edges"]):
        z = th.SE2(tensor=torch.from_numpy(g["pg_meas"][k]), name=f"Z{k}")
        objective.add(th.Between(poses[int(i)], poses[int(j)], z,
                                 th.DiagonalCostWeight(th.Variable(torch.tensor([[10.0, 10.0, 20.0]], dtype=dt), name=f"W{k}")), name=f"btw{k}"))
    objective.add(th.Difference(poses[0], th.SE2(tensor=torch.from_numpy(g["pg_prior"]), name="P0_prior"),
                                th.ScaleCostWeight(torch.tensor(5.0, dtype=dt)), name="prior"))
    objective.to(device)
    return objective, poses
