"""Multifrontal block-sparse Cholesky (layout "front"): the symbolic plan (theseus_b200/frontal.py) executed by a numpy interpreter of
the very arrays the kernels consume, and the kernels of theseus_b200/csrc/thb_front.cu on the host emulation (tests/simt; the DMMA
m8n8k4 fragment semantics emulated with shuffles) driven through BaspachoSparseSolver's own host code.  The big-front path (dense DMMA
kernel in partial mode) needs the device: tests/test_gpu_front.py."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from theseus_b200 import _lib
from theseus_b200.frontal import build_front_plan, execute_numpy, nested_dissection_order, small_smem_bytes
from theseus_b200.structure import build_structure


def _graph_csr(N, edges):
    adj = [set([i]) for i in range(N)]
    for a, b in edges:
        adj[a].add(b); adj[b].add(a)
    ptrs = np.zeros(N + 1, dtype=np.int64)
    inds = []
    for i in range(N):
        inds.extend(sorted(adj[i])); ptrs[i + 1] = len(inds)
    return ptrs, np.array(inds, dtype=np.int64)


def _grid(r, c):
    e = []
    for i in range(r):
        for j in range(c):
            if j + 1 < c: e.append((i * c + j, i * c + j + 1))
            if i + 1 < r: e.append((i * c + j, (i + 1) * c + j))
    return e


def _check_plan(N, edges, dims, seed=0, **kw):
    rng = np.random.default_rng(seed)
    ptrs, inds = _graph_csr(N, edges)
    plan = build_front_plan(dims, ptrs, inds, **kw)
    n = plan.n
    start = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(int)
    M = np.zeros((n, n))
    for (a, b) in list(edges) + [(i, i) for i in range(N)]:
        idx = np.arange(start[a], start[a] + dims[a]) if a == b else np.concatenate(
            [np.arange(start[a], start[a] + dims[a]), np.arange(start[b], start[b] + dims[b])])
        J = rng.standard_normal((4, len(idx)))
        M[np.ix_(idx, idx)] += J.T @ J
    M += np.eye(n)
    panels = np.zeros(plan.data_size)
    f = plan.gram_out_offsets()
    for a in range(N):
        for b_ in inds[ptrs[a]:ptrs[a + 1]]:
            if plan.pos[a] >= plan.pos[b_]:
                off, ld, _ = f(a, int(b_))
                blk = M[start[a]:start[a] + dims[a], start[b_]:start[b_] + dims[b_]]
                for p in range(dims[a]):
                    panels[off + p * ld:off + p * ld + dims[b_]] = blk[p]
    rhs = rng.standard_normal(n)
    x, _ = execute_numpy(plan, panels, rhs)
    xr = np.linalg.solve(M, rhs)
    assert np.abs(x - xr).max() <= 1e-10 * np.abs(xr).max()
    # structural invariants the kernels rely on
    A = plan.arrays
    assert sorted(plan.order.tolist()) == list(range(N)) and sorted(A["sched"].tolist()) == list(range(plan.S))
    for t in range(plan.S):
        p = int(A["f_parent"][t])
        rel = A["f_rel"][A["rel_ptr"][t]:A["rel_ptr"][t + 1]]
        assert (p < 0) == (A["f_b"][t] == 0)
        if p >= 0:
            assert A["f_depth"][t] == A["f_depth"][p] + 1 and len(rel) == A["f_b"][t]
            assert np.all(np.diff(rel) > 0) and rel[-1] < A["f_w"][p] + A["f_b"][p]
        if A["f_class"][t] < 3:
            assert small_smem_bytes(int(A["f_w"][t]), int(A["f_b"][t]), int(A["child_ptr"][t + 1] - A["child_ptr"][t])) <= 220 * 1024
    depths = plan.launches[:, 0]
    assert np.all(np.diff(depths) <= 0)
    return plan


def test_plans_solve_random_systems():
    rng = np.random.default_rng(1)
    for N in (1, 2, 5, 12, 40):
        edges = [(int(a), int(b)) for a, b in rng.integers(0, N, size=(2 * N, 2)) if a != b]
        _check_plan(N, edges, rng.choice([1, 2, 3, 6], size=N))
    _check_plan(100, _grid(10, 10), np.full(100, 6))
    _check_plan(100, _grid(10, 10), np.full(100, 6), ordering="mindeg")
    _check_plan(64, _grid(8, 8), np.full(64, 3), ordering="natural")
    cams, pts = 6, 60
    e = [(cams + p, int(c)) for p in range(pts) for c in rng.choice(cams, size=3, replace=False)] + [(a, b) for a in range(cams) for b in range(a)]
    plan = _check_plan(cams + pts, e, np.array([6] * cams + [3] * pts))
    assert plan.stats["flops"] < 3 * plan.stats["column_flops"]     # amalgamation must not blow a bundle-adjustment structure up


def test_big_fronts_and_disconnected_components_in_the_plan():
    # small_limit forces "big" fronts (padded dense matrices in the arena): same interpreter, layout arrays checked
    e = _grid(12, 12)
    plan = _check_plan(144, e, np.full(144, 6), small_limit=60)
    A = plan.arrays
    big = np.nonzero(A["f_class"] == 3)[0]
    assert len(big) > 0
    for t in big:
        assert A["f_wpad"][t] % 64 == 0 and A["f_np"][t] % 128 == 0 and A["f_np"][t] >= A["f_wpad"][t] + A["f_b"][t]
        assert A["f_cb_ld"][t] == A["f_np"][t] and A["f_cb_off"][t] == A["f_fr_off"][t] + A["f_wpad"][t] * A["f_np"][t] + A["f_wpad"][t]
    two = e + [(144 + a, 144 + b) for a, b in _grid(3, 3)]
    _check_plan(153, two, np.full(153, 6))


def test_nested_dissection_beats_minimum_degree_on_the_c5_topology():
    N = 2500
    e = [(i, i + 1) for i in range(N - 1)] + [(i, i + 50) for i in range(N - 50)]
    ptrs, inds = _graph_csr(N, e)
    plan = build_front_plan(np.full(N, 6), ptrs, inds)
    assert plan.stats["ordering"] == "nd"
    assert plan.stats["column_flops"] < 0.30e9 and plan.stats["flops"] < 0.33e9     # round 1's minimum degree: 0.415 GFLOP
    assert sorted(nested_dissection_order(N, ptrs, inds).tolist()) == list(range(N))


# ------------------------------------------------------------------------------------------------ emulated kernels
def _ring_structure(N, dims=None, chord=7):
    dims = dims or [6] * N
    costs = [(3, [i, (i + 1) % N]) for i in range(N)] + [(3, [i, (i + chord) % N]) for i in range(N)] + [(dims[i], [i]) for i in range(N)]
    return build_structure(dims, [(d, sorted(vs)) for d, vs in costs])


def _dense_system(S, A_val, b):
    B = A_val.shape[0]
    A = np.zeros((B, S.num_rows, S.num_cols))
    for r in range(S.num_rows):
        A[:, r, S.A_col_ind[S.A_row_ptr[r]:S.A_row_ptr[r + 1]]] = A_val[:, S.A_row_ptr[r]:S.A_row_ptr[r + 1]]
    return np.einsum("bri,brj->bij", A, A), np.einsum("bri,br->bi", A, b)


@pytest.fixture(scope="module")
def emulated():
    import importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("emulation_mode", os.path.join(here, "simt", "emulation_mode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.load_emulated_lib()


def _solve(monkeypatch, lib, S, A_val, b, alpha, **kw):
    monkeypatch.setattr(_lib, "load", lambda: lib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    solver = th.BaspachoSparseSolver.from_structure(S, layout="front", **kw)
    solver.linearization.A_val, solver.linearization.b = torch.from_numpy(A_val), torch.from_numpy(b)
    x = solver.solve(damping=torch.from_numpy(alpha), ellipsoidal_damping=True, damping_eps=1e-6)
    return solver, x.numpy()


@pytest.mark.parametrize("case", ["ring40", "mixed", "ba", "islands"])
def test_front_kernels_on_the_host_emulation(monkeypatch, emulated, case):
    rng = np.random.default_rng(7)
    if case == "ring40":
        S, B = _ring_structure(40), 3
    elif case == "islands":
        # disconnected components (several roots), an isolated variable that only has a prior, dof 1 and 7 next to 6, batch of ONE
        dims = [6, 6, 6, 6, 6, 1, 7, 3, 3, 3]
        costs = [(3, [0, 1]), (3, [1, 2]), (3, [2, 3]), (3, [0, 3]), (2, [7, 8]), (2, [8, 9]), (4, [5, 6])] + [(dims[i], [i]) for i in range(len(dims))]
        S, B = build_structure(dims, [(d, sorted(vs)) for d, vs in costs]), 1
    elif case == "mixed":
        dims = [6, 3, 6, 2, 1, 6, 3, 3, 6, 6, 2, 6]
        S, B = _ring_structure(len(dims), dims, chord=5), 2
    else:
        P_, Cn, B = 8, 4, 2          # (<= 8 children per front: more go to the dense path's scatter assembly, which needs the device)
        dims = [6] * Cn + [3] * P_
        costs = [(2, sorted([int(c), Cn + p])) for p in range(P_) for c in rng.choice(Cn, size=3, replace=False)]
        costs += [(dims[i], [i]) for i in range(len(dims))]
        S = build_structure(dims, costs)
    A_val = rng.standard_normal((B, S.nnz)); b = rng.standard_normal((B, S.num_rows)); alpha = rng.random(B) * 0.1
    solver, x = _solve(monkeypatch, emulated, S, A_val, b, alpha)
    assert solver._plan.S >= 1 and (solver._plan.arrays["f_class"] < 3).all()
    AtA, Atb = _dense_system(S, A_val, b)
    idx = np.arange(S.num_cols)
    M = AtA.copy(); M[:, idx, idx] = M[:, idx, idx] * (1 + alpha[:, None]) + 1e-6
    ref = np.linalg.solve(M, Atb[..., None])[..., 0]
    assert np.abs(x - ref).max() <= 1e-11 * np.linalg.cond(M).max() * max(1.0, np.abs(ref).max())
    # A^T b and diag(A^T A) of the Gram pass: what the LM accept test reads from the linearization instead of a second Atb kernel
    bufs = solver._dev["bufs"]
    assert np.allclose(bufs["Atb"].numpy(), Atb, rtol=1e-12, atol=1e-12)
    assert np.allclose(bufs["AtA_diag"].numpy(), AtA[:, idx, idx], rtol=1e-12, atol=1e-12)


def test_front_kernels_report_a_non_positive_pivot(monkeypatch, emulated):
    S2 = build_structure([2, 2], [(2, [0, 1])])
    monkeypatch.setattr(_lib, "load", lambda: emulated)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    solver = th.BaspachoSparseSolver.from_structure(S2, layout="front")
    A2 = torch.ones(3, S2.nnz, dtype=torch.float64)
    for r in range(S2.num_rows):
        cols = S2.A_col_ind[S2.A_row_ptr[r]:S2.A_row_ptr[r + 1]]
        A2[:, S2.A_row_ptr[r]:S2.A_row_ptr[r + 1]][:, cols >= 2] = 0.0
    A2[:, 1] = 2.0
    solver.linearization.A_val = A2
    solver.linearization.b = torch.ones(3, S2.num_rows, dtype=torch.float64)
    with pytest.raises(RuntimeError, match=r"batch element 0: matrix is not positive definite"):
        solver.solve()


def test_front_solver_processes_the_batch_in_chunks(monkeypatch, emulated):
    """front_options['chunk']: the update-matrix arena is sized for a chunk of the batch, the factor stays resident for every item."""
    rng = np.random.default_rng(9)
    S, B = _ring_structure(12), 3
    A_val = rng.standard_normal((B, S.nnz)); b = rng.standard_normal((B, S.num_rows)); alpha = rng.random(B) * 0.1
    _, x1 = _solve(monkeypatch, emulated, S, A_val, b, alpha)
    solver, x2 = _solve(monkeypatch, emulated, S, A_val, b, alpha, front_options=dict(chunk=2))
    assert solver._dev["bufs"]["arena"].shape[1] == 2
    assert np.array_equal(x1, x2)
