"""Mirror of the reference's native solver module `theseus.extlib.baspacho_solver` (pybind11 API of extlib/baspacho_solver.cpp:326-358):

    s = SymbolicDecomposition(param_size, sparse_struct_ptrs, sparse_struct_inds, dev)     # int64 CPU tensors, dev = "cuda"
    f = s.create_numeric_decomposition(batch_size)
    f.add_M(val [B,nnz] f64, ptrs, inds)     # scalar CSR, entries with column <= row are added (baspacho_solver_cuda.cu:19-58)
    f.add_MtM(val [B,nnz] f64, ptrs, inds)   # += M^T M for a scalar-CSR M (baspacho_solver_cuda.cu:96-134)
    f.damp(alpha [B], beta [B])              # diag <- diag (1 + alpha) + beta (baspacho_solver_cuda.cu:171-185)
    f.factor()
    f.solve(x [B,n])                         # in place

on the multifrontal layout of this package (frontal.py + csrc/thb_front.cu).  The hot path does not go through this class
(BaspachoSparseSolver drives the fused Gram / damping / factor calls directly); it exists so that code written against the reference's
extension module -- and the reference's own tests of it (tests/theseus_tests/extlib/test_baspacho*.py) -- run unchanged.  The index
tables of add_M / add_MtM are built on the host once per CSR pattern; the scatter itself is a device index_add_ (no atomics, fixed order).
"""
import ctypes as C
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from .frontal import FrontPlan, build_front_plan


class SymbolicDecomposition:
    def __init__(self, param_size: torch.Tensor, sparse_struct_ptrs: torch.Tensor, sparse_struct_inds: torch.Tensor, dev: str = "cuda"):
        if str(dev) not in ("cuda",) and not str(dev).startswith("cuda"):
            raise RuntimeError("theseus_b200.extlib: only dev='cuda' exists (no CPU implementation of the factorisation in this package)")
        ps = np.asarray(param_size.cpu().numpy() if torch.is_tensor(param_size) else param_size, dtype=np.int64)
        ptrs = np.asarray(sparse_struct_ptrs.cpu().numpy() if torch.is_tensor(sparse_struct_ptrs) else sparse_struct_ptrs, dtype=np.int64)
        inds = np.asarray(sparse_struct_inds.cpu().numpy() if torch.is_tensor(sparse_struct_inds) else sparse_struct_inds, dtype=np.int64)
        N = int(ps.shape[0])
        if ptrs.shape[0] != N + 1:
            raise RuntimeError("sparse_struct_ptrs must have len(param_size) + 1 entries")
        # the reference is handed either the full symmetric block pattern (baspacho_sparse_solver.py:93-113) or its lower triangle
        # (tests/theseus_tests/extlib/test_baspacho_simple.py): symmetrise, add the diagonal
        nbr = [set([i]) for i in range(N)]
        for i in range(N):
            for j in inds[ptrs[i]:ptrs[i + 1]]:
                nbr[i].add(int(j)); nbr[int(j)].add(i)
        fp = np.zeros(N + 1, dtype=np.int64)
        fi = []
        for i in range(N):
            fi.extend(sorted(nbr[i])); fp[i + 1] = len(fi)
        self.param_size = ps
        self.plan: FrontPlan = build_front_plan(ps, fp, np.array(fi, dtype=np.int64))
        self.device = torch.device(dev if str(dev) != "cuda" else "cuda")
        self.param_start = np.concatenate([[0], np.cumsum(ps)[:-1]]).astype(np.int64)
        self.to_param = np.repeat(np.arange(N), ps)
        self._dev = None
        self._tables: Dict[Tuple, Tuple[torch.Tensor, ...]] = {}

    def create_numeric_decomposition(self, batch_size: int) -> "NumericDecomposition":
        return NumericDecomposition(self, int(batch_size))

    # ---- host-side tables ----
    def _device_plan(self):
        if self._dev is None:
            P = self.plan
            dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in P.arrays.items()}
            keys = ("f_w", "f_b", "f_first", "f_class", "f_wpad", "f_np", "f_cb_ld", "f_depth", "f_panel_off", "f_cb_off", "f_fr_off", "f_u_off",
                    "child_ptr", "child_list", "rel_ptr", "f_rel", "rows_ptr", "f_rows", "sched", "perm", "c_jw", "c_sp_ptr", "c_sp", "c_inv_ptr", "c_inv", "fd", "pc", "pmap")
            st = _lib.FrontPlanStruct(S=P.S, n=P.n, data_size=P.data_size, arena_size=P.arena_size, varena_size=P.varena_size,
                                      **{k: dev[k].data_ptr() for k in keys})
            launches = np.ascontiguousarray(P.launches, dtype=np.int64)
            big = launches[launches[:, 1] == 3]
            self._dev = dict(front=st, keep=dev, launches=launches, max_np=int(big[:, 5].max()) if len(big) else 0)
        return self._dev

    def _offset_of(self, r: np.ndarray, c: np.ndarray) -> np.ndarray:
        """Offset in one item's factor storage of the scalar entries (r, c), ORIGINAL scalar indices, for pairs whose elimination
        positions satisfy pos(r) >= pos(c) after the caller's swap."""
        P = self.plan
        A = P.arrays
        pr, pc = self.to_param[r], self.to_param[c]
        posr, posc = P.pos[pr], P.pos[pc]
        swap = posr < posc
        r2, c2 = np.where(swap, c, r), np.where(swap, r, c)
        pr, pc = self.to_param[r2], self.to_param[c2]
        posr, posc = P.pos[pr], P.pos[pc]
        same = posr == posc
        sw2 = same & (r2 < c2)                       # inside a diagonal block: keep the lower triangle
        r2, c2 = np.where(sw2, c2, r2), np.where(sw2, r2, c2)
        gr = P.pstart[posr] + (r2 - self.param_start[pr])          # permuted scalar indices
        gc = P.pstart[posc] + (c2 - self.param_start[pc])
        s = P.front_of_pos[posc]
        first, w = A["f_first"][s].astype(np.int64), A["f_w"][s].astype(np.int64)
        lcol = gc - first
        out = np.empty(r.shape[0], dtype=np.int64)
        for k in range(r.shape[0]):                  # (host, once per pattern)
            if gr[k] < first[k] + w[k]:
                lrow = gr[k] - first[k]
                if lrow < lcol[k]:                   # both pivots of one front, other orientation: the panel stores the lower triangle
                    lrow, lc = lcol[k], gr[k] - first[k]
                    out[k] = A["f_panel_off"][s[k]] + lrow * w[k] + lc
                    continue
            else:
                rows = P.border_rows[int(s[k])]
                j = int(np.searchsorted(rows, gr[k]))
                assert rows[j] == gr[k], "entry outside the symbolic structure"
                lrow = w[k] + j
            out[k] = A["f_panel_off"][s[k]] + lrow * w[k] + lcol[k]
        return out

    def _add_m_table(self, ptrs: np.ndarray, inds: np.ndarray):
        key = ("M", ptrs.tobytes(), inds.tobytes())
        if key not in self._tables:
            rows = np.repeat(np.arange(ptrs.shape[0] - 1), np.diff(ptrs))
            keep = np.nonzero(inds <= rows)[0]                      # c > r is skipped (baspacho_solver_cuda.cu:37-39)
            off = self._offset_of(rows[keep], inds[keep])
            self._tables[key] = (torch.from_numpy(keep).to(self.device), torch.from_numpy(off).to(self.device))
        return self._tables[key]

    def _add_mtm_table(self, ptrs: np.ndarray, inds: np.ndarray):
        key = ("MtM", ptrs.tobytes(), inds.tobytes())
        if key not in self._tables:
            jj, ii = [], []
            for q in range(ptrs.shape[0] - 1):
                a, b = int(ptrs[q]), int(ptrs[q + 1])
                for j in range(a, b):
                    for i in range(a, j + 1):
                        jj.append(j); ii.append(i)
            jj, ii = np.array(jj, dtype=np.int64), np.array(ii, dtype=np.int64)
            off = self._offset_of(inds[jj], inds[ii])
            order = np.argsort(off, kind="stable")                  # index_add_ then sums each target's contributions in a fixed order
            self._tables[key] = tuple(torch.from_numpy(x[order]).to(self.device) for x in (jj, ii, off))
        return self._tables[key]

    def _diag_offsets(self):
        if "diag" not in self._tables:
            idx = np.arange(self.plan.n)
            self._tables["diag"] = (torch.from_numpy(self._offset_of(idx, idx)).to(self.device),)
        return self._tables["diag"][0]


class NumericDecomposition:
    def __init__(self, symbolic: SymbolicDecomposition, batch_size: int):
        self.sym, self.B = symbolic, batch_size
        P = symbolic.plan
        self.data = torch.zeros(batch_size, P.data_size, dtype=torch.float64, device=symbolic.device)   # baspacho_solver_cuda.cu:13-17
        self._bufs = None

    @staticmethod
    def _np(t):
        return np.ascontiguousarray(t.cpu().numpy() if torch.is_tensor(t) else t, dtype=np.int64)

    def add_M(self, val: torch.Tensor, ptrs, inds):
        keep, off = self.sym._add_m_table(self._np(ptrs), self._np(inds))
        self.data.index_add_(1, off, val.to(self.data.device, torch.float64)[:, keep])

    def add_MtM(self, val: torch.Tensor, ptrs, inds):
        jj, ii, off = self.sym._add_mtm_table(self._np(ptrs), self._np(inds))
        v = val.to(self.data.device, torch.float64)
        self.data.index_add_(1, off, v[:, jj] * v[:, ii])

    def damp(self, alpha: torch.Tensor, beta: torch.Tensor):
        d = self.sym._diag_offsets()
        diag = self.data[:, d]
        self.data[:, d] = diag * (1.0 + alpha.to(diag).view(-1, 1)) + beta.to(diag).view(-1, 1)

    def _buffers(self):
        if self._bufs is None:
            P, dev = self.sym.plan, self.sym._device_plan()
            lib = _lib.load()
            ws_bytes = int(lib.thb_potrf_partial_workspace_bytes(self.B, dev["max_np"])) if dev["max_np"] else 0
            kw = dict(dtype=torch.float64, device=self.data.device)
            self._bufs = dict(arena=torch.empty(2, self.B, P.arena_size, **kw), varena=torch.empty(2, self.B, P.varena_size, **kw),
                              work=torch.empty(self.B, P.n, **kw), ws=torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=self.data.device),
                              info=torch.empty(self.B, dtype=torch.int32, device=self.data.device))
        return self._bufs

    def factor(self):
        dev, bufs, lib = self.sym._device_plan(), self._buffers(), _lib.load()
        L = dev["launches"]
        _lib.check(lib.thb_front_factor_f64(C.byref(dev["front"]), L.ctypes.data, L.shape[0], _lib.ptr(self.data), None, 0, None, None, _lib.ptr(bufs["arena"]),
                                            _lib.ptr(bufs["ws"]) if dev["max_np"] else None, bufs["ws"].numel(), _lib.ptr(bufs["info"]), self.B,
                                            _lib.stream_ptr()), "front_factor")
        bad = bufs["info"].nonzero()
        if bad.numel() > 0:
            k = int(bad[0, 0])
            raise RuntimeError(f"block-sparse Cholesky: batch element {k}: matrix is not positive definite (pivot {int(bufs['info'][k])})")

    def solve(self, x: torch.Tensor):
        """In place, like NumericDecomposition::solve (baspacho_solver.cpp:229-257)."""
        dev, bufs, lib = self.sym._device_plan(), self._buffers(), _lib.load()
        L = dev["launches"]
        rhs = x.detach().to(torch.float64).contiguous().clone()
        out = torch.empty_like(rhs)
        _lib.check(lib.thb_front_solve_f64(C.byref(dev["front"]), L.ctypes.data, L.shape[0], _lib.ptr(self.data), _lib.ptr(rhs), _lib.ptr(out),
                                           _lib.ptr(bufs["work"]), _lib.ptr(bufs["varena"]), self.B, _lib.stream_ptr()), "front_solve")
        x.copy_(out)
        return x
