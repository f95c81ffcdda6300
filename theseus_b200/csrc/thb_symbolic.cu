// Symbolic analysis of the batched block-sparse Cholesky (host code, batch independent, once per structure).
//
// Replaces what the reference delegates to compiled third-party code: BaSpaCho's createSolver behind
// SymbolicDecomposition(param_size, sparse_struct_ptrs, sparse_struct_inds, device) (theseus/extlib/baspacho_solver.cpp:259-319),
// cusolverSpXcsrsymamdHost + csrluAnalysisHost (extlib/cusolver_lu_solver.cpp:95-196) and CHOLMOD's analyze_AAt
// (optimizer/linear/cholmod_sparse_solver.py:38-54).  Input = exactly what the reference hands over: param_size [N] and the CSR
// (ptrs, inds) of the symmetric block pattern of AtA (optimizer/linear/baspacho_sparse_solver.py:93-113).
//
// Steps: (1) greedy minimum (weighted external) degree ordering with a deterministic tie-break, (2) symbolic factorisation by
// elimination -> column structures + elimination tree, (3) tree levels, (4) factor layout (per column: diagonal block, then the
// sub-diagonal blocks, row-major), (5) left-looking update-pair lists per block, (6) per-level work lists for both numeric back
// ends (thb_sparse.cu: u_/f_/t_/s_ arrays; thb_sparse_lane.cu: ln_* arrays + launch list).  The result is a bag of named host
// arrays; theseus_b200/sparse.py uploads them and fills thb_sparse_plan / thb_sparse_lane_plan.  sparse.py:analyze_py is the same
// algorithm in Python (the executable specification; tests/test_sparse_symbolic.py checks the two agree array by array).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <queue>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/thb200.h"

namespace {

struct Arr {
  std::vector<char> bytes;
  int elem = 8;
  int64_t count() const { return (int64_t)bytes.size() / elem; }
};

template <typename T>
Arr make_arr(const std::vector<T>& v) {
  Arr a;
  a.elem = (int)sizeof(T);
  a.bytes.resize(v.size() * sizeof(T));
  if (!v.empty()) std::memcpy(a.bytes.data(), v.data(), a.bytes.size());
  return a;
}

}  // namespace

struct thb_symbolic {
  std::map<std::string, Arr> arrays;
  std::map<std::string, double> stats;
};

namespace {

constexpr int kLaneHeavy = 8;  // == sparse.py:LANE_HEAVY

std::vector<int64_t> min_degree_order(int64_t N, const int64_t* ptrs, const int64_t* inds, const int64_t* w) {
  std::vector<std::set<int>> adj(N);
  for (int64_t i = 0; i < N; i++)
    for (int64_t q = ptrs[i]; q < ptrs[i + 1]; q++)
      if (inds[q] != i) adj[i].insert((int)inds[q]);
  std::vector<int64_t> deg(N);
  typedef std::pair<int64_t, int> Key;
  std::priority_queue<Key, std::vector<Key>, std::greater<Key>> heap;
  for (int64_t i = 0; i < N; i++) {
    int64_t d = 0;
    for (int a : adj[i]) d += w[a];
    deg[i] = d;
    heap.push(Key(d, (int)i));
  }
  std::vector<char> done(N, 0);
  std::vector<int64_t> order;
  order.reserve(N);
  while (!heap.empty()) {
    const Key k = heap.top();
    heap.pop();
    const int v = k.second;
    if (done[v] || k.first != deg[v]) continue;
    done[v] = 1;
    order.push_back(v);
    const std::vector<int> nb(adj[v].begin(), adj[v].end());  // sorted
    for (int a : nb) adj[a].erase(v);
    for (size_t ai = 0; ai < nb.size(); ai++)  // clique among the neighbours
      for (size_t bi = ai + 1; bi < nb.size(); bi++)
        if (adj[nb[ai]].insert(nb[bi]).second) adj[nb[bi]].insert(nb[ai]);
    for (int a : nb) {
      int64_t nd = 0;
      for (int x : adj[a]) nd += w[x];
      deg[a] = nd;
      heap.push(Key(nd, a));
    }
    adj[v].clear();
  }
  return order;
}

}  // namespace

extern "C" {

int thb_symbolic_create(const int64_t* param_size, int64_t N, const int64_t* blk_ptrs, const int64_t* blk_inds, int32_t ordering,
                        thb_symbolic** out) {
  if (out == nullptr || N < 0 || (N > 0 && (param_size == nullptr || blk_ptrs == nullptr || blk_inds == nullptr))) return THB_ERR_BAD_ARG;
  if (ordering != 0 && ordering != 1) return THB_ERR_BAD_ARG;
  typedef int64_t i64;
  typedef int32_t i32;
  typedef int16_t i16;
  std::vector<i64> order;
  if (ordering == 0) {
    order = min_degree_order(N, blk_ptrs, blk_inds, param_size);
  } else {
    order.resize(N);
    for (i64 i = 0; i < N; i++) order[i] = i;
  }
  std::vector<i64> pos(N), dims(N), orig_start(N), col_start(N), pstart(N);
  for (i64 k = 0; k < N; k++) pos[order[k]] = k;
  i64 n = 0;
  for (i64 v = 0; v < N; v++) { orig_start[v] = n; n += param_size[v]; }
  {
    i64 acc = 0;
    for (i64 k = 0; k < N; k++) { dims[k] = param_size[order[k]]; col_start[k] = orig_start[order[k]]; pstart[k] = acc; acc += dims[k]; }
  }
  // ---- symbolic factorisation in elimination order ----
  std::vector<std::set<i64>> adj(N);
  for (i64 v = 0; v < N; v++)
    for (i64 q = blk_ptrs[v]; q < blk_ptrs[v + 1]; q++) {
      const i64 pu = pos[blk_inds[q]];
      if (pu != pos[v]) adj[pos[v]].insert(pu);
    }
  std::vector<std::vector<i64>> st(N);
  std::vector<i64> parent(N, -1), level(N, 0);
  for (i64 j = 0; j < N; j++) {
    for (i64 x : adj[j]) if (x > j) st[j].push_back(x);  // std::set iterates in increasing order
    if (!st[j].empty()) {
      const i64 p = st[j][0];
      parent[j] = p;
      for (i64 x : st[j]) if (x != p) adj[p].insert(x);
    }
  }
  for (i64 j = 0; j < N; j++)
    if (parent[j] >= 0) level[parent[j]] = std::max(level[parent[j]], level[j] + 1);
  i64 nlev = 0;
  for (i64 j = 0; j < N; j++) nlev = std::max(nlev, level[j] + 1);
  // ---- factor layout ----
  std::vector<i64> col_blk0(N + 1, 0);  // first block id of each column
  std::vector<i64> blk_off, blk_i, blk_j;
  std::vector<i32> blk_rows, blk_cols;
  i64 off = 0;
  for (i64 j = 0; j < N; j++) {
    col_blk0[j] = (i64)blk_off.size();
    for (size_t q = 0; q <= st[j].size(); q++) {
      const i64 i = q == 0 ? j : st[j][q - 1];
      blk_off.push_back(off); blk_i.push_back(i); blk_j.push_back(j);
      blk_rows.push_back((i32)dims[i]); blk_cols.push_back((i32)dims[j]);
      off += dims[i] * dims[j];
    }
  }
  col_blk0[N] = (i64)blk_off.size();
  const i64 data_size = off, nblk = (i64)blk_off.size();
  auto blk_id = [&](i64 i, i64 j) -> i64 {  // block (i,j), i == j or i in st[j]
    if (i == j) return col_blk0[j];
    const auto it = std::lower_bound(st[j].begin(), st[j].end(), i);
    return col_blk0[j] + 1 + (it - st[j].begin());
  };
  std::vector<i64> winv_off(N), diag_off(N);
  i64 winv_size = 0;
  for (i64 j = 0; j < N; j++) { winv_off[j] = winv_size; winv_size += dims[j] * dims[j]; diag_off[j] = blk_off[col_blk0[j]]; }
  // ---- left-looking update lists ----
  std::vector<std::vector<std::pair<i64, i64>>> upd(nblk);
  double flops = 0;
  for (i64 k = 0; k < N; k++) {
    const std::vector<i64>& s = st[k];
    const i64 dk = dims[k];
    flops += (double)(dk * dk * dk / 3);
    for (size_t bi = 0; bi < s.size(); bi++) {
      flops += (double)(dims[s[bi]] * dk * dk);
      const i64 idb = col_blk0[k] + 1 + (i64)bi;
      for (size_t ai = bi; ai < s.size(); ai++) {
        upd[blk_id(s[ai], s[bi])].push_back(std::make_pair(col_blk0[k] + 1 + (i64)ai, idb));
        flops += (double)(2 * dims[s[ai]] * dims[s[bi]] * dk);
      }
    }
  }
  std::vector<i64> up_ptr(nblk + 1, 0);
  for (i64 t = 0; t < nblk; t++) up_ptr[t + 1] = up_ptr[t] + (i64)upd[t].size();
  std::vector<i64> up_a(up_ptr[nblk]), up_b(up_ptr[nblk]);
  std::vector<i32> up_k(up_ptr[nblk]);
  for (i64 t = 0; t < nblk; t++)
    for (size_t q = 0; q < upd[t].size(); q++) {
      up_a[up_ptr[t] + q] = blk_off[upd[t][q].first];
      up_b[up_ptr[t] + q] = blk_off[upd[t][q].second];
      up_k[up_ptr[t] + q] = blk_cols[upd[t][q].first];
    }
  // ---- per-level work items ----
  std::vector<std::vector<i64>> cols_by_level(nlev);
  for (i64 j = 0; j < N; j++) cols_by_level[level[j]].push_back(j);
  std::vector<i64> u_ptr(1, 0), u_tgt, u_p0, u_p1, f_ptr(1, 0), f_off, f_w, t_ptr(1, 0), t_off, t_w, s_ptr(1, 0);
  std::vector<i16> u_r, u_c, u_ld, t_r, t_dim;
  std::vector<i32> f_dim, f_col, s_col;
  // lane lists
  std::vector<i64> ln_u_tgt, ln_u_p0, ln_u_p1, ln_t_off, ln_t_diag, ln_t_dl;
  std::vector<i32> ln_t_pstart, ln_s_col, launches;
  for (i64 lv = 0; lv < nlev; lv++) {
    std::map<std::vector<i64>, std::vector<i64>> ucls;   // (heavy, di, dj) -> blocks
    std::map<std::pair<i64, i64>, std::vector<i64>> tcls;  // (di, dj) -> blocks
    std::map<i64, std::vector<i64>> scls;
    for (i64 j : cols_by_level[lv]) {
      const i64 dj = dims[j];
      scls[dj].push_back(j);
      for (i64 t = col_blk0[j]; t < col_blk0[j + 1]; t++) {
        const i64 i = blk_i[t], di = dims[i];
        const i64 np = up_ptr[t + 1] - up_ptr[t];
        if (np > 0) {
          u_tgt.push_back(blk_off[t]); u_r.push_back((i16)di); u_c.push_back((i16)dj); u_ld.push_back(i == j ? 1 : 0);
          u_p0.push_back(up_ptr[t]); u_p1.push_back(up_ptr[t + 1]);
          ucls[{np >= kLaneHeavy ? 1 : 0, di, dj}].push_back(t);
        }
        if (i != j)
          for (i64 r = 0; r < di; r++) { t_off.push_back(blk_off[t]); t_r.push_back((i16)r); t_dim.push_back((i16)dj); t_w.push_back(winv_off[j]); }
        tcls[std::make_pair(di, dj)].push_back(t);
      }
      f_off.push_back(diag_off[j]); f_dim.push_back((i32)dj); f_w.push_back(winv_off[j]); f_col.push_back((i32)j);
      s_col.push_back((i32)j);
    }
    u_ptr.push_back((i64)u_tgt.size()); f_ptr.push_back((i64)f_off.size()); t_ptr.push_back((i64)t_off.size()); s_ptr.push_back((i64)s_col.size());
    for (const auto& kv : ucls) {
      const i32 b0 = (i32)ln_u_tgt.size();
      for (i64 t : kv.second) { ln_u_tgt.push_back(blk_off[t]); ln_u_p0.push_back(up_ptr[t]); ln_u_p1.push_back(up_ptr[t + 1]); }
      const i32 row[5] = {kv.first[0] ? THB_LANE_UH : THB_LANE_U, (i32)kv.first[1], (i32)kv.first[2], b0, (i32)ln_u_tgt.size()};
      launches.insert(launches.end(), row, row + 5);
    }
    for (const auto& kv : tcls) {
      const i32 b0 = (i32)ln_t_off.size();
      for (i64 t : kv.second) {
        const i64 j = blk_j[t];
        ln_t_off.push_back(blk_off[t]); ln_t_diag.push_back(diag_off[j]); ln_t_dl.push_back(winv_off[j]); ln_t_pstart.push_back((i32)pstart[j]);
      }
      const i32 row[5] = {THB_LANE_T, (i32)kv.first.first, (i32)kv.first.second, b0, (i32)ln_t_off.size()};
      launches.insert(launches.end(), row, row + 5);
    }
    for (const auto& kv : scls) {
      const i32 b0 = (i32)ln_s_col.size();
      for (i64 j : kv.second) ln_s_col.push_back((i32)j);
      const i32 row[5] = {THB_LANE_S, (i32)kv.first, (i32)kv.first, b0, (i32)ln_s_col.size()};
      launches.insert(launches.end(), row, row + 5);
    }
  }
  // ---- substitution lists ----
  std::vector<std::vector<std::pair<i64, i64>>> row_lists(N);  // row j: (offset of L_jk, k)
  for (i64 k = 0; k < N; k++)
    for (size_t q = 0; q < st[k].size(); q++) row_lists[st[k][q]].push_back(std::make_pair(blk_off[col_blk0[k] + 1 + (i64)q], k));
  std::vector<i64> fr_ptr(1, 0), fr_off, bc_ptr(1, 0), bc_off;
  std::vector<i32> fr_k, bc_i, fr_p, fr_d, bc_p, bc_d;
  for (i64 j = 0; j < N; j++) {
    for (const auto& e : row_lists[j]) { fr_off.push_back(e.first); fr_k.push_back((i32)e.second); fr_p.push_back((i32)pstart[e.second]); fr_d.push_back((i32)dims[e.second]); }
    fr_ptr.push_back((i64)fr_off.size());
    for (size_t q = 0; q < st[j].size(); q++) {
      const i64 i = st[j][q];
      bc_off.push_back(blk_off[col_blk0[j] + 1 + (i64)q]); bc_i.push_back((i32)i); bc_p.push_back((i32)pstart[i]); bc_d.push_back((i32)dims[i]);
    }
    bc_ptr.push_back((i64)bc_off.size());
  }
  std::vector<i64> struct_ptr(1, 0), struct_idx;
  i64 max_front = 0;
  for (i64 j = 0; j < N; j++) {
    struct_idx.insert(struct_idx.end(), st[j].begin(), st[j].end());
    struct_ptr.push_back((i64)struct_idx.size());
    max_front = std::max(max_front, (i64)st[j].size() + 1);
  }
  // ---- chains (fundamental supernodes): j, j+1, ... with parent(j) = j+1, |struct(j)| = |struct(j+1)| + 1, j the only child ----
  std::vector<i64> nchild(N, 0), chain_of(N, -1), chain_level;
  for (i64 j = 0; j < N; j++) if (parent[j] >= 0) nchild[parent[j]]++;
  i64 nchains = 0;
  for (i64 j = 0; j < N; j++) {
    if (chain_of[j] >= 0) continue;
    chain_of[j] = nchains;
    i64 k = j;
    while (true) {
      const i64 p = parent[k];
      if (p != k + 1 || nchild[p] != 1 || st[k].size() != st[p].size() + 1) break;
      chain_of[p] = nchains;
      k = p;
    }
    nchains++;
  }
  chain_level.assign(nchains, 0);
  for (i64 j = 0; j < N; j++) {
    const i64 p = parent[j];
    if (p >= 0 && chain_of[p] != chain_of[j]) chain_level[chain_of[p]] = std::max(chain_level[chain_of[p]], chain_level[chain_of[j]] + 1);
  }
  i64 chain_levels = 0;
  for (i64 c = 0; c < nchains; c++) chain_levels = std::max(chain_levels, chain_level[c] + 1);
  auto to32 = [](const std::vector<i64>& v) { return std::vector<i32>(v.begin(), v.end()); };

  thb_symbolic* S = new thb_symbolic();
  auto& A = S->arrays;
  A["order"] = make_arr(order); A["pos"] = make_arr(pos); A["level"] = make_arr(level);
  A["dims64"] = make_arr(dims); A["col_start64"] = make_arr(col_start); A["pstart64"] = make_arr(pstart);
  A["struct_ptr"] = make_arr(struct_ptr); A["struct_idx"] = make_arr(struct_idx);
  A["chain_of"] = make_arr(chain_of); A["chain_level"] = make_arr(chain_level);
  A["blk_off"] = make_arr(blk_off); A["blk_i"] = make_arr(blk_i); A["blk_j"] = make_arr(blk_j);
  A["blk_rows"] = make_arr(blk_rows); A["blk_cols"] = make_arr(blk_cols); A["up_ptr"] = make_arr(up_ptr);
  // thb_sparse_plan arrays (names == struct fields)
  A["dims"] = make_arr(to32(dims)); A["col_start"] = make_arr(to32(col_start)); A["pstart"] = make_arr(to32(pstart));
  A["winv_off"] = make_arr(winv_off); A["diag_off"] = make_arr(diag_off);
  A["up_a"] = make_arr(up_a); A["up_b"] = make_arr(up_b); A["up_k"] = make_arr(up_k);
  A["u_ptr"] = make_arr(u_ptr); A["u_tgt"] = make_arr(u_tgt); A["u_r"] = make_arr(u_r); A["u_c"] = make_arr(u_c); A["u_ld"] = make_arr(u_ld);
  A["u_p0"] = make_arr(u_p0); A["u_p1"] = make_arr(u_p1);
  A["f_ptr"] = make_arr(f_ptr); A["f_off"] = make_arr(f_off); A["f_dim"] = make_arr(f_dim); A["f_w"] = make_arr(f_w); A["f_col"] = make_arr(f_col);
  A["t_ptr"] = make_arr(t_ptr); A["t_off"] = make_arr(t_off); A["t_r"] = make_arr(t_r); A["t_dim"] = make_arr(t_dim); A["t_w"] = make_arr(t_w);
  A["s_ptr"] = make_arr(s_ptr); A["s_col"] = make_arr(s_col);
  A["fr_ptr"] = make_arr(fr_ptr); A["fr_off"] = make_arr(fr_off); A["fr_k"] = make_arr(fr_k);
  A["bc_ptr"] = make_arr(bc_ptr); A["bc_off"] = make_arr(bc_off); A["bc_i"] = make_arr(bc_i);
  // thb_sparse_lane_plan arrays
  A["ln_u_tgt"] = make_arr(ln_u_tgt); A["ln_u_p0"] = make_arr(ln_u_p0); A["ln_u_p1"] = make_arr(ln_u_p1);
  A["ln_t_off"] = make_arr(ln_t_off); A["ln_t_diag"] = make_arr(ln_t_diag); A["ln_t_dl"] = make_arr(ln_t_dl); A["ln_t_pstart"] = make_arr(ln_t_pstart);
  A["ln_s_col"] = make_arr(ln_s_col); A["ln_launches"] = make_arr(launches);
  A["ln_fr_p"] = make_arr(fr_p); A["ln_fr_d"] = make_arr(fr_d); A["ln_bc_p"] = make_arr(bc_p); A["ln_bc_d"] = make_arr(bc_d);
  S->stats["N"] = (double)N; S->stats["n"] = (double)n; S->stats["data_size"] = (double)data_size; S->stats["winv_size"] = (double)winv_size;
  S->stats["nnz_L"] = (double)data_size; S->stats["flops"] = flops; S->stats["levels"] = (double)nlev;
  S->stats["max_front"] = (double)max_front; S->stats["num_updates"] = (double)up_ptr[nblk];
  S->stats["num_chains"] = (double)nchains; S->stats["chain_levels"] = (double)chain_levels;
  *out = S;
  return THB_OK;
}

void thb_symbolic_destroy(thb_symbolic* s) { delete s; }

int64_t thb_symbolic_array_count(const thb_symbolic* s, const char* name) {
  if (s == nullptr || name == nullptr) return -1;
  const auto it = s->arrays.find(name);
  return it == s->arrays.end() ? -1 : it->second.count();
}

int32_t thb_symbolic_array_elem_bytes(const thb_symbolic* s, const char* name) {
  if (s == nullptr || name == nullptr) return -1;
  const auto it = s->arrays.find(name);
  return it == s->arrays.end() ? -1 : it->second.elem;
}

int thb_symbolic_array_copy(const thb_symbolic* s, const char* name, void* dst, int64_t dst_bytes) {
  if (s == nullptr || name == nullptr) return THB_ERR_BAD_ARG;
  const auto it = s->arrays.find(name);
  if (it == s->arrays.end() || dst_bytes != (int64_t)it->second.bytes.size()) return THB_ERR_BAD_ARG;
  if (dst_bytes > 0) {
    if (dst == nullptr) return THB_ERR_BAD_ARG;
    std::memcpy(dst, it->second.bytes.data(), (size_t)dst_bytes);
  }
  return THB_OK;
}

double thb_symbolic_stat(const thb_symbolic* s, const char* name) {
  if (s == nullptr || name == nullptr) return -1.0;
  const auto it = s->stats.find(name);
  return it == s->stats.end() ? -1.0 : it->second;
}

}  // extern "C"
