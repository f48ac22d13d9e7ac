// host_plan.cpp — host-only planning logic of the contraction path (no CUDA calls):
//   * ProcGrid            (reference: src/TiledArray/proc_grid.h:97-260)
//   * CyclicPmap::owner   (reference: src/TiledArray/pmap/cyclic_pmap.h:123-134)
//   * GEMM permutation planning (reference: src/TiledArray/expressions/permopt.h:254-376,
//     binary_engine.h:101-178, cont_engine.h:354-529)
//   * the SUMMA step / tile-pair schedule (reference: dist_eval/contraction_eval.h:655-676,
//     925-1001, 1311-1384)
// These are restated from the reference's behaviour for the new engine's process model (one
// process per GPU, replicated shapes); they run on any host and are covered by CPU tests.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "common.h"
#include "summa_schedule.h"

// ---------------------------------------------------------------------------------------------
// ProcGrid
namespace {

// positive real root of Nn(2x^4 - x^3) + Mm(Px - 2P^2) = 0 by Newton-Raphson from sqrt(P)
int64_t optimal_proc_row(double nprocs, double Mm, double Nn) {
  double x = std::sqrt(nprocs);
  const double PMm = nprocs * Mm, two_P = nprocs + nprocs;
  unsigned it = 0;
  double r = 0.0;
  do {
    const double x2 = x * x, Nx2 = Nn * x2;
    const double f = Nx2 * (2.0 * x2 - x) + PMm * (x - two_P);
    const double df = Nx2 * (8.0 * x - 3.0) + PMm;
    const double xn = x - f / df;
    r = std::abs(xn - x);
    x = xn;
  } while (r > 0.1 && ++it < 21u);
  return (int64_t)(x + 0.5);
}

void minimize_unused_procs(int64_t& x, int64_t& y, int64_t nprocs, int64_t min_x, int64_t max_x) {
  int64_t unused = x * y;  // sic: the reference seeds the search with x*y
  if (unused == 0) return;
  const int64_t delta = std::max<int64_t>(1, (int64_t)std::log2((double)nprocs));
  const int64_t optimal_x = x;
  int64_t diff = 0;
  const int64_t min_test_x = std::max<int64_t>(min_x, x - delta);
  for (int64_t test_x = std::min(x + delta, max_x); test_x >= min_test_x; --test_x) {
    const int64_t test_y = nprocs / test_x;
    const int64_t test_unused = nprocs - test_x * test_y;
    const int64_t test_diff = std::llabs(optimal_x - test_x);
    if (test_unused < unused || (test_unused == unused && test_diff < diff)) {
      x = test_x; y = test_y; unused = test_unused; diff = test_diff;
    }
  }
}

}  // namespace

extern "C" int tadev_proc_grid_make(int rank, int nprocs, int64_t rows, int64_t cols, int64_t row_size,
                                    int64_t col_size, tadev_proc_grid* out) {
  TADEV_REQUIRE(out, "tadev_proc_grid_make: null out");
  TADEV_REQUIRE(nprocs >= 1 && rank >= 0, "tadev_proc_grid_make: bad rank/nprocs");
  TADEV_REQUIRE(rows >= 1 && cols >= 1, "tadev_proc_grid_make: empty tile grid");
  const int64_t size = rows * cols;
  int64_t pr, pc, ps;
  tadev_proc_grid g{};
  g.rank_row = g.rank_col = -1;
  if (nprocs == 1) {
    pr = pc = ps = 1;
    if (rank < ps) { g.rank_row = g.rank_col = 0; g.local_rows = rows; g.local_cols = cols; g.local_size = size; }
  } else if (size <= nprocs) {
    pr = rows; pc = cols; ps = size;
    if (rank < ps) {
      g.rank_row = (int32_t)(rank / pc); g.rank_col = (int32_t)(rank % pc);
      g.local_rows = g.local_cols = g.local_size = 1;
    }
  } else {
    const int64_t min_pr = std::max<int64_t>((nprocs + cols - 1) / cols, 1);
    const int64_t max_pr = std::min<int64_t>(nprocs, rows);
    pr = std::max<int64_t>(min_pr, std::min<int64_t>(optimal_proc_row((double)nprocs, (double)row_size, (double)col_size), max_pr));
    pc = nprocs / pr;
    if (pr > min_pr && pr < max_pr) minimize_unused_procs(pr, pc, nprocs, min_pr, max_pr);
    ps = pr * pc;
    if (rank < ps) {
      g.rank_row = (int32_t)(rank / pc); g.rank_col = (int32_t)(rank % pc);
      g.local_rows = rows / pr + (g.rank_row < rows % pr ? 1 : 0);
      g.local_cols = cols / pc + (g.rank_col < cols % pc ? 1 : 0);
      g.local_size = g.local_rows * g.local_cols;
    }
  }
  g.proc_rows = (int32_t)pr; g.proc_cols = (int32_t)pc; g.proc_size = (int32_t)ps;
  *out = g;
  return TADEV_OK;
}

extern "C" int tadev_cyclic_owner(int64_t tile, int64_t cols, int proc_rows, int proc_cols, int* owner) {
  TADEV_REQUIRE(owner && cols >= 1 && proc_rows >= 1 && proc_cols >= 1 && tile >= 0, "tadev_cyclic_owner: bad args");
  const int64_t tr = tile / cols, tc = tile % cols;
  *owner = (int)((tr % proc_rows) * proc_cols + (tc % proc_cols));
  return TADEV_OK;
}

// ---------------------------------------------------------------------------------------------
// contraction planning
namespace {

std::vector<std::string> split_indices(const char* s) {
  std::vector<std::string> out;
  std::string cur;
  for (const char* p = s; *p; ++p) {
    if (*p == ',') { out.push_back(cur); cur.clear(); }
    else if (*p != ' ' && *p != '\t') cur.push_back(*p);
  }
  if (!cur.empty() || !out.empty()) out.push_back(cur);
  return out;
}
std::string join_indices(const std::vector<std::string>& v) {
  std::string s;
  for (size_t i = 0; i < v.size(); ++i) { if (i) s += ","; s += v[i]; }
  return s;
}
size_t find_idx(const std::vector<std::string>& v, const std::string& x) {
  return std::find(v.begin(), v.end(), x) - v.begin();
}
bool all_unique(const std::vector<std::string>& v) {
  for (size_t i = 0; i < v.size(); ++i) {
    if (v[i].empty()) return false;
    for (size_t j = i + 1; j < v.size(); ++j) if (v[i] == v[j]) return false;
  }
  return true;
}

}  // namespace

extern "C" int tadev_plan_contraction(const char* target, const char* left, const char* right,
                                      tadev_contraction_plan* out) {
  TADEV_REQUIRE(target && left && right && out, "tadev_plan_contraction: null");
  const auto L = split_indices(left), Rr = split_indices(right), T = split_indices(target);
  TADEV_REQUIRE(all_unique(L) && all_unique(Rr) && (T.empty() || all_unique(T)), "tadev_plan_contraction: repeated or empty index");
  TADEV_REQUIRE(L.size() <= 16 && Rr.size() <= 16 && T.size() <= 16, "tadev_plan_contraction: rank > 16");
  const unsigned left_rank = (unsigned)L.size(), right_rank = (unsigned)Rr.size();
  std::vector<std::string> tl, tr, res;
  for (unsigned i = 0; i < left_rank; ++i) {
    if (find_idx(Rr, L[i]) == right_rank) { tl.push_back(L[i]); res.push_back(L[i]); }
    else tr.push_back(L[i]);
  }
  const unsigned inner_rank = (unsigned)tr.size(), left_outer_rank = (unsigned)tl.size();
  const unsigned right_outer_rank = right_rank - inner_rank;
  int lt = 3, rt = 3;  // general
  if (inner_rank == 0) {
    for (unsigned i = 0; i < right_rank; ++i) { tr.push_back(Rr[i]); res.push_back(Rr[i]); }
  } else {
    bool ordered = true, l_nt = true, l_t = true, r_nt = true, r_t = true;
    const bool perm_left = (left_rank < right_rank) || (left_rank == right_rank);  // prefer_to_permute_left
    for (unsigned i = 0; i < right_rank; ++i) {
      const std::string& idx = Rr[i];
      const unsigned j = (unsigned)find_idx(L, idx);
      if (j == left_rank) { tr.push_back(idx); res.push_back(idx); }
      else {
        const unsigned x = (unsigned)tl.size() - left_outer_rank;
        ordered = ordered && (tr[x] == idx);
        l_nt = l_nt && (j >= left_outer_rank);
        l_t = l_t && (j < inner_rank);
        r_nt = r_nt && (i < inner_rank);
        r_t = r_t && (i >= right_outer_rank);
        if (ordered) tl.push_back(idx);
        else if (perm_left) { tl.push_back(idx); tr[x] = idx; l_nt = l_t = false; }
        else { tl.push_back(tr[x]); r_nt = r_t = false; }
      }
    }
    lt = l_nt ? 1 : (l_t ? 2 : 3);
    rt = r_nt ? 1 : (r_t ? 2 : 3);
  }
  // every target index must be a result index and vice versa
  if (!T.empty()) {
    TADEV_REQUIRE(T.size() == res.size(), "tadev_plan_contraction: target rank %zu != result rank %zu", T.size(), res.size());
    for (auto& x : T) TADEV_REQUIRE(find_idx(res, x) < res.size(), "tadev_plan_contraction: target index '%s' is not a free index", x.c_str());
  }
  tadev_contraction_plan P{};
  P.left_rank = (int32_t)left_rank; P.right_rank = (int32_t)right_rank;
  P.result_rank = (int32_t)res.size(); P.inner_rank = (int32_t)inner_rank;
  P.left_permtype = lt; P.right_permtype = rt;
  P.opA = lt == 2 ? TADEV_OP_T : TADEV_OP_N;
  P.opB = rt == 2 ? TADEV_OP_T : TADEV_OP_N;
  for (int i = 0; i < 16; ++i) P.perm_left[i] = P.perm_right[i] = P.perm_result[i] = -1;
  auto make_perm = [](const std::vector<std::string>& from, const std::vector<std::string>& to, int32_t* perm) {
    bool ident = true;
    for (size_t i = 0; i < from.size(); ++i) if (from[i] != to[i]) ident = false;
    if (ident) return;
    for (size_t i = 0; i < from.size(); ++i) perm[i] = (int32_t)find_idx(to, from[i]);  // image form
  };
  if (lt == 3) make_perm(L, tl, P.perm_left);
  if (rt == 3) make_perm(Rr, tr, P.perm_right);
  if (!T.empty()) make_perm(res, T, P.perm_result);
  snprintf(P.left_target, sizeof(P.left_target), "%s", join_indices(tl).c_str());
  snprintf(P.right_target, sizeof(P.right_target), "%s", join_indices(tr).c_str());
  snprintf(P.result_gemm, sizeof(P.result_gemm), "%s", join_indices(res).c_str());
  *out = P;
  return TADEV_OK;
}

// The reference always produces the result in (left outer, right outer) order and permutes every
// result tile when the target order differs (contract_reduce.h:370-378). C = A*B and C^T = B^T*A^T
// are the same set of products summed in the same k order, so the engine may exchange the
// operands when that removes explicit tile permutations (BASELINE config 4:
// R("a,b,i,j") = T("c,d,i,j") * V("a,b,c,d") is a plain NN product V*T with no permutation at all,
// but two transposed operands plus a result permutation as written).
extern "C" int tadev_plan_contraction_opt(const char* target, const char* left, const char* right,
                                          tadev_contraction_plan* out, int32_t* swapped) {
  TADEV_REQUIRE(out && swapped, "tadev_plan_contraction_opt: null");
  tadev_contraction_plan a, b;
  int rc = tadev_plan_contraction(target, left, right, &a);
  if (rc) return rc;
  *out = a;
  *swapped = 0;
  auto cost = [](const tadev_contraction_plan& p) {
    return (p.perm_left[0] >= 0) + (p.perm_right[0] >= 0) + (p.perm_result[0] >= 0);
  };
  if (cost(a) == 0 || a.inner_rank == 0) return TADEV_OK;
  if (tadev_plan_contraction(target, right, left, &b) != TADEV_OK) return TADEV_OK;
  if (cost(b) < cost(a)) { *out = b; *swapped = 1; }
  return TADEV_OK;
}

// ---------------------------------------------------------------------------------------------
// SUMMA schedule
SummaSchedule make_summa_schedule(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a,
                                  const float* b, const float* cn, float thr) {
  SummaSchedule S;
  auto a_nz = [&](int i, int k) { return !a || a[(size_t)i * Kt + k] >= thr; };
  auto b_nz = [&](int k, int j) { return !b || b[(size_t)k * Nt + j] >= thr; };
  auto c_nz = [&](int i, int j) { return !cn || cn[(size_t)i * Nt + j] >= thr; };
  for (int k = 0; k < Kt; ++k) {
    SummaStep st;
    st.k = k;
    // this rank's column-of-A panel (rows i == r mod Pr) and row-of-B panel (cols j == c mod Pc)
    for (int i = r; i < Mt; i += Pr) if (a_nz(i, k)) st.a_rows.push_back(i);
    for (int j = c; j < Nt; j += Pc) if (b_nz(k, j)) st.b_cols.push_back(j);
    // group-consistent broadcast decisions (every member of the group evaluates the same
    // predicate from the replicated shapes): a panel travels iff it is non-empty and the
    // opposite operand has any non-zero tile in this k (otherwise no rank can use it).
    bool any_a = false, any_b = false;
    for (int i = 0; i < Mt && !any_a; ++i) any_a = a_nz(i, k);
    for (int j = 0; j < Nt && !any_b; ++j) any_b = b_nz(k, j);
    st.bcast_a = Pc > 1 && !st.a_rows.empty() && any_b;
    st.bcast_b = Pr > 1 && !st.b_cols.empty() && any_a;
    // local contraction happens iff both of my panels are non-empty (iterate_sparse, :974-1001)
    st.compute = !st.a_rows.empty() && !st.b_cols.empty();
    if (st.compute) {
      st.pair_begin = (int64_t)S.pair_i.size();
      for (int i : st.a_rows)
        for (int j : st.b_cols)
          if (c_nz(i, j)) { S.pair_i.push_back(i); S.pair_j.push_back(j); }
      st.pair_end = (int64_t)S.pair_i.size();
    } else {
      st.pair_begin = st.pair_end = (int64_t)S.pair_i.size();
    }
    if (st.compute || st.bcast_a || st.bcast_b) S.steps.push_back(std::move(st));
    else ++S.nskipped;
  }
  return S;
}

extern "C" int tadev_summa_schedule(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a_norms,
                                    const float* b_norms, const float* c_norms, float threshold, int32_t* step_k,
                                    int32_t* step_pair_begin, int32_t* nsteps_out, int32_t* pair_i, int32_t* pair_j,
                                    int64_t pair_capacity, int64_t* npairs_out) {
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && r >= 0 && r < Pr && c >= 0 && c < Pc, "tadev_summa_schedule: bad grid position");
  TADEV_REQUIRE(Mt >= 0 && Nt >= 0 && Kt >= 0, "tadev_summa_schedule: negative extents");
  TADEV_REQUIRE(nsteps_out && npairs_out, "tadev_summa_schedule: null outputs");
  SummaSchedule S = make_summa_schedule(Pr, Pc, r, c, Mt, Nt, Kt, a_norms, b_norms, c_norms, threshold);
  int n = 0;
  for (auto& st : S.steps) {
    if (!st.compute) continue;  // report only steps with local contractions
    if (step_k) step_k[n] = st.k;
    if (step_pair_begin) step_pair_begin[n] = (int32_t)st.pair_begin;
    ++n;
  }
  if (step_pair_begin) step_pair_begin[n] = (int32_t)S.pair_i.size();
  *nsteps_out = n;
  *npairs_out = (int64_t)S.pair_i.size();
  if (pair_i && pair_j) {
    TADEV_REQUIRE((int64_t)S.pair_i.size() <= pair_capacity, "tadev_summa_schedule: pair capacity %lld < %zu",
                  (long long)pair_capacity, S.pair_i.size());
    std::copy(S.pair_i.begin(), S.pair_i.end(), pair_i);
    std::copy(S.pair_j.begin(), S.pair_j.end(), pair_j);
  }
  return TADEV_OK;
}

extern "C" int tadev_summa_steps(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a_norms,
                                 const float* b_norms, const float* c_norms, float threshold, int32_t* step_k,
                                 int32_t* step_flags, int32_t* a_begin, int32_t* a_rows, int32_t* b_begin,
                                 int32_t* b_cols, int32_t* nsteps_out) {
  TADEV_REQUIRE(Pr >= 1 && Pc >= 1 && r >= 0 && r < Pr && c >= 0 && c < Pc, "tadev_summa_steps: bad grid position");
  TADEV_REQUIRE(Mt >= 0 && Nt >= 0 && Kt >= 0, "tadev_summa_steps: negative extents");
  TADEV_REQUIRE(step_k && step_flags && a_begin && a_rows && b_begin && b_cols && nsteps_out, "tadev_summa_steps: null outputs");
  SummaSchedule S = make_summa_schedule(Pr, Pc, r, c, Mt, Nt, Kt, a_norms, b_norms, c_norms, threshold);
  int n = 0, na = 0, nb = 0;
  for (auto& st : S.steps) {
    step_k[n] = st.k;
    step_flags[n] = (st.compute ? 1 : 0) | (st.bcast_a ? 2 : 0) | (st.bcast_b ? 4 : 0);
    a_begin[n] = na;
    b_begin[n] = nb;
    for (int i : st.a_rows) a_rows[na++] = i;
    for (int j : st.b_cols) b_cols[nb++] = j;
    ++n;
  }
  a_begin[n] = na;
  b_begin[n] = nb;
  *nsteps_out = n;
  return TADEV_OK;
}
