// rls_group.cu — single-process multi-device: one call shards a system over the GPUs of a box.
//
// "multi-GPU handled inside one call (single process, ncclGroupStart/End)".  A Julia user calls
// createLinearSolver(FISTA, A; ...) / solve!(solver, b) from ONE process; the shim hands the whole host matrix (or a
// Philox recipe) to rls_group_mat_create, which row-partitions it over the devices of the group (contiguous row blocks,
// the same rule as the process-per-GPU path), and rls_group_solver_solve_host runs the row-sharded solve of SURVEY 8e:
// local A_i'(A_i x), one sum-all-reduce of the n-vector per apply, replicated epilogues.
//
// Mechanism: one context per device, one NCCL communicator over them (ncclCommInitAll), and one short-lived host thread
// per device for the duration of a call.  Every thread runs the unchanged single-context entry points on its device and
// joins the collectives from its own thread — NCCL's thread-per-rank mode — so no code path is duplicated and the
// replicas stay bit-identical exactly as in the process-per-GPU harness (tests/test_gpu_multi.py).
#include "rls_common.cuh"

#include <functional>
#include <thread>

struct rls_group_s {
  std::vector<rls_ctx_s*> ctx;
};
struct rls_gmat_s {
  rls_group_s* g = nullptr;
  int32_t dtype = 0;
  int64_t m = 0, n = 0;
  std::vector<rls_mat_s*> part;
  std::vector<int64_t> lo, hi;
  std::atomic<int> refs{1};   // the creator + every group solver built on it (handles are freed in arbitrary order)
};
struct rls_gsolver_s {
  rls_gmat_s* A = nullptr;
  std::vector<rls_normal_t> op;
  std::vector<rls_solver_t> sol;
  std::vector<std::vector<char>> xbuf;   // host landing buffers of the replicas other than device 0
};

namespace {

// fn(i) on one host thread per device; the first failure (in device order) is reported with its message
int32_t run_all(int n, const std::function<int32_t(int)>& fn) {
  std::vector<int32_t> st(n, RLS_OK);
  std::vector<std::string> msg(n);
  auto body = [&](int i) {
    st[i] = fn(i);
    if (st[i] != RLS_OK) msg[i] = rls_last_error();   // the error text is thread-local
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back(body, i);
  body(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (st[i] != RLS_OK) {
      rls_set_error("device %d of the group: %s", i, msg[i].c_str());
      return st[i];
    }
  return RLS_OK;
}

// contiguous row blocks, multiples of 4 rows except possibly the last (the rule of dist.py row_range)
void row_range(int64_t m, int rank, int nranks, int64_t* lo, int64_t* hi) {
  int64_t per = (m + nranks - 1) / nranks;
  per = (per + 3) / 4 * 4;
  *lo = std::min<int64_t>(m, (int64_t)rank * per);
  *hi = std::min<int64_t>(m, *lo + per);
}

}  // namespace

extern "C" int32_t rls_group_create(int32_t ndev, const int32_t* dev_ids, rls_group_t* out) {
  RLS_CHECK_ARG(out && ndev >= 1 && ndev <= RLS_MAX_PEERS, "group of 1..%d devices", RLS_MAX_PEERS);
  rls_group_s* g = new rls_group_s();
  int32_t st = RLS_OK;
  for (int i = 0; i < ndev && st == RLS_OK; ++i) {
    rls_ctx_s* c = nullptr;
    st = rls_ctx_create(dev_ids ? dev_ids[i] : i, &c);
    if (st == RLS_OK) g->ctx.push_back(c);
  }
  if (st == RLS_OK && ndev > 1) st = rls_comm_init_all(g->ctx.data(), ndev);
  if (st != RLS_OK) {
    for (rls_ctx_s* c : g->ctx) rls_ctx_destroy(c);
    delete g;
    return st;
  }
  *out = g;
  return RLS_OK;
}

extern "C" int32_t rls_group_destroy(rls_group_t g) {
  if (!g) return RLS_OK;
  for (rls_ctx_s* c : g->ctx) rls_ctx_destroy(c);   // reference-counted: matrices / solvers still alive keep their context
  delete g;
  return RLS_OK;
}

extern "C" int32_t rls_group_size(rls_group_t g, int32_t* ndev) {
  RLS_CHECK_ARG(g && ndev, "NULL argument");
  *ndev = (int32_t)g->ctx.size();
  return RLS_OK;
}

extern "C" int32_t rls_group_ctx(rls_group_t g, int32_t i, rls_ctx_t* ctx) {
  RLS_CHECK_ARG(g && ctx && i >= 0 && i < (int32_t)g->ctx.size(), "bad group member %d", i);
  *ctx = g->ctx[i];
  return RLS_OK;
}

extern "C" int32_t rls_group_mat_create(rls_group_t g, int32_t dtype, int64_t m, int64_t n, const void* host, int64_t ld, rls_gmat_t* out) {
  RLS_CHECK_ARG(g && out, "NULL argument");
  RLS_CHECK_ARG(m >= 0 && n >= 0 && (!host || ld >= m), "bad shape / leading dimension");
  const int nd = (int)g->ctx.size();
  rls_gmat_s* A = new rls_gmat_s();
  A->g = g; A->dtype = dtype; A->m = m; A->n = n;
  A->part.assign(nd, nullptr); A->lo.resize(nd); A->hi.resize(nd);
  for (int i = 0; i < nd; ++i) row_range(m, i, nd, &A->lo[i], &A->hi[i]);
  const size_t es = rls_elem_size(dtype);
  int32_t st = run_all(nd, [&](int i) {
    // rows [lo, hi) of a column-major host matrix: same leading dimension, shifted base
    const void* hp = host ? (const char*)host + (size_t)A->lo[i] * es : nullptr;
    return rls_mat_create_layout(g->ctx[i], dtype, A->hi[i] - A->lo[i], n, hp, ld, RLS_LAYOUT_AUTO, &A->part[i]);
  });
  if (st != RLS_OK) { rls_group_mat_destroy(A); return st; }
  *out = A;
  return RLS_OK;
}

extern "C" int32_t rls_group_mat_fill_philox(rls_gmat_t A, uint64_t seed, int32_t dist, float scale) {
  RLS_CHECK_ARG(A, "NULL argument");
  return run_all((int)A->part.size(), [&](int i) { return rls_mat_fill_philox(A->part[i], seed, dist, scale, A->lo[i], A->m); });
}

extern "C" int32_t rls_group_mat_part(rls_gmat_t A, int32_t i, rls_mat_t* part, int64_t* row_lo, int64_t* row_hi) {
  RLS_CHECK_ARG(A && i >= 0 && i < (int32_t)A->part.size(), "bad group member %d", i);
  if (part) *part = A->part[i];
  if (row_lo) *row_lo = A->lo[i];
  if (row_hi) *row_hi = A->hi[i];
  return RLS_OK;
}

extern "C" int32_t rls_group_mat_destroy(rls_gmat_t A) {
  if (!A || A->refs.fetch_sub(1) != 1) return RLS_OK;
  for (rls_mat_s* p : A->part) rls_mat_destroy(p);
  delete A;
  return RLS_OK;
}

extern "C" int32_t rls_group_solver_create(rls_gmat_t A, int32_t normal_form, const rls_solver_desc* desc, rls_gsolver_t* out) {
  RLS_CHECK_ARG(A && desc && out, "NULL argument");
  const int nd = (int)A->part.size();
  rls_gsolver_s* s = new rls_gsolver_s();
  s->A = A;
  A->refs.fetch_add(1);
  s->op.assign(nd, nullptr); s->sol.assign(nd, nullptr); s->xbuf.resize(nd);
  int32_t st = run_all(nd, [&](int i) {
    RLS_TRY(rls_normal_create(A->part[i], normal_form, &s->op[i]));
    return rls_solver_create(A->part[i], s->op[i], desc, &s->sol[i]);
  });
  if (st != RLS_OK) { rls_group_solver_destroy(s); return st; }
  *out = s;
  return RLS_OK;
}

extern "C" int32_t rls_group_solver_destroy(rls_gsolver_t s) {
  if (!s) return RLS_OK;
  for (size_t i = 0; i < s->sol.size(); ++i) {
    if (s->sol[i]) rls_solver_destroy(s->sol[i]);
    if (s->op[i]) rls_normal_destroy(s->op[i]);
  }
  rls_group_mat_destroy(s->A);
  delete s;
  return RLS_OK;
}

// solve!(solver, b): b (m elements) and x (n elements) on the host; every device gets its rows of b, device 0's replica
// of x is returned (the replicas are bit-identical: deterministic epilogues on identical all-reduced vectors)
extern "C" int32_t rls_group_solver_solve_host(rls_gsolver_t s, const void* b_host, int64_t b_len, void* x_host, int64_t x_len,
                                               int32_t* iterations_done, rls_solver_scalars* scalars) {
  RLS_CHECK_ARG(s && b_host && x_host, "NULL argument");
  rls_gmat_s* A = s->A;
  RLS_CHECK_ARG(b_len == A->m && x_len == A->n, "solve!: b has %lld elements (A has %lld rows), x %lld (A has %lld columns)",
                (long long)b_len, (long long)A->m, (long long)x_len, (long long)A->n);
  RlsNvtxRange nvtx("rls: solve! (device group)");
  const int nd = (int)s->sol.size();
  const size_t es = rls_elem_size(A->dtype);
  std::vector<int32_t> its(nd, 0);
  std::vector<rls_solver_scalars> sc(nd);
  int32_t st = run_all(nd, [&](int i) {
    void* xi = x_host;
    if (i > 0) { s->xbuf[i].resize((size_t)x_len * es); xi = s->xbuf[i].data(); }
    return rls_solver_solve_host(s->sol[i], (const char*)b_host + (size_t)A->lo[i] * es, A->hi[i] - A->lo[i], xi, x_len, &its[i], &sc[i]);
  });
  RLS_TRY(st);
  for (int i = 1; i < nd; ++i)
    if (its[i] != its[0] || memcmp(s->xbuf[i].data(), x_host, (size_t)x_len * es) != 0) {
      rls_set_error("the replicas of the device group diverged (device %d: %d iterations, device 0: %d)", i, its[i], its[0]);
      return RLS_ERR_CUDA;
    }
  if (iterations_done) *iterations_done = its[0];
  if (scalars) *scalars = sc[0];
  return RLS_OK;
}
