// Two-site infinite DMRG on the device (MPSKit `IDMRG2`).
//
// Replaces `find_groundstate(psi, H, IDMRG2(trscheme = truncbelow(cut), tol))`
// (/root/reference/src/HubbardFunctions.jl:1010): sweeps L->R, edge, R->L, edge
// over the unit cell; per bond the lowest eigenvector of H_AC2 (Lanczos), a truncated SVD that
// re-defines the bond space, and growth of the environments by one site transfer.  Statement by
// statement the algorithm of oracle/twosite.py:idmrg2.  Bond spaces change every step, so tensors,
// environments and contraction programs are re-planned on the host per step; all block data stay in
// HBM (only singular values and a few scalars travel).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

#include "htn_linalg.hpp"

using namespace htn;

extern "C" {
bool htn_same_structure(const htn_tensor* x, const htn_tensor* y);
int32_t htn_heff_run(htn_plan* p, const double* x, double* y, int mask);
int32_t htn_upload_locked(htn_tensor* t, const double* host, int64_t nelem);
int32_t htn_download_locked(const htn_tensor* t, double* host, int64_t nelem);
}

namespace {

#define RC(call)             \
  do {                       \
    int32_t rc_ = (call);    \
    if (rc_ < 0) return rc_; \
  } while (0)

bool same_space(const htn_space& a, const htn_space& b) { return a.sec == b.sec && a.mult == b.mult; }

// out[l,s,r] = A[l,s,r] . B[r]  (right = true)   or   B[l] . A[l,s,r]  (right = false); out is created
int32_t mul_bond(htn_ctx* ctx, const htn_tensor* A, const htn_tensor* B, bool right, htn_tensor** out) {
  if (!same_space(right ? A->s1 : A->s0, B->s0)) return ctx->fail(HTN_ERR_SHAPE, "idmrg2: bond space mismatch in A.C product");
  RC(htn_tensor_create_like(A, out));
  Program pg;
  std::vector<GemmTaskH> tasks;
  for (const Block& b : A->blocks) {
    const Block& c = B->blocks[right ? b.lab[2] : b.lab[0]];
    GemmTaskH t;
    t.C = Opnd{2, b.off};
    t.ldc = b.ld;
    t.M = b.rows;
    t.N = b.cols;
    if (right)
      t.segs.push_back(GemmSegH{Opnd{0, b.off}, b.ld, Opnd{1, c.off}, c.ld, b.cols});
    else
      t.segs.push_back(GemmSegH{Opnd{1, c.off}, c.ld, Opnd{0, b.off}, b.ld, b.rows});
    tasks.push_back(std::move(t));
  }
  pg.add_gemm(tasks, TAG_L);
  int32_t rc = pg.finalize(ctx, 3);
  if (rc == HTN_OK) {
    const double* slots[3] = {A->d, B->d, (*out)->d};
    rc = pg.run(slots);
  }
  cudaStreamSynchronize(ctx->stream);
  pg.destroy();
  return rc;
}

// out = C^-1 block by block (MPSKit idmrg2: `inv(psi.C[end])`).  The bonds produced by the truncated SVD are diagonal
// and are inverted on the device; the caller's initial C[L-1] is a general matrix (triangular after the QR gauge,
// dense after VUMPS): those blocks go through an LU inverse with partial pivoting on the host (<= 342^2 each, once per
// run -- every later edge bond is diagonal).
int32_t inv_bond(htn_ctx* ctx, const htn_tensor* C, htn_tensor** out) {
  RC(htn_tensor_create_like(C, out));
  if (C->blocks.empty()) return HTN_OK;
  std::vector<double> host(C->hsize);
  RC(htn_download_locked(C, host.data(), C->hsize));
  bool all_diag = true;
  for (const Block& b : C->blocks) {
    const double* B = host.data() + b.hoff;
    const int n = b.rows;
    for (int i = 0; i < n && all_diag; ++i)
      for (int j = 0; j < n; ++j)
        if (i != j && B[(size_t)i * n + j] != 0.0) {
          all_diag = false;
          break;
        }
    if (!all_diag) break;
  }
  if (all_diag) {
    std::vector<FillBlock> fb;
    for (const Block& b : C->blocks) fb.push_back(FillBlock{b.off, b.rows, b.cols, b.ld, 0});
    FillBlock* d = nullptr;
    if (cudaMalloc(&d, fb.size() * sizeof(FillBlock)) != cudaSuccess) return ctx->fail(HTN_ERR_OOM, "idmrg2: table allocation failed");
    htn::h2d_on_stream(d, fb.data(), fb.size() * sizeof(FillBlock), ctx->stream);
    launch_diag_inv(d, (int)fb.size(), C->d, (*out)->d, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    return HTN_OK;
  }
  std::vector<double> inv(C->hsize, 0.0);
  for (const Block& b : C->blocks) {
    const int n = b.rows;
    if (b.cols != n) return ctx->fail(HTN_ERR_SHAPE, "idmrg2: bond blocks must be square");
    std::vector<double> A(host.begin() + b.hoff, host.begin() + b.hoff + (size_t)n * n);
    double* X = inv.data() + b.hoff;  // starts as the identity, ends as A^-1 (Gauss-Jordan with row pivoting)
    for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0;
    double amax = 0.0;
    for (double v : A) amax = std::max(amax, std::fabs(v));
    for (int k = 0; k < n; ++k) {
      int piv = k;
      for (int i = k + 1; i < n; ++i)
        if (std::fabs(A[(size_t)i * n + k]) > std::fabs(A[(size_t)piv * n + k])) piv = i;
      if (std::fabs(A[(size_t)piv * n + k]) <= 1e-300 * (1.0 + amax))
        return ctx->fail(HTN_ERR_INVALID, "idmrg2: singular bond matrix on the unit-cell edge");
      if (piv != k)
        for (int j = 0; j < n; ++j) {
          std::swap(A[(size_t)k * n + j], A[(size_t)piv * n + j]);
          std::swap(X[(size_t)k * n + j], X[(size_t)piv * n + j]);
        }
      const double d = 1.0 / A[(size_t)k * n + k];
      for (int j = 0; j < n; ++j) {
        A[(size_t)k * n + j] *= d;
        X[(size_t)k * n + j] *= d;
      }
      for (int i = 0; i < n; ++i) {
        if (i == k) continue;
        const double f = A[(size_t)i * n + k];
        if (f == 0.0) continue;
        for (int j = 0; j < n; ++j) {
          A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
          X[(size_t)i * n + j] -= f * X[(size_t)k * n + j];
        }
      }
    }
  }
  return htn_upload_locked(*out, inv.data(), C->hsize);
}

// temporaries of one bond update: destroyed on every early return, handed over with release() once they replace state
struct Temps {
  std::vector<htn_tensor**> slots;
  void own(htn_tensor** t) { slots.push_back(t); }
  void release(htn_tensor** t) { slots.erase(std::remove(slots.begin(), slots.end(), t), slots.end()); }
  ~Temps() {
    for (htn_tensor** t : slots)
      if (*t) {
        htn_tensor_destroy(*t);
        *t = nullptr;
      }
  }
};

void replace(htn_tensor*& slot, htn_tensor* nw) {
  if (slot && slot != nw) htn_tensor_destroy(slot);
  slot = nw;
}

// singular values of every block of a bond tensor on the host (diagonal blocks: |diag|; general
// blocks -- only the caller's initial C -- via the eigenvalues of B^T B)
int32_t bond_singular_values(const htn_tensor* C, std::vector<std::vector<double>>& out) {
  std::vector<double> host(C->hsize);
  RC(htn_download_locked(C, host.data(), C->hsize));
  out.assign(C->blocks.size(), {});
  for (size_t bi = 0; bi < C->blocks.size(); ++bi) {
    const Block& b = C->blocks[bi];
    const int n = b.rows;
    const double* B = host.data() + b.hoff;
    bool diag = true;
    for (int i = 0; i < n && diag; ++i)
      for (int j = 0; j < n; ++j)
        if (i != j && B[(size_t)i * n + j] != 0.0) {
          diag = false;
          break;
        }
    std::vector<double> sv(n);
    if (diag) {
      for (int i = 0; i < n; ++i) sv[i] = std::fabs(B[(size_t)i * n + i]);
    } else {
      // cyclic Jacobi on G = B^T B
      std::vector<double> G((size_t)n * n, 0.0);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0.0;
          for (int k = 0; k < n; ++k) s += B[(size_t)k * n + i] * B[(size_t)k * n + j];
          G[(size_t)i * n + j] = s;
        }
      for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < n; ++i)
          for (int j = i + 1; j < n; ++j) off += G[(size_t)i * n + j] * G[(size_t)i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; ++p)
          for (int q = p + 1; q < n; ++q) {
            const double apq = G[(size_t)p * n + q];
            if (std::fabs(apq) < 1e-300) continue;
            const double tau = (G[(size_t)q * n + q] - G[(size_t)p * n + p]) / (2.0 * apq);
            const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
            const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
            for (int k = 0; k < n; ++k) {
              const double akp = G[(size_t)k * n + p], akq = G[(size_t)k * n + q];
              G[(size_t)k * n + p] = c * akp - s * akq;
              G[(size_t)k * n + q] = s * akp + c * akq;
            }
            for (int k = 0; k < n; ++k) {
              const double apk = G[(size_t)p * n + k], aqk = G[(size_t)q * n + k];
              G[(size_t)p * n + k] = c * apk - s * aqk;
              G[(size_t)q * n + k] = s * apk + c * aqk;
            }
          }
      }
      for (int i = 0; i < n; ++i) sv[i] = std::sqrt(std::max(G[(size_t)i * n + i], 0.0));
    }
    std::sort(sv.begin(), sv.end(), [](double a, double b) { return a > b; });
    out[bi] = sv;
  }
  return HTN_OK;
}

struct Idmrg {
  htn_ctx* ctx;
  int L, chi;
  std::vector<htn_tensor*> AL, AR, AC, C, GL, GR;
  std::vector<const htn_mpo*> W;
  int krylovdim;
  double eig_tol, cut;
  int maxdim;
  long applies = 0;
  double t_plan = 0, t_eig = 0, t_svd = 0, t_env = 0;  // seconds (stream-synchronised phase timers)

  double now() {
    cudaStreamSynchronize(ctx->stream);
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }

  ~Idmrg() {
    for (htn_tensor* t : GL) htn_tensor_destroy(t);
    for (htn_tensor* t : GR) htn_tensor_destroy(t);
  }

  int32_t unit_env(int side, const htn_space* V, const htn_legs* M, int level, htn_tensor** out) {
    RC(htn_tensor_create_env(ctx, side, V, M, level, out));
    return t_fill_level(*out, level, 1);
  }

  // GL of site i+1 from site i
  int32_t grow_left(int i) {
    const double t0 = now();
    struct Acc {
      Idmrg* d;
      double t0;
      ~Acc() { d->t_env += d->now() - t0; }
    } acc{this, t0};
    const int j = (i + 1) % L;
    htn_tensor *At = nullptr, *nw = nullptr;
    htn_plan* p = nullptr;
    int32_t rc = htn_tensor_create_transposed(AL[i], &At);
    if (rc == HTN_OK) rc = t_transpose(AL[i], At, 0);
    if (rc == HTN_OK) rc = htn_tensor_create_env(ctx, HTN_SIDE_LEFT, &AL[i]->s1, &W[i]->Mr, 0, &nw);
    if (rc == HTN_OK) rc = htn_plan_transfer(ctx, HTN_SIDE_LEFT, W[i], AL[i], At, GL[i], nw, &p);
    if (rc == HTN_OK) rc = htn_transfer_apply(p, AL[i], At, GL[i], nw);
    cudaStreamSynchronize(ctx->stream);
    if (p) htn_plan_destroy(p);
    if (At) htn_tensor_destroy(At);
    if (rc < 0) {
      if (nw) htn_tensor_destroy(nw);
      return rc;
    }
    replace(GL[j], nw);
    return HTN_OK;
  }

  // GR of site j-1 from site j
  int32_t grow_right(int j) {
    const double t0 = now();
    struct Acc {
      Idmrg* d;
      double t0;
      ~Acc() { d->t_env += d->now() - t0; }
    } acc{this, t0};
    const int i = (j + L - 1) % L;
    htn_tensor *At = nullptr, *nw = nullptr;
    htn_plan* p = nullptr;
    int32_t rc = htn_tensor_create_transposed(AR[j], &At);
    if (rc == HTN_OK) rc = t_transpose(AR[j], At, 0);
    if (rc == HTN_OK) rc = htn_tensor_create_env(ctx, HTN_SIDE_RIGHT, &AR[j]->s0, &W[j]->Ml, chi - 1, &nw);
    if (rc == HTN_OK) rc = htn_plan_transfer(ctx, HTN_SIDE_RIGHT, W[j], AR[j], At, GR[j], nw, &p);
    if (rc == HTN_OK) rc = htn_transfer_apply(p, AR[j], At, GR[j], nw);
    cudaStreamSynchronize(ctx->stream);
    if (p) htn_plan_destroy(p);
    if (At) htn_tensor_destroy(At);
    if (rc < 0) {
      if (nw) htn_tensor_destroy(nw);
      return rc;
    }
    replace(GR[i], nw);
    return HTN_OK;
  }

  // x2 = A1 . A2 ; lowest eigenvector of H_AC2(i, j) ; truncated SVD -> (al, c normalised, ar)
  int32_t solve_and_split(int i, int j, const htn_tensor* A1, const htn_tensor* A2, htn_tensor** al, htn_tensor** c,
                          htn_tensor** ar) {
    htn_tensor *x2 = nullptr, *y2 = nullptr;
    htn_plan* p = nullptr;
    htn_space* Vm = nullptr;
    const double t0 = now();
    int32_t rc = htn_tensor_create_mps2(ctx, &A1->s0, &A1->legs, &A2->legs, &A2->s1, &x2);
    if (rc == HTN_OK) rc = htn_contract_two_site(A1, A2, x2);
    if (rc == HTN_OK) rc = htn_tensor_create_like(x2, &y2);
    if (krylovdim <= 0) {
      // truncation-only sweep (MPSKit `changebonds(psi, SvdCut(trscheme))`): no eigensolve, x2 is split as is
      if (rc == HTN_OK) rc = t_copy(x2, y2);
    } else {
      if (rc == HTN_OK) rc = htn_plan_heff_ac2(ctx, GL[i], W[i], W[j], GR[j], x2, &p);
      const double t1 = now();
      t_plan += t1 - t0;
      if (rc == HTN_OK) {
        ApplyFn op = [&](const double* a, double* b) -> int32_t { return htn_heff_run(p, a, b, 0xF); };
        KrylovInfo info;
        rc = lanczos_lowest(x2, op, x2->d, y2->d, krylovdim, eig_tol, 6, &info);  // 30 + 5 x 12 applies: the budget of 3 explicit restarts
        applies += info.applies;
        if (rc > 0) rc = HTN_OK;
      }
      t_eig += now() - t1;
    }
    const double t2 = now();
    if (rc == HTN_OK) rc = htn_tsvd(y2, cut, maxdim, &Vm, al, c, ar, nullptr, nullptr);
    if (rc == HTN_OK) rc = t_normalize(*c, (*c)->d);
    t_svd += now() - t2;
    cudaStreamSynchronize(ctx->stream);
    if (p) htn_plan_destroy(p);
    if (x2) htn_tensor_destroy(x2);
    if (y2) htn_tensor_destroy(y2);
    if (Vm) htn_space_destroy(Vm);
    return rc;
  }

  // standard bond update (sites i, i+1 inside the cell)
  int32_t update_bond(int i, const htn_tensor* A1, const htn_tensor* A2) {
    htn_tensor *al = nullptr, *c = nullptr, *ar = nullptr, *ac1 = nullptr, *ac2 = nullptr;
    Temps tmp;
    tmp.own(&al);
    tmp.own(&c);
    tmp.own(&ar);
    tmp.own(&ac1);
    tmp.own(&ac2);
    RC(solve_and_split(i, i + 1, A1, A2, &al, &c, &ar));
    RC(mul_bond(ctx, al, c, true, &ac1));
    RC(mul_bond(ctx, ar, c, false, &ac2));
    tmp.slots.clear();  // from here on the state owns them
    replace(AL[i], al);
    replace(C[i], c);
    replace(AR[i + 1], ar);
    replace(AC[i], ac1);
    replace(AC[i + 1], ac2);
    RC(grow_left(i));
    RC(grow_right(i + 1));
    return HTN_OK;
  }
};

}  // namespace

extern "C" {

// In/out: AL, AR, C, AC handle arrays (the driver destroys the tensors it replaces and stores the new
// handles; the caller destroys the final ones).  log rows of 8 doubles: eps, sum of D_red over bonds, cumulative
// H_AC2 applies, cumulative seconds in planning / Lanczos / SVD / environment growth, 0.
int32_t htn_idmrg2(htn_ctx* ctx, int32_t nsites, htn_tensor** AL, htn_tensor** AR, htn_tensor** C, htn_tensor** AC,
                   const htn_mpo* const* W, double cut, double tol, int32_t maxiter, int32_t krylovdim, double eig_tol,
                   int32_t maxdim, double* delta, int32_t* iterations, double* log, int32_t log_cap) {
  if (!ctx || nsites < 2 || !AL || !AR || !C || !AC || !W) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  const int L = nsites;
  Idmrg D;
  D.ctx = ctx;
  D.L = L;
  D.chi = (int)W[0]->Ml.sec.size();
  D.krylovdim = krylovdim;
  D.eig_tol = eig_tol;
  D.cut = cut;
  D.maxdim = maxdim;
  for (int i = 0; i < L; ++i) {
    if (!AL[i] || !AR[i] || !C[i] || !AC[i] || !W[i]) return HTN_ERR_INVALID;
    if ((int)W[i]->Ml.sec.size() != D.chi || (int)W[i]->Mr.sec.size() != D.chi)
      return ctx->fail(HTN_ERR_SHAPE, "idmrg2: every site must carry the same MPO levels");
    D.AL.push_back(AL[i]);
    D.AR.push_back(AR[i]);
    D.C.push_back(C[i]);
    D.AC.push_back(AC[i]);
    D.W.push_back(W[i]);
  }
  auto sync_out = [&]() {
    for (int i = 0; i < L; ++i) {
      AL[i] = D.AL[i];
      AR[i] = D.AR[i];
      C[i] = D.C[i];
      AC[i] = D.AC[i];
    }
  };
  D.GL.assign(L, nullptr);
  D.GR.assign(L, nullptr);
  int32_t rc = HTN_OK;
  for (int i = 0; i < L && rc == HTN_OK; ++i) {
    rc = D.unit_env(HTN_SIDE_LEFT, &D.AL[i]->s0, &W[i]->Ml, 0, &D.GL[i]);
    if (rc == HTN_OK) rc = D.unit_env(HTN_SIDE_RIGHT, &D.AR[i]->s1, &W[i]->Mr, D.chi - 1, &D.GR[i]);
  }
  static const bool unit_start = getenv("HTN_IDMRG_UNIT_ENV") != nullptr;  // experiments: textbook iDMRG start
  if (krylovdim > 0 && !unit_start) {
    // MPSKit: find_groundstate(psi, H, alg::IDMRG2, envs = environments(psi, H)) -- the sweeps start from
    // the infinite environments of the initial uniform state (oracle/twosite.py:idmrg2, init_env="infinite");
    // empty environments let charge escape to the edges and can trap the edge bond in a one-multiplet sector
    double el = 0, er = 0;
    if (rc == HTN_OK) {
      rc = htn_environments(ctx, L, D.AL.data(), D.AR.data(), D.C.data(), W, D.GL.data(), D.GR.data(), 1e-10, 30, 200, &el, &er);
      if (rc > 0) rc = HTN_OK;  // not fully converged environments of a random state are still a valid start
    }
  } else {
    for (int i = 0; i + 1 < L && rc == HTN_OK; ++i) rc = D.grow_left(i);
    for (int i = L - 1; i > 0 && rc == HTN_OK; --i) rc = D.grow_right(i);
  }
  if (rc < 0) return rc;
  double eps = 1e300;
  int it = 0;
  auto body = [&]() -> int32_t {
    for (it = 1; it <= maxiter; ++it) {
      std::vector<std::vector<double>> sv_old;
      std::vector<Sector> sec_old = D.C[L - 1]->s0.sec;
      RC(bond_singular_values(D.C[L - 1], sv_old));
      // ---- left -> right ----
      for (int i = 0; i + 1 < L; ++i) RC(D.update_bond(i, D.AC[i], D.AR[i + 1]));
      // ---- edge (sites L-1, 0) ----
      {
        htn_tensor *ci = nullptr, *left = nullptr, *right = nullptr, *al = nullptr, *c = nullptr, *ar = nullptr;
        htn_tensor *ac1 = nullptr, *ac0 = nullptr, *c0i = nullptr, *al0 = nullptr;
        Temps tmp;
        for (htn_tensor** t : {&ci, &left, &right, &al, &c, &ar, &ac1, &ac0, &c0i, &al0}) tmp.own(t);
        RC(inv_bond(ctx, D.C[L - 1], &ci));
        RC(mul_bond(ctx, D.AC[L - 1], ci, true, &left));
        RC(mul_bond(ctx, D.AL[0], D.C[0], true, &right));
        RC(D.solve_and_split(L - 1, 0, left, right, &al, &c, &ar));
        RC(mul_bond(ctx, al, c, true, &ac1));
        RC(mul_bond(ctx, ar, c, false, &ac0));
        for (htn_tensor** t : {&al, &c, &ar, &ac1, &ac0}) tmp.release(t);  // the state owns them from here on
        replace(D.AL[L - 1], al);
        replace(D.C[L - 1], c);
        replace(D.AR[0], ar);
        replace(D.AC[L - 1], ac1);
        replace(D.AC[0], ac0);
        RC(inv_bond(ctx, D.C[0], &c0i));
        RC(mul_bond(ctx, D.AC[0], c0i, true, &al0));
        tmp.release(&al0);
        replace(D.AL[0], al0);
        RC(D.grow_left(L - 1));
        RC(D.grow_right(0));
      }
      // ---- right -> left ----
      for (int i = L - 2; i >= 0; --i) RC(D.update_bond(i, D.AL[i], D.AC[i + 1]));
      // ---- edge again ----
      {
        htn_tensor *ci = nullptr, *left = nullptr, *right = nullptr, *al = nullptr, *c = nullptr, *ar = nullptr;
        htn_tensor *ac1 = nullptr, *ac0 = nullptr, *cmi = nullptr, *arl = nullptr;
        Temps tmp;
        for (htn_tensor** t : {&ci, &left, &right, &al, &c, &ar, &ac1, &ac0, &cmi, &arl}) tmp.own(t);
        RC(inv_bond(ctx, D.C[L - 1], &ci));
        RC(mul_bond(ctx, D.AC[0], ci, false, &right));
        RC(mul_bond(ctx, D.AR[L - 1], D.C[L - 2], false, &left));
        RC(D.solve_and_split(L - 1, 0, left, right, &al, &c, &ar));
        RC(mul_bond(ctx, al, c, true, &ac1));
        RC(mul_bond(ctx, ar, c, false, &ac0));
        for (htn_tensor** t : {&al, &c, &ar, &ac1, &ac0}) tmp.release(t);
        replace(D.AL[L - 1], al);
        replace(D.C[L - 1], c);
        replace(D.AR[0], ar);
        replace(D.AC[L - 1], ac1);
        replace(D.AC[0], ac0);
        RC(inv_bond(ctx, D.C[L - 2], &cmi));
        RC(mul_bond(ctx, D.AC[L - 1], cmi, false, &arl));
        tmp.release(&arl);
        replace(D.AR[L - 1], arl);
        RC(D.grow_left(L - 1));
        RC(D.grow_right(0));
      }
      // ---- error on the common subspace of the edge bond ----
      std::vector<std::vector<double>> sv_new;
      RC(bond_singular_values(D.C[L - 1], sv_new));
      const std::vector<Sector>& sec_new = D.C[L - 1]->s0.sec;
      double d2 = 0.0;
      for (size_t a = 0; a < sec_old.size(); ++a)
        for (size_t b = 0; b < sec_new.size(); ++b)
          if (sec_old[a] == sec_new[b]) {
            const size_t n = std::min(sv_old[a].size(), sv_new[b].size());
            for (size_t k = 0; k < n; ++k) d2 += sdim(D.C[L - 1]->sym, sec_new[b]) * (sv_old[a][k] - sv_new[b][k]) * (sv_old[a][k] - sv_new[b][k]);
          }
      eps = std::sqrt(d2);
      if (log && it <= log_cap) {
        double dsum = 0;
        for (int i = 0; i < L; ++i)
          for (int m : D.C[i]->s0.mult) dsum += m;
        double* row = log + 8 * (it - 1);
        row[0] = eps;
        row[1] = dsum;
        row[2] = (double)D.applies;
        row[3] = D.t_plan;  // cumulative seconds: tensors + two-site contraction + H_AC2 planning (host)
        row[4] = D.t_eig;   // Lanczos
        row[5] = D.t_svd;   // truncated SVD
        row[6] = D.t_env;   // environment growth (planning + transfers)
        row[7] = 0.0;
      }
      if (eps < tol) break;
    }
    return HTN_OK;
  };
  rc = body();
  sync_out();
  cudaStreamSynchronize(ctx->stream);
  if (rc < 0) return rc;
  if (delta) *delta = eps;
  if (iterations) *iterations = std::min(it, (int)maxiter);
  return eps < tol ? HTN_OK : HTN_NOT_CONVERGED;
}

// out = A . C (right) or C . A (left): the AC = AL C = C AR products of MPSKit (`_mul_tail` / `_mul_front`)
int32_t htn_mul_bond(const htn_tensor* A, const htn_tensor* Cb, int32_t right, htn_tensor** out) {
  if (!A || !Cb || !out) return HTN_ERR_INVALID;
  htn_ctx* ctx = A->ctx;
  *out = nullptr;
  if (A->kind != HTN_T_MPS || Cb->kind != HTN_T_BOND) return ctx->fail(HTN_ERR_INVALID, "mul_bond: needs an MPS and a bond tensor");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  int32_t rc = mul_bond(ctx, A, Cb, right != 0, out);
  if (rc != HTN_OK && *out) {
    htn_tensor_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

// Gauge-fix site tensors into a consistent uniform MPS in mixed gauge (MPSKit `InfiniteMPS(A...)`).
//   from_right = 0: AL[i] hold (approximate) left isometries: AL <- Q(AL), then AR, C by the iterated LQ;
//   from_right = 1: AR[i] hold (approximate) right isometries -- the list whose bond spaces chain after
//                   an IDMRG2 iteration: AL, C by the iterated QR from AR, then AR, C by the iterated LQ.
// All of AL[i], AR[i], AC[i] share one block structure; C[i] is a bond tensor on the right space of
// site i; C_guess lives on the last bond.  AC = AL C on exit.
int32_t htn_mixed_gauge(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, const htn_tensor* C_guess,
                        htn_tensor* const* AR, htn_tensor* const* C, htn_tensor* const* AC, int32_t from_right,
                        double tol, int32_t maxiter, int32_t* iterations) {
  if (!ctx || nsites <= 0 || !AL || !C_guess || !AR || !C || !AC) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  int32_t rc = HTN_OK;
  if (from_right) {
    int32_t it1 = 0;
    rc = htn_gauge_left(ctx, nsites, AR, C_guess, AL, C, tol, maxiter, &it1, nullptr);
    if (rc < 0) return rc;
  } else {
    for (int i = 0; i < nsites; ++i) {
      htn_tensor* R = nullptr;
      RC(htn_tensor_create_like(C[i], &R));
      int32_t r2 = t_qr_inplace(AL[i], R);
      cudaStreamSynchronize(ctx->stream);
      htn_tensor_destroy(R);
      RC(r2);
    }
  }
  htn_tensor* guess = nullptr;
  RC(htn_tensor_create_like(C[nsites - 1], &guess));
  if (from_right)
    t_copy(C[nsites - 1], guess);
  else
    t_copy(C_guess, guess);
  double d = 0;
  rc = htn_gauge_right(ctx, nsites, AL, guess, AR, C, tol, maxiter, iterations, &d);
  htn_tensor_destroy(guess);
  if (rc < 0) return rc;
  for (int i = 0; i < nsites; ++i) {
    htn_tensor* ac = nullptr;
    RC(mul_bond(ctx, AL[i], C[i], true, &ac));
    int32_t r2 = t_copy(ac, AC[i]);
    cudaStreamSynchronize(ctx->stream);
    htn_tensor_destroy(ac);
    RC(r2);
  }
  return rc;
}

}  // extern "C"
