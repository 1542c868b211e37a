// Tensor-level linear algebra of the gauge and Krylov steps: blockwise transposes, level fills,
// positive QR / LQ per coupled sector, Lanczos and GMRES over device-resident vectors.
//
// Replaces: TensorKit `leftorth!(.., QRpos())` / `rightorth!(.., LQpos())`
// (LAPACK geqrf under MKL_jll 2025.0.1, Manifest.toml:716) and KrylovKit 0.9.5 `eigsolve`
// (Lanczos) / `linsolve` (GMRES) (Manifest.toml:548).  The Krylov recurrences are driven from the
// host, but every vector stays in HBM: per iteration one fused "all inner products" kernel pair, one
// rank-k update, and a single readback of <= 130 scalars.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>

#include "htn_linalg.hpp"

using namespace htn;

extern "C" bool htn_same_structure(const htn_tensor* x, const htn_tensor* y);

namespace htn {

static int32_t cuda_rc(htn_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return HTN_OK;
}

template <class T>
static int32_t cached_table(htn_tensor* owner, const void* partner, int mode, const std::vector<T>& host, T** dev,
                            int* n) {
  htn_ctx* ctx = owner->ctx;
  auto key = std::make_pair(partner, mode);
  auto it = owner->devtables.find(key);
  if (it != owner->devtables.end()) {
    *dev = static_cast<T*>(it->second.first);
    *n = it->second.second;
    return HTN_OK;
  }
  T* d = nullptr;
  if (!host.empty()) {
    if (cudaMalloc(&d, host.size() * sizeof(T)) != cudaSuccess) return ctx->fail(HTN_ERR_OOM, "device table allocation failed");
    if (h2d_on_stream(d, host.data(), host.size() * sizeof(T), ctx->stream) != cudaSuccess) {  // do not cache a table that never arrived
      cudaFree(d);
      return ctx->fail(HTN_ERR_CUDA, "device table upload failed");
    }
  }
  owner->devtables[key] = {d, (int)host.size()};
  *dev = d;
  *n = (int)host.size();
  return HTN_OK;
}

int32_t t_copy(const htn_tensor* src, htn_tensor* dst) {
  htn_ctx* ctx = src->ctx;
  if (src->dsize != dst->dsize) return ctx->fail(HTN_ERR_SHAPE, "copy: tensors differ in structure");
  if (src->d == dst->d) return HTN_OK;
  cudaMemcpyAsync(dst->d, src->d, src->dsize * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
  return cuda_rc(ctx, "copy");
}

// mode 0: plain; 1: scale sqrt(d_r/d_l); 2: scale 1/sqrt(d_r/d_l)   (MPS <-> MPST only)
int32_t t_transpose(const htn_tensor* src, htn_tensor* dst, int mode) {
  htn_ctx* ctx = src->ctx;
  const bool ok = (src->kind == HTN_T_MPS && dst->kind == HTN_T_MPST) || (src->kind == HTN_T_MPST && dst->kind == HTN_T_MPS) ||
                  (src->kind == HTN_T_BOND && dst->kind == HTN_T_BOND);
  if (!ok || src->blocks.size() != dst->blocks.size() || src->d == dst->d)
    return ctx->fail(HTN_ERR_INVALID, "transpose: incompatible tensors");
  htn_tensor* owner = const_cast<htn_tensor*>(src);
  TrBlock* dt = nullptr;
  int nt = 0;
  const void* pk = reinterpret_cast<const void*>(static_cast<uintptr_t>(dst->uid));
  auto key = std::make_pair(pk, mode);
  if (owner->devtables.find(key) == owner->devtables.end()) {
    std::vector<TrBlock> tiles;
    for (const Block& sb : src->blocks) {
      const int di = dst->find(sb.lab[0], sb.lab[1], sb.lab[2]);
      if (di < 0) return ctx->fail(HTN_ERR_SHAPE, "transpose: block structures differ");
      const Block& db = dst->blocks[di];
      if (db.rows != sb.cols || db.cols != sb.rows) return ctx->fail(HTN_ERR_SHAPE, "transpose: block shapes differ");
      double scale = 1.0;
      if (mode != 0 && src->kind != HTN_T_BOND) {
        const double w = std::sqrt((double)sdim(src->sym, src->s1.sec[sb.lab[2]]) / sdim(src->sym, src->s0.sec[sb.lab[0]]));
        scale = mode == 1 ? w : 1.0 / w;
      }
      for (int r0 = 0; r0 < sb.rows; r0 += 32)
        for (int c0 = 0; c0 < sb.cols; c0 += 32)
          tiles.push_back(TrBlock{sb.off + (int64_t)r0 * sb.ld + c0, db.off + (int64_t)c0 * db.ld + r0,
                                  std::min(32, sb.rows - r0), std::min(32, sb.cols - c0), sb.ld, db.ld, scale});
    }
    int32_t rc = cached_table(owner, pk, mode, tiles, &dt, &nt);
    if (rc) return rc;
  } else {
    dt = static_cast<TrBlock*>(owner->devtables[key].first);
    nt = owner->devtables[key].second;
  }
  launch_transpose(dt, nt, src->d, dst->d, ctx->stream);
  return cuda_rc(ctx, "transpose");
}

// set all blocks of one MPO level of an environment to zero (mode 0) or the unit matrix (mode 1)
int32_t t_fill_level(htn_tensor* env, int level, int mode) {
  htn_ctx* ctx = env->ctx;
  FillBlock* d = nullptr;
  int n = 0;
  auto key = std::make_pair((const void*)nullptr, 1000 + 2 * level + mode);
  if (env->devtables.find(key) == env->devtables.end()) {
    std::vector<FillBlock> fb;
    for (const Block& b : env->blocks) {
      const bool hit = env->kind == HTN_T_BOND ? level == 0 : b.lab[0] == level;
      if (!hit) continue;
      if (mode == 1 && b.lab[1] != b.lab[2]) return ctx->fail(HTN_ERR_INVALID, "fill_level: identity on an off-diagonal block");
      fb.push_back(FillBlock{b.off, b.rows, b.cols, b.ld, mode});
    }
    int32_t rc = cached_table(env, nullptr, 1000 + 2 * level + mode, fb, &d, &n);
    if (rc) return rc;
  } else {
    d = static_cast<FillBlock*>(env->devtables[key].first);
    n = env->devtables[key].second;
  }
  launch_fill(d, n, env->d, ctx->stream);
  return cuda_rc(ctx, "fill_level");
}

// A = Q R in place (Q overwrites A), one panel per coupled sector; R is a BOND tensor on the
// column space.  MPS: panels = right sectors; MPST: panels = left sectors; BOND: every block.
int32_t t_qr_inplace(htn_tensor* A, htn_tensor* R) {
  htn_ctx* ctx = A->ctx;
  if (R->kind != HTN_T_BOND) return ctx->fail(HTN_ERR_INVALID, "qr: R must be a bond tensor");
  const htn_space& Vc = A->kind == HTN_T_MPS ? A->s1 : A->s0;
  if (R->s0.sec != Vc.sec || R->s0.mult != Vc.mult) return ctx->fail(HTN_ERR_SHAPE, "qr: R space does not match the column space of A");
  QrPanel* d = nullptr;
  int n = 0;
  const void* pk = reinterpret_cast<const void*>(static_cast<uintptr_t>(R->uid));
  auto key = std::make_pair(pk, 2000);
  if (A->devtables.find(key) == A->devtables.end()) {
    std::vector<QrPanel> panels;
    const int gidx = A->kind == HTN_T_MPS ? 2 : 0;
    size_t i = 0;
    while (i < A->blocks.size()) {
      const Block& b0 = A->blocks[i];
      size_t j = i;
      int m = 0;
      if (A->kind == HTN_T_BOND) {
        m = b0.rows;
        j = i + 1;
      } else {
        while (j < A->blocks.size() && A->blocks[j].lab[gidx] == b0.lab[gidx]) {
          m += A->blocks[j].rows;
          ++j;
        }
      }
      const int g = A->kind == HTN_T_BOND ? b0.lab[0] : b0.lab[gidx];
      const Block& rb = R->blocks[g];
      if (b0.cols > 512) return ctx->fail(HTN_ERR_SHAPE, "qr: sector multiplicity above 512 is not supported by the panel kernel");
      if (m < b0.cols) return ctx->fail(HTN_ERR_SHAPE, "qr: a coupled-sector panel has fewer rows than columns (spaces not full rank)");
      panels.push_back(QrPanel{b0.off, rb.off, m, b0.cols, b0.ld, rb.ld});
      i = j;
    }
    int32_t rc = cached_table(A, pk, 2000, panels, &d, &n);
    if (rc) return rc;
  } else {
    d = static_cast<QrPanel*>(A->devtables[key].first);
    n = A->devtables[key].second;
  }
  cudaMemsetAsync(R->d, 0, R->dsize * sizeof(double), ctx->stream);
  int max_m = 0;  // panel heights: group sizes (host side of the cached table)
  {
    const int gidx = A->kind == HTN_T_MPS ? 2 : 0;
    size_t i = 0;
    while (i < A->blocks.size()) {
      size_t j = i;
      int m = 0;
      if (A->kind == HTN_T_BOND) {
        m = A->blocks[i].rows;
        j = i + 1;
      } else {
        while (j < A->blocks.size() && A->blocks[j].lab[gidx] == A->blocks[i].lab[gidx]) m += A->blocks[j++].rows;
      }
      max_m = std::max(max_m, m);
      i = j;
    }
  }
  launch_qr(d, n, max_m, A->d, R->d, ctx->d_status, ctx->stream);
  return cuda_rc(ctx, "qr");
}

void htn_drop_krylov_graphs(htn_ctx* ctx) {
  for (auto& kv : ctx->kry_graphs) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(kv.second));
  ctx->kry_graphs.clear();
}

int32_t ensure_krylov(htn_ctx* ctx, int64_t nvec, int64_t stride, int64_t nchunks) {
  const int64_t need = nvec * stride;
  if (ctx->kry_cap < need) {
    cudaStreamSynchronize(ctx->stream);
    htn_drop_krylov_graphs(ctx);
    if (ctx->kry_V) cudaFree(ctx->kry_V);
    ctx->kry_V = nullptr;
    ctx->kry_cap = 0;
    if (cudaMalloc(&ctx->kry_V, need * sizeof(double)) != cudaSuccess) return ctx->fail(HTN_ERR_OOM, "Krylov basis allocation failed");
    ctx->kry_cap = need;
  }
  const int64_t pneed = (nchunks + 1) * MD_MAXVEC_HOST;
  if (ctx->kry_partial_cap < pneed) {
    cudaStreamSynchronize(ctx->stream);
    htn_drop_krylov_graphs(ctx);
    if (ctx->kry_partial) cudaFree(ctx->kry_partial);
    ctx->kry_partial = nullptr;
    ctx->kry_partial_cap = 0;
    if (cudaMalloc(&ctx->kry_partial, pneed * sizeof(double)) != cudaSuccess) return ctx->fail(HTN_ERR_OOM, "Krylov scratch allocation failed");
    ctx->kry_partial_cap = pneed;
  }
  return HTN_OK;
}

// device scalar slots inside ctx->kry_scal
enum { S_H = 0, S_H2 = 64, S_BETA = 128, S_NRM = 129, S_Y = 192, S_TMP = 300, S_A1 = 512, S_A2 = 576, S_B = 640, S_Z = 1024 };

int32_t t_dot_dev(const htn_tensor* like, const double* x, const double* y, double* out_dev) {
  htn_ctx* ctx = like->ctx;
  int32_t rc = ensure_krylov(ctx, 0, 0, like->nchunks);
  if (rc) return rc;
  if (like->nchunks == 0) {
    cudaMemsetAsync(out_dev, 0, sizeof(double), ctx->stream);
    return HTN_OK;
  }
  launch_multidot(like->dblocks, like->dchunks, like->nchunks, x, 0, 1, y, ctx->kry_partial, out_dev, ctx->stream);
  return cuda_rc(ctx, "dot");
}

int32_t t_dot_host(const htn_tensor* like, const double* x, const double* y, double* out) {
  htn_ctx* ctx = like->ctx;
  int32_t rc = t_dot_dev(like, x, y, ctx->kry_scal + S_TMP);
  if (rc) return rc;
  cudaMemcpyAsync(ctx->kry_scal_host + S_TMP, ctx->kry_scal + S_TMP, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return cuda_rc(ctx, "dot readback");
  *out = ctx->kry_scal_host[S_TMP];
  return HTN_OK;
}

// x <- x / ||x||   (norm stays on the device)
int32_t t_normalize(const htn_tensor* like, double* x) {
  htn_ctx* ctx = like->ctx;
  int32_t rc = t_dot_dev(like, x, x, ctx->kry_scal + S_NRM);
  if (rc) return rc;
  launch_scale_dev(x, ctx->kry_scal + S_NRM, 2, x, like->dsize, ctx->stream);
  return cuda_rc(ctx, "normalize");
}

// ---- small dense symmetric eigenproblem (cyclic Jacobi) on the host ----------------------
static void jacobi_eigh(int n, std::vector<double>& A, std::vector<double>& V, std::vector<double>& w) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = A[(size_t)i * n + i];
}

// Lowest eigenpair of the symmetric operator `apply`: Lanczos with full CGS2 re-orthogonalisation and THICK
// restarts (Krylov-Schur form, as KrylovKit's `eigsolve(.., Lanczos(krylovdim))` does): when the basis is full
// the lowest `keep` Ritz vectors are kept, the projected matrix becomes diag(theta) plus one coupling row to the
// residual vector, and the recurrence continues from there.  x0: start vector; x_out: eigenvector (unit norm,
// <x0, x_out> >= 0).  x0 may alias x_out.
// Semi-eager: the coefficients of every step are parked in device arrays and read back (one synchronisation)
// only every LANCZOS_CHECK steps, so the launches queue up instead of paying a host round trip per step.
// One Lanczos step behind the apply: two rounds of classical Gram-Schmidt of w against V[0..j], beta^2 = <w,w>, the three
// coefficients parked for the host, and (scale_after) w <- w / beta.  Twelve tiny launches: at D_red <= 256 they, not the
// apply, set the pace of the eigensolver, so the sequence is captured once per (vector structure, j) into a CUDA graph
// and replayed (HTN_LANCZOS_GRAPH=0 switches back to plain launches).
static void lanczos_ortho_launches(htn_ctx* ctx, const htn_tensor* like, double* V, int64_t n, int j, double* w, bool scale_after) {
  double* sc = ctx->kry_scal;
  cudaStream_t st = ctx->stream;
  launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H, st);
  launch_multiaxpy(V, n, j + 1, sc + S_H, -1.0, w, n, st);
  launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H2, st);
  launch_multiaxpy(V, n, j + 1, sc + S_H2, -1.0, w, n, st);
  launch_multidot(like->dblocks, like->dchunks, like->nchunks, w, 0, 1, w, ctx->kry_partial, sc + S_BETA, st);
  cudaMemcpyAsync(sc + S_A1 + j, sc + S_H + j, sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(sc + S_A2 + j, sc + S_H2 + j, sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(sc + S_B + j, sc + S_BETA, sizeof(double), cudaMemcpyDeviceToDevice, st);
  if (scale_after) launch_scale_dev(w, sc + S_BETA, 3, w, n, st);  // next basis vector (0 on breakdown)
}

static int32_t lanczos_ortho_step(htn_ctx* ctx, const htn_tensor* like, double* V, int64_t n, int j, double* w, bool scale_after) {
  static int use_graph = -1;
  if (use_graph < 0) {
    const char* e = getenv("HTN_LANCZOS_GRAPH");
    use_graph = e ? atoi(e) : 0;
  }
  if (!use_graph || like->nchunks <= 0) {
    lanczos_ortho_launches(ctx, like, V, n, j, w, scale_after);
    return HTN_OK;
  }
  const std::array<long long, 6> key{(long long)(intptr_t)like->dblocks, (long long)(intptr_t)V, (long long)n, (long long)j,
                                     (long long)like->nchunks, (long long)((intptr_t)like->dchunks ^ (scale_after ? 1 : 0))};
  auto it = ctx->kry_graphs.find(key);
  if (it == ctx->kry_graphs.end()) {
    cudaStream_t st = ctx->stream;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      lanczos_ortho_launches(ctx, like, V, n, j, w, scale_after);
      return HTN_OK;
    }
    lanczos_ortho_launches(ctx, like, V, n, j, w, scale_after);
    if (cudaStreamEndCapture(st, &g) != cudaSuccess || !g || cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
      cudaGetLastError();
      if (g) cudaGraphDestroy(g);
      use_graph = 0;  // capture is not available: plain launches from here on (the captured work was NOT executed)
      lanczos_ortho_launches(ctx, like, V, n, j, w, scale_after);
      return HTN_OK;
    }
    cudaGraphDestroy(g);
    if (ctx->kry_graphs.size() > 4096) htn_drop_krylov_graphs(ctx);
    it = ctx->kry_graphs.emplace(key, (void*)ex).first;
  }
  if (cudaGraphLaunch(static_cast<cudaGraphExec_t>(it->second), ctx->stream) != cudaSuccess) return cuda_rc(ctx, "lanczos graph");
  return HTN_OK;
}

int32_t lanczos_lowest(const htn_tensor* like, const ApplyFn& apply, const double* x0, double* x_out, int krylovdim,
                       double tol, int maxiter, KrylovInfo* info) {
  htn_ctx* ctx = like->ctx;
  const int kd = std::max(2, std::min(krylovdim, 60));
  const int keep = std::max(1, std::min(kd - 2, (3 * kd) / 5));
  const int64_t n = like->dsize;
  int32_t rc = ensure_krylov(ctx, kd + 2 + keep, n, like->nchunks);
  if (rc) return rc;
  double* V = ctx->kry_V;                            // slots 0..kd: basis (+ residual)
  double* xsave = V + (int64_t)(kd + 1) * n;         // copy of x0 for the final sign convention
  double* S = V + (int64_t)(kd + 2) * n;             // `keep` scratch vectors for the restart rotation
  double* sc = ctx->kry_scal;
  double* sh = ctx->kry_scal_host;
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(xsave, x0, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(V, x0, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if ((rc = t_normalize(like, V))) return rc;
  constexpr int LANCZOS_CHECK = 5;
  int applies = 0, m = 0, j0 = 0, k = 0;  // k: number of kept Ritz vectors at the head of the basis
  double theta = 0.0, res = 1e300;
  bool converged = false;
  std::vector<double> thetas, arrow, Hm, Z, ev, y, alphas, betas;
  for (int cycle = 0; cycle < maxiter && !converged; ++cycle) {
    bool full = false;
    for (int j = j0; j < kd; ++j) {
      double* w = V + (int64_t)(j + 1) * n;
      if ((rc = apply(V + (int64_t)j * n, w))) return rc;
      ++applies;
      const bool check = ((j - j0) % LANCZOS_CHECK) == LANCZOS_CHECK - 1 || j == kd - 1;
      if ((rc = lanczos_ortho_step(ctx, like, V, n, j, w, !check))) return rc;
      if (!check) continue;
      cudaMemcpyAsync(sh + S_A1, sc + S_A1, 192 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_rc(ctx, "lanczos");
      m = j + 1;
      alphas.assign(m, 0.0);
      betas.assign(m, 0.0);
      for (int i = k; i < m; ++i) {
        alphas[i] = sh[S_A1 + i] + sh[S_A2 + i];
        betas[i] = std::sqrt(std::max(sh[S_B + i], 0.0));
      }
      for (int i = k; i < m; ++i)
        if (betas[i] < 1e-14) {  // invariant subspace reached at step i
          m = i + 1;
          break;
        }
      // projected matrix: diag(theta_0..theta_{k-1}) + coupling row/column k + tridiagonal tail
      Hm.assign((size_t)m * m, 0.0);
      for (int i = 0; i < k && i < m; ++i) {
        Hm[(size_t)i * m + i] = thetas[i];
        if (k < m) Hm[(size_t)i * m + k] = Hm[(size_t)k * m + i] = arrow[i];
      }
      for (int i = k; i < m; ++i) {
        Hm[(size_t)i * m + i] = alphas[i];
        if (i + 1 < m) Hm[(size_t)i * m + i + 1] = Hm[(size_t)(i + 1) * m + i] = betas[i];
      }
      jacobi_eigh(m, Hm, Z, ev);
      int lo = 0;
      for (int i = 1; i < m; ++i)
        if (ev[i] < ev[lo]) lo = i;
      theta = ev[lo];
      y.assign(m, 0.0);
      for (int i = 0; i < m; ++i) y[i] = Z[(size_t)i * m + lo];
      const double bm = betas[m - 1];
      res = std::fabs(bm * y[m - 1]);
      if (res < tol || bm < 1e-14 || m < j + 1) break;
      if (j == kd - 1) {
        full = true;
        break;
      }
      launch_scale_dev(w, sc + S_BETA, 3, w, n, st);
    }
    converged = res < tol || (m > 0 && m < kd && !full && res < 1e299 && betas[m - 1] < 1e-14);
    if (converged || !full || cycle == maxiter - 1) break;
    // ---- thick restart: keep the lowest `keep` Ritz vectors + the residual direction ----
    std::vector<int> order(m);
    for (int i = 0; i < m; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a2, int b2) { return ev[a2] < ev[b2]; });
    const int kn = std::min(keep, m - 1);
    const double bm = betas[m - 1];
    thetas.assign(kn, 0.0);
    arrow.assign(kn, 0.0);
    for (int i = 0; i < kn; ++i) {
      thetas[i] = ev[order[i]];
      arrow[i] = bm * Z[(size_t)(m - 1) * m + order[i]];
      for (int r = 0; r < m; ++r) sh[S_Z + i * m + r] = Z[(size_t)r * m + order[i]];
    }
    cudaMemcpyAsync(sc + S_Z, sh + S_Z, (size_t)kn * m * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(S, 0, (size_t)kn * n * sizeof(double), st);
    for (int i = 0; i < kn; ++i) launch_multiaxpy(V, n, m, sc + S_Z + i * m, 1.0, S + (int64_t)i * n, n, st);
    launch_scale_dev(V + (int64_t)m * n, sc + S_BETA, 3, V + (int64_t)kn * n, n, st);  // residual -> slot kn
    cudaMemcpyAsync(V, S, (size_t)kn * n * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_rc(ctx, "lanczos restart");
    k = kn;
    j0 = kn;
  }
  // Ritz vector of the last evaluation -> V[0]
  if (m == 0) return ctx->fail(HTN_ERR_INVALID, "lanczos: no iteration performed");
  for (int i = 0; i < m; ++i) sh[S_Y + i] = y[i];
  cudaMemcpyAsync(sc + S_Y, sh + S_Y, m * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(S, 0, n * sizeof(double), st);
  launch_multiaxpy(V, n, m, sc + S_Y, 1.0, S, n, st);
  if ((rc = t_normalize(like, S))) return rc;
  // sign convention: positive overlap with the start vector
  double ov = 0.0;
  if ((rc = t_dot_host(like, xsave, S, &ov))) return rc;
  launch_axpby(ov < 0 ? -1.0 : 1.0, S, 0.0, x_out, n, st);
  if ((rc = cuda_rc(ctx, "lanczos"))) return rc;
  converged = res < tol;
  if (info) {
    info->value = theta;
    info->residual = res;
    info->applies = applies;
    info->converged = converged ? 1 : 0;
  }
  return converged ? HTN_OK : HTN_NOT_CONVERGED;
}

// Restarted GMRES for apply(x) = b, x0 = 0 (same recurrence as oracle/krylov.py:gmres).
int32_t gmres_solve(const htn_tensor* like, const ApplyFn& apply, const double* b, double* x, int krylovdim, double tol,
                    int maxiter, KrylovInfo* info) {
  htn_ctx* ctx = like->ctx;
  krylovdim = std::max(2, std::min(krylovdim, 60));
  const int64_t n = like->dsize;
  int32_t rc = ensure_krylov(ctx, krylovdim + 2, n, like->nchunks);
  if (rc) return rc;
  double* V = ctx->kry_V;
  double* tmp = V + (int64_t)(krylovdim + 1) * n;
  double* sc = ctx->kry_scal;
  double* sh = ctx->kry_scal_host;
  cudaStream_t st = ctx->stream;
  cudaMemsetAsync(x, 0, n * sizeof(double), st);
  double bnorm2 = 0.0;
  if ((rc = t_dot_host(like, b, b, &bnorm2))) return rc;
  const double bnorm = std::sqrt(std::max(bnorm2, 0.0));
  int applies = 0;
  double res = 0.0;
  bool converged = bnorm == 0.0;
  for (int restart = 0; restart < maxiter && !converged; ++restart) {
    // r = b - A x
    if (restart == 0) {
      cudaMemcpyAsync(V, b, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
    } else {
      if ((rc = apply(x, tmp))) return rc;
      ++applies;
      cudaMemcpyAsync(V, b, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
      launch_axpby(-1.0, tmp, 1.0, V, n, st);
    }
    double beta2 = 0.0;
    if ((rc = t_dot_host(like, V, V, &beta2))) return rc;
    const double beta = std::sqrt(std::max(beta2, 0.0));
    res = beta / bnorm;
    if (res < tol) {
      converged = true;
      break;
    }
    launch_axpby(1.0 / beta, V, 0.0, V, n, st);
    // Arnoldi with CGS2; least squares by Givens rotations
    std::vector<double> H((size_t)(krylovdim + 1) * krylovdim, 0.0), cs(krylovdim), sn(krylovdim), gvec(krylovdim + 1, 0.0);
    gvec[0] = beta;
    int m = 0;
    for (int j = 0; j < krylovdim; ++j) {
      double* w = V + (int64_t)(j + 1) * n;
      if ((rc = apply(V + (int64_t)j * n, w))) return rc;
      ++applies;
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H, st);
      launch_multiaxpy(V, n, j + 1, sc + S_H, -1.0, w, n, st);
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H2, st);
      launch_multiaxpy(V, n, j + 1, sc + S_H2, -1.0, w, n, st);
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, w, 0, 1, w, ctx->kry_partial, sc + S_BETA, st);
      cudaMemcpyAsync(sh, sc, 130 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_rc(ctx, "gmres");
      const int ldh = krylovdim;
      for (int i = 0; i <= j; ++i) H[(size_t)i * ldh + j] = sh[S_H + i] + sh[S_H2 + i];
      const double hn = std::sqrt(std::max(sh[S_BETA], 0.0));
      H[(size_t)(j + 1) * ldh + j] = hn;
      // apply previous rotations to the new column
      for (int i = 0; i < j; ++i) {
        const double t0 = cs[i] * H[(size_t)i * ldh + j] + sn[i] * H[(size_t)(i + 1) * ldh + j];
        H[(size_t)(i + 1) * ldh + j] = -sn[i] * H[(size_t)i * ldh + j] + cs[i] * H[(size_t)(i + 1) * ldh + j];
        H[(size_t)i * ldh + j] = t0;
      }
      const double h0 = H[(size_t)j * ldh + j], h1 = H[(size_t)(j + 1) * ldh + j];
      const double rr = std::hypot(h0, h1);
      cs[j] = rr > 0 ? h0 / rr : 1.0;
      sn[j] = rr > 0 ? h1 / rr : 0.0;
      H[(size_t)j * ldh + j] = rr;
      H[(size_t)(j + 1) * ldh + j] = 0.0;
      gvec[j + 1] = -sn[j] * gvec[j];
      gvec[j] = cs[j] * gvec[j];
      m = j + 1;
      res = std::fabs(gvec[j + 1]) / bnorm;
      if (res < tol || hn < 1e-14) break;
      launch_scale_dev(w, sc + S_BETA, 2, w, n, st);
    }
    // back substitution, x += V y
    std::vector<double> y(m, 0.0);
    for (int i = m - 1; i >= 0; --i) {
      double s = gvec[i];
      for (int k = i + 1; k < m; ++k) s -= H[(size_t)i * krylovdim + k] * y[k];
      y[i] = s / H[(size_t)i * krylovdim + i];
    }
    for (int i = 0; i < m; ++i) sh[S_Y + i] = y[i];
    cudaMemcpyAsync(sc + S_Y, sh + S_Y, m * sizeof(double), cudaMemcpyHostToDevice, st);
    launch_multiaxpy(V, n, m, sc + S_Y, 1.0, x, n, st);
    cudaStreamSynchronize(st);
    if (res < tol) converged = true;
  }
  if ((rc = cuda_rc(ctx, "gmres"))) return rc;
  if (info) {
    info->value = 0.0;
    info->residual = res;
    info->applies = applies;
    info->converged = converged ? 1 : 0;
  }
  return converged ? HTN_OK : HTN_NOT_CONVERGED;
}

// ---- dominant eigenpair of a small real upper-Hessenberg matrix on the host ------------------
// Eigenvalues by the single-shift QR algorithm in complex arithmetic (Wilkinson shifts, Givens
// rotations, deflation); the eigenvector of the eigenvalue of largest modulus (real for a transfer
// map) by inverse iteration.  m <= 60.
static bool hessenberg_eigenvalues(int m, const std::vector<double>& Hin, int ld, std::vector<std::complex<double>>& ev) {
  typedef std::complex<double> cd;
  std::vector<cd> H((size_t)m * m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) H[(size_t)i * m + j] = (j >= i - 1) ? Hin[(size_t)i * ld + j] : 0.0;
  ev.assign(m, cd(0.0));
  int hi = m - 1, iter = 0;
  std::vector<cd> cs(m), sn(m);
  while (hi >= 0) {
    if (hi == 0) {
      ev[0] = H[0];
      break;
    }
    int l = hi;
    for (; l > 0; --l) {
      const double sub = std::abs(H[(size_t)l * m + l - 1]);
      const double scale = std::abs(H[(size_t)l * m + l]) + std::abs(H[(size_t)(l - 1) * m + l - 1]);
      if (sub <= 1e-16 * (scale > 0 ? scale : 1.0)) {
        H[(size_t)l * m + l - 1] = 0.0;
        break;
      }
    }
    if (l == hi) {
      ev[hi] = H[(size_t)hi * m + hi];
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 300) return false;
    // Wilkinson shift: eigenvalue of the trailing 2x2 closer to H[hi][hi]
    const cd a = H[(size_t)(hi - 1) * m + hi - 1], b = H[(size_t)(hi - 1) * m + hi], c = H[(size_t)hi * m + hi - 1],
             d = H[(size_t)hi * m + hi];
    const cd tr = a + d, det = a * d - b * c;
    const cd disc = std::sqrt(tr * tr - 4.0 * det);
    const cd l1 = 0.5 * (tr + disc), l2 = 0.5 * (tr - disc);
    cd mu = std::abs(l1 - d) < std::abs(l2 - d) ? l1 : l2;
    if (iter % 11 == 10) mu += cd(std::abs(c), 0.0);  // exceptional shift
    for (int i = l; i <= hi; ++i) H[(size_t)i * m + i] -= mu;
    for (int k = l; k < hi; ++k) {
      const cd x = H[(size_t)k * m + k], y = H[(size_t)(k + 1) * m + k];
      const double r = std::sqrt(std::norm(x) + std::norm(y));
      if (r == 0.0) {
        cs[k] = 1.0;
        sn[k] = 0.0;
        continue;
      }
      cs[k] = x / r;
      sn[k] = y / r;
      for (int j = k; j <= hi; ++j) {
        const cd u = H[(size_t)k * m + j], v = H[(size_t)(k + 1) * m + j];
        H[(size_t)k * m + j] = std::conj(cs[k]) * u + std::conj(sn[k]) * v;
        H[(size_t)(k + 1) * m + j] = -sn[k] * u + cs[k] * v;
      }
    }
    for (int k = l; k < hi; ++k) {
      const int top = std::min(k + 2, hi);
      for (int i = l; i <= top; ++i) {
        const cd u = H[(size_t)i * m + k], v = H[(size_t)i * m + k + 1];
        H[(size_t)i * m + k] = u * cs[k] + v * sn[k];
        H[(size_t)i * m + k + 1] = -u * std::conj(sn[k]) + v * std::conj(cs[k]);
      }
    }
    for (int i = l; i <= hi; ++i) H[(size_t)i * m + i] += mu;
  }
  return true;
}

static bool small_dominant_eig(int m, const std::vector<double>& H, int ld, double* theta, std::vector<double>& y) {
  std::vector<std::complex<double>> ev;
  if (!hessenberg_eigenvalues(m, H, ld, ev)) return false;
  int k = 0;
  for (int i = 1; i < m; ++i)
    if (std::abs(ev[i]) > std::abs(ev[k])) k = i;
  if (std::fabs(ev[k].imag()) > 1e-8 * std::abs(ev[k])) return false;  // not a transfer-map fixed point
  const double lam = ev[k].real();
  const double sigma = lam * (1.0 + 1e-10) + 1e-300;
  y.assign(m, 1.0 / std::sqrt((double)m));
  for (int it = 0; it < 4; ++it) {
    std::vector<double> A((size_t)m * m), b(y);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) A[(size_t)i * m + j] = H[(size_t)i * ld + j] - (i == j ? sigma : 0.0);
    for (int c = 0; c < m; ++c) {
      int piv = c;
      for (int r = c + 1; r < m; ++r)
        if (std::fabs(A[(size_t)r * m + c]) > std::fabs(A[(size_t)piv * m + c])) piv = r;
      if (std::fabs(A[(size_t)piv * m + c]) < 1e-300) A[(size_t)piv * m + c] = 1e-300;
      if (piv != c) {
        for (int j = 0; j < m; ++j) std::swap(A[(size_t)piv * m + j], A[(size_t)c * m + j]);
        std::swap(b[piv], b[c]);
      }
      for (int r = c + 1; r < m; ++r) {
        const double f = A[(size_t)r * m + c] / A[(size_t)c * m + c];
        if (f == 0.0) continue;
        for (int j = c; j < m; ++j) A[(size_t)r * m + j] -= f * A[(size_t)c * m + j];
        b[r] -= f * b[c];
      }
    }
    for (int i = m - 1; i >= 0; --i) {
      double sacc = b[i];
      for (int j = i + 1; j < m; ++j) sacc -= A[(size_t)i * m + j] * b[j];
      b[i] = sacc / A[(size_t)i * m + i];
    }
    double n = 0.0;
    for (int i = 0; i < m; ++i) n += b[i] * b[i];
    n = std::sqrt(n);
    if (!(n > 0.0) || !std::isfinite(n)) return false;
    for (int i = 0; i < m; ++i) y[i] = b[i] / n;
  }
  if (y[0] < 0.0)
    for (int i = 0; i < m; ++i) y[i] = -y[i];
  *theta = lam;
  return true;
}

}  // namespace htn

// test hook (host only): dominant eigenpair of a small upper-Hessenberg matrix H[m][m] (row-major)
extern "C" int32_t htn_test_hessenberg_dominant(int32_t m, const double* H, double* theta, double* y) {
  if (m <= 0 || m > 60 || !H || !theta || !y) return HTN_ERR_INVALID;
  std::vector<double> h(H, H + (size_t)m * m), v;
  if (!htn::small_dominant_eig(m, h, m, theta, v)) return HTN_NOT_CONVERGED;
  for (int i = 0; i < m; ++i) y[i] = v[i];
  return HTN_OK;
}

namespace htn {

// Dominant eigenvector (largest |lambda|, assumed real and simple: the fixed point of a transfer map) by
// Arnoldi with CGS2 and explicit restarts (KrylovKit `eigsolve(f, x0, 1, :LM, Arnoldi(krylovdim, tol))`
// as used by MPSKit's uniform_leftorth!/rightorth!).  x: start vector in, eigenvector (unit norm) out.
int32_t arnoldi_dominant(const htn_tensor* like, const ApplyFn& apply, double* x, int krylovdim, double tol, int maxiter,
                         KrylovInfo* info) {
  htn_ctx* ctx = like->ctx;
  krylovdim = std::max(2, std::min(krylovdim, 60));
  const int64_t n = like->dsize;
  int32_t rc = ensure_krylov(ctx, krylovdim + 2, n, like->nchunks);
  if (rc) return rc;
  double* V = ctx->kry_V;
  double* sc = ctx->kry_scal;
  double* sh = ctx->kry_scal_host;
  cudaStream_t st = ctx->stream;
  cudaMemcpyAsync(V, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if ((rc = t_normalize(like, V))) return rc;
  int applies = 0;
  double theta = 0.0, res = 1e300;
  bool converged = false;
  for (int restart = 0; restart < maxiter && !converged; ++restart) {
    const int ldh = krylovdim;
    std::vector<double> H((size_t)(krylovdim + 1) * krylovdim, 0.0), y;
    int m = 0;
    for (int j = 0; j < krylovdim; ++j) {
      double* w = V + (int64_t)(j + 1) * n;
      if ((rc = apply(V + (int64_t)j * n, w))) return rc;
      ++applies;
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H, st);
      launch_multiaxpy(V, n, j + 1, sc + S_H, -1.0, w, n, st);
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, V, n, j + 1, w, ctx->kry_partial, sc + S_H2, st);
      launch_multiaxpy(V, n, j + 1, sc + S_H2, -1.0, w, n, st);
      launch_multidot(like->dblocks, like->dchunks, like->nchunks, w, 0, 1, w, ctx->kry_partial, sc + S_BETA, st);
      cudaMemcpyAsync(sh, sc, 130 * sizeof(double), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return cuda_rc(ctx, "arnoldi");
      for (int i = 0; i <= j; ++i) H[(size_t)i * ldh + j] = sh[S_H + i] + sh[S_H2 + i];
      const double hn = std::sqrt(std::max(sh[S_BETA], 0.0));
      H[(size_t)(j + 1) * ldh + j] = hn;
      m = j + 1;
      if (hn < 1e-14 || j == krylovdim - 1) {
        // (no eager evaluation: the small problem is only solved once the space is complete)
        const bool ok = small_dominant_eig(m, H, ldh, &theta, y);
        res = ok ? std::fabs(hn * y[m - 1]) : 1e300;
        break;
      }
      launch_scale_dev(w, sc + S_BETA, 2, w, n, st);
    }
    if (!(res < 1e299)) {  // the small eigenproblem did not settle: give up, the caller keeps its vector
      if (info) {
        info->value = theta;
        info->residual = res;
        info->applies = applies;
        info->converged = 0;
      }
      return HTN_NOT_CONVERGED;
    }
    for (int i = 0; i < m; ++i) sh[S_Y + i] = y[i];
    cudaMemcpyAsync(sc + S_Y, sh + S_Y, m * sizeof(double), cudaMemcpyHostToDevice, st);
    double* acc = V + (int64_t)m * n;
    cudaMemsetAsync(acc, 0, n * sizeof(double), st);
    launch_multiaxpy(V, n, m, sc + S_Y, 1.0, acc, n, st);
    if ((rc = t_normalize(like, acc))) return rc;
    cudaMemcpyAsync(V, acc, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaStreamSynchronize(st);
    converged = res < tol;
  }
  cudaMemcpyAsync(x, V, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
  if ((rc = cuda_rc(ctx, "arnoldi"))) return rc;
  if (info) {
    info->value = theta;
    info->residual = res;
    info->applies = applies;
    info->converged = converged ? 1 : 0;
  }
  return converged ? HTN_OK : HTN_NOT_CONVERGED;
}

}  // namespace htn
