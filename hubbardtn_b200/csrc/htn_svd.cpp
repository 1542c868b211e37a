// Per-sector truncated SVD of a two-site tensor (device one-sided Jacobi + host truncation logic).
//
// Replaces TensorKit `tsvd!(ac2; trunc = truncbelow(cut) | truncdim(D))` as called by MPSKit's IDMRG2
// and `changebonds(.., SvdCut)` (/root/reference/src/HubbardFunctions.jl:1010,1013,1018,1363-1365;
// SURVEY.md 8(a) a9).  Same conventions as oracle/twosite.py:tsvd:
//   M_m[(s1,l),(r,s2)] = sqrt(d_r/d_m) x2[l,s1,m,s2,r] = U S V^T  per middle sector m,
//   AL[l,s1,m] = rows of U, C[m] = diag(S), AR[m,s2,r] = columns of V^T / sqrt(d_r/d_m);
// truncation is global over all sectors: keep sigma >= cut * ||x2|| (and > 1e-14 ||x2||), at most
// `maxdim` multiplets (largest first); maxdim < 0 caps the FULL dimension sum_c dim(c) n_c at |maxdim| (TensorKit truncdim).  Only the singular values travel to the host.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "htn_linalg.hpp"

using namespace htn;

namespace {
struct PanelInfo {
  int m;  // index into x2->mid
  std::vector<std::pair<int, int>> rows, cols;  // (s1,l) / (r,s2), sorted
  std::map<std::pair<int, int>, int> ro, co;
  int nrow = 0, ncol = 0;
  bool u_in_g = false;  // G = M^T (k = ncol vectors of length nrow)
  int k = 0, len = 0, ldg = 0, ldq = 0;
  int64_t offG = 0, offQ = 0, offS = 0;
};
}  // namespace

extern "C" int32_t htn_tsvd(const htn_tensor* x2, double cut, int32_t maxdim, htn_space** Vm_out, htn_tensor** AL_out,
                            htn_tensor** C_out, htn_tensor** AR_out, double* discarded_weight, int32_t* kept) {
  if (!x2 || !Vm_out || !AL_out || !C_out || !AR_out) return HTN_ERR_INVALID;
  htn_ctx* ctx = x2->ctx;
  if (x2->kind != HTN_T_MPS2) return ctx->fail(HTN_ERR_INVALID, "tsvd: x2 must be a two-site tensor");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int sym = x2->sym;
  *Vm_out = nullptr;
  *AL_out = *C_out = *AR_out = nullptr;
  double *dG = nullptr, *dQ = nullptr, *dG2 = nullptr, *dQ2 = nullptr, *dS = nullptr;
  SvdPanel* dP = nullptr;
  TrBlock *dT1 = nullptr, *dT2 = nullptr;
  htn_space* Vm = nullptr;
  htn_tensor *AL = nullptr, *Cb = nullptr, *AR = nullptr;
  auto cleanup = [&]() {
    cudaFree(dG);
    cudaFree(dQ);
    cudaFree(dG2);
    cudaFree(dQ2);
    cudaFree(dS);
    cudaFree(dP);
    cudaFree(dT1);
    cudaFree(dT2);
  };
  auto fail = [&](int32_t code, const std::string& msg) {
    cleanup();
    if (AL) htn_tensor_destroy(AL);
    if (Cb) htn_tensor_destroy(Cb);
    if (AR) htn_tensor_destroy(AR);
    if (Vm) htn_space_destroy(Vm);
    return ctx->fail(code, msg);
  };
  const bool dbg = getenv("HTN_DEBUG_SYNC") != nullptr;
  auto dbg_sync = [&](const char* what) {
    if (!dbg) return;
    cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, "tsvd debug: %s -> %s\n", what, cudaGetErrorString(e));
  };
  dbg_sync("entry (kernels queued before tsvd)");
  try {
    // ---- panels ----
    std::map<int, PanelInfo> pm;
    for (const Block& b : x2->blocks) {
      PanelInfo& p = pm[b.lab[2]];
      p.m = b.lab[2];
      p.rows.push_back({b.lab[1], b.lab[0]});
      p.cols.push_back({b.lab[4], b.lab[3]});
    }
    std::vector<PanelInfo*> panels;
    int64_t totG = 0, totQ = 0, totS = 0;
    for (auto& kv : pm) {
      PanelInfo& p = kv.second;
      std::sort(p.rows.begin(), p.rows.end());
      p.rows.erase(std::unique(p.rows.begin(), p.rows.end()), p.rows.end());
      std::sort(p.cols.begin(), p.cols.end());
      p.cols.erase(std::unique(p.cols.begin(), p.cols.end()), p.cols.end());
      for (auto& r : p.rows) {
        p.ro[r] = p.nrow;
        p.nrow += x2->s0.mult[r.second];
      }
      for (auto& c : p.cols) {
        p.co[c] = p.ncol;
        p.ncol += x2->s1.mult[c.first];
      }
      if (p.nrow == 0 || p.ncol == 0) continue;
      p.u_in_g = p.nrow >= p.ncol;
      p.k = std::min(p.nrow, p.ncol);
      p.len = std::max(p.nrow, p.ncol);
      if (p.k > 1024) return fail(HTN_ERR_SHAPE, "tsvd: coupled block with more than 1024 singular values is not supported");
      p.ldg = even_up(p.len);
      p.ldq = even_up(p.k);
      p.offG = totG;
      p.offQ = totQ;
      p.offS = totS;
      totG = align_up(totG + (int64_t)p.k * p.ldg, 16);
      totQ = align_up(totQ + (int64_t)p.k * p.ldq, 16);
      totS += p.k;
      panels.push_back(&p);
    }
    if (panels.empty()) return fail(HTN_ERR_SHAPE, "tsvd: empty tensor");
    if (cudaMalloc(&dG, totG * 8) != cudaSuccess || cudaMalloc(&dQ, totQ * 8) != cudaSuccess ||
        cudaMalloc(&dG2, totG * 8) != cudaSuccess || cudaMalloc(&dQ2, totQ * 8) != cudaSuccess ||
        cudaMalloc(&dS, std::max<int64_t>(totS, 1) * 8) != cudaSuccess)
      return fail(HTN_ERR_OOM, "tsvd: workspace allocation failed");
    cudaMemsetAsync(dG, 0, totG * 8, st);
    // ---- gather x2 blocks into the panels (weighted; transposed when G = M^T) ----
    std::vector<TrBlock> tT, tC;  // transposing / straight tiles
    for (const Block& b : x2->blocks) {
      const PanelInfo& p = pm[b.lab[2]];
      if (p.k == 0) continue;
      const double w = std::sqrt((double)sdim(sym, x2->s1.sec[b.lab[4]]) / sdim(sym, x2->mid[b.lab[2]]));
      const int r0 = p.ro.at({b.lab[1], b.lab[0]}), c0 = p.co.at({b.lab[4], b.lab[3]});
      for (int rr = 0; rr < b.rows; rr += 32)
        for (int cc = 0; cc < b.cols; cc += 32) {
          TrBlock t{};
          t.soff = b.off + (int64_t)rr * b.ld + cc;
          t.rows = std::min(32, b.rows - rr);
          t.cols = std::min(32, b.cols - cc);
          t.lds = b.ld;
          t.ldd = p.ldg;
          t.scale = w;
          if (p.u_in_g) {  // G[c0+cc+c][r0+rr+r] = w x[r][c]
            t.doff = p.offG + (int64_t)(c0 + cc) * p.ldg + (r0 + rr);
            tT.push_back(t);
          } else {  // G[r0+rr+r][c0+cc+c]
            t.doff = p.offG + (int64_t)(r0 + rr) * p.ldg + (c0 + cc);
            tC.push_back(t);
          }
        }
    }
    auto upload_tiles = [&](const std::vector<TrBlock>& v, TrBlock** d) -> bool {
      if (v.empty()) return true;
      if (cudaMalloc(d, v.size() * sizeof(TrBlock)) != cudaSuccess) return false;
      return h2d_on_stream(*d, v.data(), v.size() * sizeof(TrBlock), st) == cudaSuccess;
    };
    if (!upload_tiles(tT, &dT1) || !upload_tiles(tC, &dT2)) return fail(HTN_ERR_OOM, "tsvd: table allocation failed");
    if (dbg) {
      fprintf(stderr, "tsvd debug: x2 dsize %lld blocks %zu totG %lld totQ %lld tiles T %zu C %zu\n", (long long)x2->dsize,
              x2->blocks.size(), (long long)totG, (long long)totQ, tT.size(), tC.size());
      for (PanelInfo* p : panels)
        fprintf(stderr, "   panel m=%d nrow=%d ncol=%d k=%d len=%d ldg=%d offG=%lld u_in_g=%d\n", p->m, p->nrow, p->ncol, p->k, p->len,
                p->ldg, (long long)p->offG, (int)p->u_in_g);
      long long maxs = 0, maxd = 0;
      for (const TrBlock& t : tT) {
        maxs = std::max(maxs, t.soff + (long long)(t.rows - 1) * t.lds + t.cols);
        maxd = std::max(maxd, t.doff + (long long)(t.cols - 1) * t.ldd + t.rows);
      }
      fprintf(stderr, "   transposing tiles: max src end %lld (dsize %lld)  max dst end %lld (totG %lld)\n", maxs, (long long)x2->dsize, maxd,
              (long long)totG);
    }
    launch_transpose(dT1, (int)tT.size(), x2->d, dG, st);
    dbg_sync("gather (transposing tiles)");
    launch_copy2d(dT2, (int)tC.size(), x2->d, dG, st);
    dbg_sync("gather (straight tiles)");
    // ---- Jacobi ----
    std::vector<SvdPanel> sp;
    for (PanelInfo* p : panels) sp.push_back(SvdPanel{p->offG, p->offQ, p->offS, p->k, p->len, p->ldg, p->ldq, p->u_in_g ? 1 : 0, 0});
    if (cudaMalloc(&dP, sp.size() * sizeof(SvdPanel)) != cudaSuccess) return fail(HTN_ERR_OOM, "tsvd: table allocation failed");
    h2d_on_stream(dP, sp.data(), sp.size() * sizeof(SvdPanel), st);
    const int svd_rc = launch_svd(dP, sp.data(), (int)sp.size(), dG, dQ, dG2, dQ2, dS, st);
    if (svd_rc < 0) return fail(HTN_ERR_CUDA, "tsvd: Jacobi launch failed");
    if (svd_rc > 0) return fail(HTN_ERR_INVALID, "tsvd: Jacobi sweeps did not converge");
    if (dbg) {
      dbg_sync("jacobi");
      for (PanelInfo* p : panels) fprintf(stderr, "   panel m=%d k=%d len=%d u_in_g=%d\n", p->m, p->k, p->len, (int)p->u_in_g);
    }
    std::vector<double> sig(totS);
    cudaMemcpyAsync(sig.data(), dS, totS * 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(HTN_ERR_CUDA, std::string("tsvd: ") + cudaGetErrorString(cudaGetLastError()));
    // ---- global truncation (oracle/twosite.py:tsvd) ----
    struct SV {
      double s;
      int panel;
    };
    std::vector<SV> all;
    double nrm2 = 0.0;
    for (size_t pi = 0; pi < panels.size(); ++pi)
      for (int i = 0; i < panels[pi]->k; ++i) {
        const double s = sig[panels[pi]->offS + i];
        all.push_back(SV{s, (int)pi});
        nrm2 += sdim(sym, x2->mid[panels[pi]->m]) * s * s;
      }
    const double nrm = std::sqrt(nrm2);
    std::stable_sort(all.begin(), all.end(), [](const SV& a, const SV& b) { return a.s > b.s; });
    std::vector<int> keep(panels.size(), 0);
    int nkept = 0;
    long long nfull = 0;  // maxdim < 0: cap on the FULL dimension sum_c dim(c) n_c (TensorKit truncdim)
    for (const SV& v : all) {
      if (v.s < cut * nrm || v.s <= 1e-14 * nrm || (maxdim > 0 && nkept >= maxdim)) break;
      const int dm = sdim(sym, x2->mid[panels[v.panel]->m]);
      if (maxdim < 0 && nfull + dm > -(long long)maxdim) break;
      ++keep[v.panel];
      ++nkept;
      nfull += dm;
    }
    if (nkept == 0) return fail(HTN_ERR_INVALID, "tsvd: truncation removed every singular value");
    double disc = 0.0;
    for (size_t pi = 0; pi < panels.size(); ++pi)
      for (int i = keep[pi]; i < panels[pi]->k; ++i) {
        const double s = sig[panels[pi]->offS + i];
        disc += sdim(sym, x2->mid[panels[pi]->m]) * s * s;
      }
    // ---- new middle space and the three factors ----
    std::vector<int32_t> labs, mult;
    for (size_t pi = 0; pi < panels.size(); ++pi)
      if (keep[pi] > 0) {
        const Sector c = x2->mid[panels[pi]->m];
        labs.push_back(c.p);
        labs.push_back(c.q);
        labs.push_back(c.n);
        mult.push_back(keep[pi]);
      }
    int32_t rc = htn_space_create(ctx, sym, (int)mult.size(), labs.data(), mult.data(), &Vm);
    if (rc) return fail(rc, ctx->err);
    htn_space vl = x2->s0, vr = x2->s1;
    if ((rc = htn_tensor_create_mps(ctx, &vl, &x2->legs, Vm, &AL)) || (rc = htn_tensor_create_bond(ctx, Vm, &Cb)) ||
        (rc = htn_tensor_create_mps(ctx, Vm, &x2->legs2, &vr, &AR)))
      return fail(rc, ctx->err);
    std::vector<TrBlock> uT, uC, vC;  // U from G2 (transposing) / from Q2 (transposing); V^T straight
    std::vector<TrBlock> uT_q, vC_g;
    std::vector<double> cpacked(Cb->hsize, 0.0);
    for (size_t pi = 0; pi < panels.size(); ++pi) {
      const int kp = keep[pi];
      if (kp == 0) continue;
      const PanelInfo& p = *panels[pi];
      const Sector cm = x2->mid[p.m];
      int mi = -1;
      for (size_t i = 0; i < Vm->sec.size(); ++i)
        if (Vm->sec[i] == cm) mi = (int)i;
      const Block& cb = Cb->blocks[mi];
      for (int i = 0; i < kp; ++i) cpacked[cb.hoff + (int64_t)i * cb.cols + i] = sig[p.offS + i];
      // AL[l,s1,mi] [n_l x kp] = U[ro:+n_l, :kp];  U[row][i] = (u_in_g ? G2 : Q2)[i][row]
      for (auto& r : p.rows) {
        const Block& ab = AL->blocks[AL->find(r.second, r.first, mi)];
        const int r0 = p.ro.at(r);
        for (int ii = 0; ii < kp; ii += 32)
          for (int cc = 0; cc < ab.rows; cc += 32) {
            TrBlock t{};
            t.rows = std::min(32, kp - ii);        // source rows = singular index
            t.cols = std::min(32, ab.rows - cc);   // source cols = row index inside the panel
            t.scale = 1.0;
            t.ldd = ab.ld;
            t.doff = ab.off + (int64_t)cc * ab.ld + ii;
            if (p.u_in_g) {
              t.soff = p.offG + (int64_t)ii * p.ldg + r0 + cc;
              t.lds = p.ldg;
              uT.push_back(t);
            } else {
              t.soff = p.offQ + (int64_t)ii * p.ldq + r0 + cc;
              t.lds = p.ldq;
              uT_q.push_back(t);
            }
          }
      }
      // AR[mi,s2,r] [kp x n_r] = V^T[:kp, co:+n_r] / w;  V^T = (u_in_g ? Q2 : G2)
      for (auto& c : p.cols) {
        const Block& ab = AR->blocks[AR->find(mi, c.second, c.first)];
        const int c0 = p.co.at(c);
        const double w = std::sqrt((double)sdim(sym, x2->s1.sec[c.first]) / sdim(sym, cm));
        for (int ii = 0; ii < kp; ii += 32)
          for (int cc = 0; cc < ab.cols; cc += 32) {
            TrBlock t{};
            t.rows = std::min(32, kp - ii);
            t.cols = std::min(32, ab.cols - cc);
            t.scale = 1.0 / w;
            t.ldd = ab.ld;
            t.doff = ab.off + (int64_t)ii * ab.ld + cc;
            if (p.u_in_g) {
              t.soff = p.offQ + (int64_t)ii * p.ldq + c0 + cc;
              t.lds = p.ldq;
              vC.push_back(t);
            } else {
              t.soff = p.offG + (int64_t)ii * p.ldg + c0 + cc;
              t.lds = p.ldg;
              vC_g.push_back(t);
            }
          }
      }
    }
    auto run_tiles = [&](const std::vector<TrBlock>& v, bool transpose, const double* src, double* dst) -> bool {
      if (v.empty()) return true;
      TrBlock* d = nullptr;
      if (cudaMalloc(&d, v.size() * sizeof(TrBlock)) != cudaSuccess) return false;
      h2d_on_stream(d, v.data(), v.size() * sizeof(TrBlock), st);
      if (transpose)
        launch_transpose(d, (int)v.size(), src, dst, st);
      else
        launch_copy2d(d, (int)v.size(), src, dst, st);
      cudaStreamSynchronize(st);
      cudaFree(d);
      return true;
    };
    if (!run_tiles(uT, true, dG2, AL->d) || !run_tiles(uT_q, true, dQ2, AL->d) || !run_tiles(vC, false, dQ2, AR->d) ||
        !run_tiles(vC_g, false, dG2, AR->d))
      return fail(HTN_ERR_OOM, "tsvd: table allocation failed");
    if ((rc = htn_tensor_upload(Cb, cpacked.data(), (int64_t)cpacked.size()))) return fail(rc, ctx->err);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(HTN_ERR_CUDA, std::string("tsvd: ") + cudaGetErrorString(e));
    cleanup();
    *Vm_out = Vm;
    *AL_out = AL;
    *C_out = Cb;
    *AR_out = AR;
    if (discarded_weight) *discarded_weight = disc / std::max(nrm2, 1e-300);
    if (kept) *kept = nkept;
    return HTN_OK;
  } catch (const std::exception& e) {
    return fail(HTN_ERR_INVALID, std::string("tsvd: ") + e.what());
  }
}

// Schmidt values of a bond matrix C per sector (descending), by the same one-sided Jacobi kernels.
// Replaces `entanglement_spectrum(psi, site)` of MPSKit (SVD of C, the observable the north star names);
// every value of sector c appears dim(c) times in the full spectrum.  out: concatenation over the
// sectors in block order, counts[c] = n_c values each (arrays sized by htn_tensor_blocktable).
extern "C" int32_t htn_entanglement_spectrum(const htn_tensor* C, double* out, int64_t nout) {
  if (!C || !out) return HTN_ERR_INVALID;
  htn_ctx* ctx = C->ctx;
  if (C->kind != HTN_T_BOND) return ctx->fail(HTN_ERR_INVALID, "entanglement_spectrum: C must be a bond tensor");
  int64_t total = 0;
  for (const Block& b : C->blocks) total += b.rows;
  if (nout != total) return ctx->fail(HTN_ERR_SHAPE, "entanglement_spectrum: output must hold sum_c n_c values");
  if (total == 0) return HTN_OK;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  std::vector<SvdPanel> sp;
  int64_t totG = 0, totS = 0;
  for (const Block& b : C->blocks) {
    if (b.rows > 1024) return ctx->fail(HTN_ERR_SHAPE, "entanglement_spectrum: sector multiplicity above 1024 is not supported");
    if (b.rows == 0) continue;
    sp.push_back(SvdPanel{totG, totG, totS, b.rows, b.cols, b.ld, b.ld, 1, 0});
    totG = align_up(totG + (int64_t)b.rows * b.ld, 16);
    totS += b.rows;
  }
  double *dG = nullptr, *dQ = nullptr, *dG2 = nullptr, *dQ2 = nullptr, *dS = nullptr;
  SvdPanel* dP = nullptr;
  auto cleanup = [&]() {
    cudaFree(dG);
    cudaFree(dQ);
    cudaFree(dG2);
    cudaFree(dQ2);
    cudaFree(dS);
    cudaFree(dP);
  };
  if (cudaMalloc(&dG, totG * 8) != cudaSuccess || cudaMalloc(&dQ, totG * 8) != cudaSuccess ||
      cudaMalloc(&dG2, totG * 8) != cudaSuccess || cudaMalloc(&dQ2, totG * 8) != cudaSuccess ||
      cudaMalloc(&dS, totS * 8) != cudaSuccess || cudaMalloc(&dP, sp.size() * sizeof(SvdPanel)) != cudaSuccess) {
    cleanup();
    return ctx->fail(HTN_ERR_OOM, "entanglement_spectrum: workspace allocation failed");
  }
  // panels = the blocks of C themselves (row-major, ld as in the tensor): copy block by block
  {
    size_t pi = 0;
    for (const Block& b : C->blocks) {
      if (b.rows == 0) continue;
      cudaMemcpyAsync(dG + sp[pi].offG, C->d + b.off, (size_t)b.rows * b.ld * 8, cudaMemcpyDeviceToDevice, st);
      ++pi;
    }
  }
  h2d_on_stream(dP, sp.data(), sp.size() * sizeof(SvdPanel), st);
  const int rc = launch_svd(dP, sp.data(), (int)sp.size(), dG, dQ, dG2, dQ2, dS, st);
  int32_t ret = HTN_OK;
  if (rc < 0)
    ret = ctx->fail(HTN_ERR_CUDA, "entanglement_spectrum: Jacobi launch failed");
  else if (rc > 0)
    ret = ctx->fail(HTN_ERR_INVALID, "entanglement_spectrum: Jacobi sweeps did not converge");
  else {
    cudaMemcpyAsync(out, dS, totS * 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) ret = ctx->fail(HTN_ERR_CUDA, "entanglement_spectrum: readback failed");
  }
  cleanup();
  return ret;
}
