// Planner + executor of the one-site effective Hamiltonian  y = sum_{a,b} GL[a] x W[a,b] GR[b].
//
// Replaces MPSKit 0.13.1 `AC_hamiltonian` / `∂AC` (reached from find_groundstate,
// /root/reference/src/HubbardFunctions.jl:1012,1017,1027) together with the TensorKit /
// TensorOperations machinery under it (permute + recouple + one BLAS gemm per coupled
// sector).  Here the fusion-tree bookkeeping is done ONCE per plan on
// the host and lowered to three device work lists (same staging as oracle/heff.py
// HeffACPlan):
//   stage L  grouped GEMM : T[a,l',l,s,r]   = GL[a,l',l] . x[l,s,r]
//   stage W  mix          : U[b,l',s',r',r] = sum coef . T      (coef = w N / dim r')
//                           y[l',s',r']     = sum coef . T|x    (identity right level)
//   stage R  grouped GEMM : y[l',s',r']    += sum_{b,r} U[b,l',s',r',r] . GR[b,r,r']
// Identity environment levels (GL[1] = 1, GR[chi] = 1) are elided exactly.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <tuple>

#include "htn_internal.hpp"

using namespace htn;

extern "C" {
// internal helpers defined in htn_api.cpp (not part of include/htn.h)
int32_t htn_upload_locked(htn_tensor* t, const double* host, int64_t nelem);
int32_t htn_download_locked(const htn_tensor* t, double* host, int64_t nelem);
bool htn_same_structure(const htn_tensor* x, const htn_tensor* y);
}

namespace {

struct TileSpec {
  int mo, mn, no, nn, layout;
};

// near-equal split of an extent into pieces of at most 64, each a multiple of the DMMA atom (8)
std::vector<std::pair<int, int>> split_flex(int n) {
  std::vector<std::pair<int, int>> out;
  int atoms = (n + 7) / 8, nt = (atoms + 7) / 8;
  int base = atoms / nt, rem = atoms % nt, o = 0;
  for (int i = 0; i < nt; ++i) {
    int len = std::min(8 * (base + (i < rem ? 1 : 0)), n - o);
    out.push_back({o, len});
    o += len;
  }
  return out;
}

// pipe cost (executed DMMA atoms, 2 per active strip and flex atom) + a small charge for idle strips
double tile_cost(int flex_ext, int fixed_ext) {
  int flex = (flex_ext + 7) / 8, strips = (fixed_ext + 15) / 16;
  return flex * 2.0 * strips + 0.25 * flex * 2.0 * (4 - strips);
}

// Cover an M x N block with CTA tiles (see the layout comment in htn_kernels.cu).  Variant 0:
// full 64-column tiles in layout A (rows split flexibly) + the remaining columns as layout-B
// tiles (64-row strips, flex = remaining columns).  Variant 1: the transpose.  Cheapest wins.
std::vector<TileSpec> tile_block(int M, int N) {
  std::vector<TileSpec> best;
  double best_cost = 1e300;
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<TileSpec> v;
    double cost = 0;
    if (variant == 0) {
      int nfull = N / 64, rn = N % 64;
      for (int j = 0; j < nfull; ++j)
        for (auto& pm : split_flex(M)) {
          v.push_back({pm.first, pm.second, j * 64, 64, 0});
          cost += tile_cost(pm.second, 64);
        }
      if (rn)
        for (int mo = 0; mo < M; mo += 64) {
          int mn = std::min(64, M - mo);
          v.push_back({mo, mn, nfull * 64, rn, 1});
          cost += tile_cost(rn, mn);
        }
    } else {
      int mfull = M / 64, rm = M % 64;
      for (int i = 0; i < mfull; ++i)
        for (auto& pn : split_flex(N)) {
          v.push_back({i * 64, 64, pn.first, pn.second, 1});
          cost += tile_cost(pn.second, 64);
        }
      if (rm)
        for (int no = 0; no < N; no += 64) {
          int nn = std::min(64, N - no);
          v.push_back({mfull * 64, rm, no, nn, 0});
          cost += tile_cost(rm, nn);
        }
    }
    if (cost < best_cost) {
      best_cost = cost;
      best = v;
    }
  }
  return best;
}

struct WsBlock {  // workspace block (T or U)
  int rows, cols, ld;
  int64_t off;
};

template <class T>
int32_t to_device(htn_ctx* ctx, const std::vector<T>& v, T** out) {
  *out = nullptr;
  if (v.empty()) return HTN_OK;
  cudaError_t e = cudaMalloc(out, v.size() * sizeof(T));
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_OOM, std::string("cudaMalloc(plan table): ") + cudaGetErrorString(e));
  e = cudaMemcpy(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_CUDA, std::string("cudaMemcpy(plan table): ") + cudaGetErrorString(e));
  return HTN_OK;
}

}  // namespace

extern "C" {

int32_t htn_plan_destroy(htn_plan* p) {
  if (!p) return HTN_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  cudaFree(p->T);
  cudaFree(p->gsrcs);
  cudaFree(p->U);
  cudaFree(p->Pp);
  cudaFree(p->itemsL);
  cudaFree(p->segsL);
  cudaFree(p->itemsR);
  cudaFree(p->segsR);
  cudaFree(p->mixT);
  cudaFree(p->mixS);
  cudaFree(p->mixC);
  if (p->like) htn_tensor_destroy(p->like);
  if (p->hx) htn_tensor_destroy(p->hx);
  if (p->hy) htn_tensor_destroy(p->hy);
  delete p;
  return HTN_OK;
}

int32_t htn_plan_heff_ac(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W, const htn_tensor* GR,
                         const htn_tensor* like, htn_plan** out) {
  if (!ctx || !GL || !W || !GR || !like || !out) return HTN_ERR_INVALID;
  *out = nullptr;
  if (GL->kind != HTN_T_ENVL || GR->kind != HTN_T_ENVR || like->kind != HTN_T_MPS)
    return ctx->fail(HTN_ERR_INVALID, "plan_heff_ac: wrong tensor kinds");
  const int sym = like->sym;
  if (GL->sym != sym || GR->sym != sym || W->sym != sym) return ctx->fail(HTN_ERR_INVALID, "plan_heff_ac: symmetry kinds differ");
  if (GL->s0.sec != like->s0.sec || GL->s0.mult != like->s0.mult)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_ac: GL bond space != left space of x");
  if (GR->s0.sec != like->s1.sec || GR->s0.mult != like->s1.mult)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_ac: GR bond space != right space of x");
  if (GL->legs.sec != W->Ml.sec || GR->legs.sec != W->Mr.sec || like->legs.sec != W->P.sec)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_ac: MPO legs do not match the environments / physical space");

  htn_tensor* like_copy = nullptr;
  int32_t rc = htn_tensor_create_like(like, &like_copy);
  if (rc) return rc;

  std::lock_guard<std::mutex> g(ctx->mu);
  htn_plan* p = nullptr;
  try {
    p = new htn_plan();
    p->ctx = ctx;
    p->GL = GL;
    p->GR = GR;
    p->like = like_copy;
    const auto& Vl = like->s0;
    const auto& Vr = like->s1;
    const auto& P = like->legs;

    // partner tables: (a,l) -> l' ; (b,r) -> r'
    std::map<std::pair<int, int>, std::vector<int>> pl, pr;
    for (const Block& b : GL->blocks) pl[{b.lab[0], b.lab[2]}].push_back(b.lab[1]);
    for (const Block& b : GR->blocks) pr[{b.lab[0], b.lab[1]}].push_back(b.lab[2]);
    std::vector<std::vector<std::pair<int, int>>> xs(P.sec.size());
    for (const Block& b : like->blocks) xs[b.lab[1]].push_back({b.lab[0], b.lab[2]});

    // term accumulation keyed by (y block, a, l, s, r, b)
    typedef std::tuple<int, int, int, int, int, int, int, int> TermKey;  // lp,sp,rp,a,l,s,r,b
    std::map<TermKey, double> terms;
    for (const MpoEntry& e : W->entries) {
      for (auto& lr : xs[e.s]) {
        int l = lr.first, r = lr.second;
        auto itl = pl.find({e.a, l});
        auto itr = pr.find({e.b, r});
        if (itl == pl.end() || itr == pr.end()) continue;
        for (int lp : itl->second)
          for (int rp : itr->second) {
            if (like->find(lp, e.sp, rp) < 0) continue;
            double n = network(sym, Vl.sec[lp], P.sec[e.sp], Vr.sec[rp], Vl.sec[l], P.sec[e.s], Vr.sec[r],
                               W->Ml.sec[e.a], W->Mr.sec[e.b], e.c);
            if (n == 0.0) continue;
            terms[TermKey(lp, e.sp, rp, e.a, l, e.s, r, e.b)] += e.w * n / sdim(sym, Vr.sec[rp]);
          }
      }
    }

    const int idL = GL->identity_level, idR = GR->identity_level;
    // workspace block lists
    std::map<std::tuple<int, int, int, int, int>, int> tindex, uindex;  // (a,lp,l,s,r) / (b,lp,sp,rp,r)
    std::vector<std::tuple<int, int, int, int, int>> tkeys, ukeys;
    std::vector<WsBlock> tb, ub;
    int64_t toff = 0, uoff = 0;
    struct Src {
      int base;
      int64_t off;
      double coef;
    };
    // mix targets: U blocks and every y block
    std::vector<std::vector<Src>> usrc;
    std::vector<std::vector<Src>> ysrc(like->blocks.size());
    for (auto& kv : terms) {
      if (kv.second == 0.0) continue;
      int lp, sp, rp, a, l, s, r, b;
      std::tie(lp, sp, rp, a, l, s, r, b) = kv.first;
      Src src;
      src.coef = kv.second;
      const int nlp = Vl.mult[lp], nr = Vr.mult[r];
      if (a == idL) {
        if (lp != l) continue;  // cannot happen: identity level is trivial
        src.base = B_X;
        src.off = like->blocks[like->find(l, s, r)].off;
      } else {
        auto key = std::make_tuple(a, lp, l, s, r);
        auto it = tindex.find(key);
        int ti;
        if (it == tindex.end()) {
          ti = (int)tb.size();
          tindex[key] = ti;
          tkeys.push_back(key);
          WsBlock w{nlp, nr, even_up(nr), toff};
          toff = align_up(toff + (int64_t)w.rows * w.ld, 16);
          tb.push_back(w);
        } else
          ti = it->second;
        src.base = B_T;
        src.off = tb[ti].off;
      }
      if (b == idR) {
        ysrc[like->find(lp, sp, rp)].push_back(src);
      } else {
        auto key = std::make_tuple(b, lp, sp, rp, r);
        auto it = uindex.find(key);
        int ui;
        if (it == uindex.end()) {
          ui = (int)ub.size();
          uindex[key] = ui;
          ukeys.push_back(key);
          WsBlock w{nlp, nr, even_up(nr), uoff};
          uoff = align_up(uoff + (int64_t)w.rows * w.ld, 16);
          ub.push_back(w);
          usrc.emplace_back();
        } else
          ui = it->second;
        usrc[ui].push_back(src);
      }
    }
    p->t_elems = std::max<int64_t>(toff, 16);
    p->u_elems = std::max<int64_t>(uoff, 16);

    // ---- stage L work list ---------------------------------------------------------------
    std::vector<GemmItem> itemsL;
    std::vector<GemmSeg> segsL;
    double flopsL = 0, flopsR = 0, padded = 0;
    for (size_t ti = 0; ti < tb.size(); ++ti) {
      int a, lp, l, s, r;
      std::tie(a, lp, l, s, r) = tkeys[ti];
      const Block& gl = GL->blocks[GL->find(a, lp, l)];
      const Block& xb = like->blocks[like->find(l, s, r)];
      const WsBlock& w = tb[ti];
      flopsL += 2.0 * w.rows * w.cols * gl.cols;
      for (const TileSpec& ts : tile_block(w.rows, w.cols)) {
        GemmSeg sg{};
        sg.a_base = B_GL;
        sg.a_off = gl.off + (int64_t)ts.mo * gl.ld;
        sg.lda = gl.ld;
        sg.b_base = B_X;
        sg.b_off = xb.off + ts.no;
        sg.ldb = xb.ld;
        sg.K = gl.cols;
        sg.nsrc = 0;
        sg.src_begin = 0;
        GemmItem it{};
        it.c_base = B_T;
        it.c_off = w.off + (int64_t)ts.mo * w.ld + ts.no;
        it.ldc = w.ld;
        it.mt = ts.mn;
        it.nt = ts.nn;
        it.layout = ts.layout;
        it.seg_begin = (int)segsL.size();
        it.seg_end = it.seg_begin + 1;
        it.nchunks = (sg.K + GEMM_BK - 1) / GEMM_BK;
        it.beta = 0;
        segsL.push_back(sg);
        itemsL.push_back(it);
        padded += 2.0 * ((ts.mn + 7) / 8 * 8) * ((ts.nn + 7) / 8 * 8) * ((sg.K + 3) / 4 * 4.0);
      }
    }

    // ---- stage R work list: one item per y tile, K-segments over all (b,r) ---------------
    std::vector<std::vector<int>> y_u(like->blocks.size());
    for (size_t ui = 0; ui < ub.size(); ++ui) {
      int b, lp, sp, rp, r;
      std::tie(b, lp, sp, rp, r) = ukeys[ui];
      y_u[like->find(lp, sp, rp)].push_back((int)ui);
    }
    std::vector<GemmItem> itemsR;
    std::vector<GemmSeg> segsR;
    // HTN_FUSE_W=1: assemble U inside the stage-R operand load instead of materialising it
    // (measured slower on B200 at D=1024: the producers become latency-bound; kept selectable)
    const char* fw = getenv("HTN_FUSE_W");
    const bool fuse_w = fw && fw[0] == '1';
    std::vector<MixSrc> gsrcs;  // mix sources of the U blocks when fused into stage R
    std::vector<int> usrc_begin(ub.size(), 0);
    for (size_t ui = 0; ui < ub.size(); ++ui) {
      usrc_begin[ui] = (int)gsrcs.size();
      for (const Src& sc : usrc[ui]) gsrcs.push_back(MixSrc{sc.off, sc.base, 0, sc.coef});
    }
    // split-K: a y block has few tiles but a K loop over every (level, sector) pair; cut the
    // segment list into nsplit parts of ~SPLIT_CHUNKS chunks, each writing its own partial copy
    // of the y tile (summed in fixed order by the final mix => deterministic)
    const int SPLIT_CHUNKS = 40, SPLIT_MAX = 32;
    std::vector<int> ysplits(like->blocks.size(), 0);
    int nsplit_max = 0;
    for (size_t yi = 0; yi < like->blocks.size(); ++yi) {
      if (y_u[yi].empty()) continue;
      const Block& yb = like->blocks[yi];
      std::sort(y_u[yi].begin(), y_u[yi].end(), [&](int i, int j) { return ukeys[i] < ukeys[j]; });
      int total_chunks = 0;
      for (int ui : y_u[yi]) total_chunks += (ub[ui].cols + GEMM_BK - 1) / GEMM_BK;
      int nsplit = std::max(1, std::min({SPLIT_MAX, (total_chunks + SPLIT_CHUNKS / 2) / SPLIT_CHUNKS,
                                         (int)y_u[yi].size()}));
      ysplits[yi] = nsplit;
      nsplit_max = std::max(nsplit_max, nsplit);
      // cut points over the segment list (balanced by chunk count)
      std::vector<int> cut(nsplit + 1, 0);
      {
        int acc = 0, sidx = 1;
        for (size_t q = 0; q < y_u[yi].size(); ++q) {
          acc += (ub[y_u[yi][q]].cols + GEMM_BK - 1) / GEMM_BK;
          while (sidx < nsplit && acc >= (long long)total_chunks * sidx / nsplit) cut[sidx++] = (int)q + 1;
        }
        for (; sidx <= nsplit; ++sidx) cut[sidx] = (int)y_u[yi].size();
        for (int q = 1; q <= nsplit; ++q) cut[q] = std::max(cut[q], cut[q - 1]);
      }
      for (const TileSpec& ts : tile_block(yb.rows, yb.cols))
        for (int sp_i = 0; sp_i < nsplit; ++sp_i) {
          if (cut[sp_i + 1] == cut[sp_i]) continue;
          GemmItem it{};
          it.c_base = B_P;
          it.c_off = (int64_t)sp_i * like->dsize + yb.off + (int64_t)ts.mo * yb.ld + ts.no;
          it.ldc = yb.ld;
          it.mt = ts.mn;
          it.nt = ts.nn;
          it.layout = ts.layout;
          it.seg_begin = (int)segsR.size();
          it.beta = 0;
          it.nchunks = 0;
          for (int q = cut[sp_i]; q < cut[sp_i + 1]; ++q) {
            int ui = y_u[yi][q];
            int b, lp, sp, rp, r;
            std::tie(b, lp, sp, rp, r) = ukeys[ui];
            const Block& gr = GR->blocks[GR->find(b, r, rp)];
            const WsBlock& w = ub[ui];
            GemmSeg sg{};
            sg.lda = w.ld;
            if (fuse_w) {
              // A operand = sum_j coef_j * source_j, rows ts.mo.. of every source block
              sg.a_base = B_X;  // unused when nsrc > 0
              sg.a_off = (int64_t)ts.mo * w.ld;
              sg.nsrc = (int)usrc[ui].size();
              sg.src_begin = usrc_begin[ui];
            } else {
              sg.a_base = B_U;
              sg.a_off = w.off + (int64_t)ts.mo * w.ld;
              sg.nsrc = 0;
              sg.src_begin = 0;
            }
            sg.b_base = B_GR;
            sg.b_off = gr.off + ts.no;
            sg.ldb = gr.ld;
            sg.K = gr.rows;
            segsR.push_back(sg);
            it.nchunks += (sg.K + GEMM_BK - 1) / GEMM_BK;
            padded += 2.0 * ((ts.mn + 7) / 8 * 8) * ((ts.nn + 7) / 8 * 8) * ((sg.K + 3) / 4 * 4.0);
          }
          it.seg_end = (int)segsR.size();
          itemsR.push_back(it);
        }
      // partial copies become extra sources of the y block in the final mix; a split whose
      // segment range is empty never writes its copy, so only non-empty splits are listed
      for (int sp_i = 0; sp_i < nsplit; ++sp_i)
        if (cut[sp_i + 1] > cut[sp_i])
          ysrc[yi].push_back(Src{B_P, (int64_t)sp_i * like->dsize + yb.off, 1.0});
    }
    p->p_elems = std::max<int64_t>((int64_t)nsplit_max * like->dsize, 16);
    for (size_t ui = 0; ui < ub.size(); ++ui) {
      int b, lp, sp, rp, r;
      std::tie(b, lp, sp, rp, r) = ukeys[ui];
      flopsR += 2.0 * ub[ui].rows * ub[ui].cols * Vr.mult[rp];
    }
    // heaviest tiles first (persistent CTAs take items round-robin)
    auto cost = [](const GemmItem& it) { return (double)it.mt * it.nt * it.nchunks; };
    std::stable_sort(itemsL.begin(), itemsL.end(), [&](const GemmItem& a, const GemmItem& b) { return cost(a) > cost(b); });
    std::stable_sort(itemsR.begin(), itemsR.end(), [&](const GemmItem& a, const GemmItem& b) { return cost(a) > cost(b); });

    // ---- stage W work list ------------------------------------------------------------------
    std::vector<MixTarget> mixT;
    std::vector<MixSrc> mixS;
    std::vector<MixChunk> mixC;
    auto add_target = [&](int base, int64_t off, int nelem, const std::vector<Src>& srcs) {
      MixTarget t{};
      t.base = base;
      t.off = off;
      t.nelem = nelem;
      t.src_begin = (int)mixS.size();
      for (const Src& s : srcs) mixS.push_back(MixSrc{s.off, s.base, 0, s.coef});
      t.src_end = (int)mixS.size();
      int ti = (int)mixT.size();
      mixT.push_back(t);
      // ~16k element-sources per CTA, chunk a multiple of 512 elements (256 threads x double2)
      int per = 16384 / std::max<int>(1, (int)srcs.size());
      per = std::max(512, std::min(8192, per / 512 * 512));
      for (int e = 0; e < nelem; e += per) mixC.push_back(MixChunk{ti, e, std::min(per, nelem - e), 0});
    };
    if (!fuse_w)
      for (size_t ui = 0; ui < ub.size(); ++ui) add_target(B_U, ub[ui].off, ub[ui].rows * ub[ui].ld, usrc[ui]);
    const int nmixCU = (int)mixC.size();
    for (size_t yi = 0; yi < like->blocks.size(); ++yi) {
      const Block& yb = like->blocks[yi];
      add_target(B_Y, yb.off, yb.rows * yb.ld, ysrc[yi]);
    }

    // ---- upload ------------------------------------------------------------------------------
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&p->T, p->t_elems * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&p->Pp, p->p_elems * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&p->U, (fuse_w ? 16 : p->u_elems) * sizeof(double)) != cudaSuccess) {
      htn_plan_destroy(p);
      return ctx->fail(HTN_ERR_OOM, "plan_heff_ac: workspace allocation failed");
    }
    {  // resolve plan-lifetime arrays to absolute pointers, keep x / y relative
      auto fix = [&](long long& off, int& base) {
        const double* b = nullptr;
        switch (base) {
          case B_GL: b = GL->d; break;
          case B_GR: b = GR->d; break;
          case B_T: b = p->T; break;
          case B_P: b = p->Pp; break;
          case B_U: b = p->U; break;
          case B_X: base = REF_X; return;
          case B_Y: base = REF_Y; return;
        }
        off = reinterpret_cast<long long>(b + off);
        base = REF_ABS;
      };
      for (auto* segs : {&segsL, &segsR})
        for (GemmSeg& sg : *segs) {
          if (sg.nsrc == 0) fix(sg.a_off, sg.a_base);
          fix(sg.b_off, sg.b_base);
        }
      for (MixSrc& sc : gsrcs) fix(sc.off, sc.base);
      for (auto* items : {&itemsL, &itemsR})
        for (GemmItem& it : *items) fix(it.c_off, it.c_base);
      for (MixTarget& t : mixT) fix(t.off, t.base);
      for (MixSrc& sc : mixS) fix(sc.off, sc.base);
    }
    cudaMemset(p->T, 0, p->t_elems * sizeof(double));
    cudaMemset(p->Pp, 0, p->p_elems * sizeof(double));
    if ((rc = to_device(ctx, itemsL, &p->itemsL)) || (rc = to_device(ctx, segsL, &p->segsL)) ||
        (rc = to_device(ctx, itemsR, &p->itemsR)) || (rc = to_device(ctx, segsR, &p->segsR)) ||
        (rc = to_device(ctx, gsrcs, &p->gsrcs)) || (rc = to_device(ctx, mixT, &p->mixT)) || (rc = to_device(ctx, mixS, &p->mixS)) ||
        (rc = to_device(ctx, mixC, &p->mixC))) {
      htn_plan_destroy(p);
      return rc;
    }
    p->nitemsL = (int)itemsL.size();
    p->nsegsL = (int)segsL.size();
    p->nitemsR = (int)itemsR.size();
    p->nsegsR = (int)segsR.size();
    p->nmixT = (int)mixT.size();
    p->nmixS = (int)mixS.size();
    p->nmixC = (int)mixC.size();
    p->nmixCU = nmixCU;
    const int cap = ctx->sm_count * gemm_max_ctas_per_sm();
    p->gridL = std::max(1, std::min(p->nitemsL, cap));
    p->gridR = std::max(1, std::min(p->nitemsR, cap));
    p->stats[0] = flopsL + flopsR;
    p->stats[1] = flopsL;
    p->stats[2] = flopsR;
    p->stats[3] = (double)tb.size();
    p->stats[4] = (double)ub.size();
    p->stats[5] = (double)mixT.size();
    p->stats[6] = (double)(mixS.size() + gsrcs.size());
    p->stats[7] = (double)(p->t_elems + p->p_elems + (fuse_w ? 0 : p->u_elems)) * sizeof(double);
    p->stats[8] = (double)itemsL.size();
    p->stats[9] = (double)itemsR.size();
    p->stats[10] = padded;
    p->stats[11] = (p->nitemsL > 0) + (p->nmixCU > 0) + (p->nitemsR > 0) + (p->nmixC > p->nmixCU);
  } catch (const std::bad_alloc&) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_OOM, "plan_heff_ac: host allocation failed");
  } catch (const std::exception& e) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_INVALID, std::string("plan_heff_ac: ") + e.what());
  }
  *out = p;
  return HTN_OK;
}

static int32_t check_xy(htn_plan* p, const htn_tensor* x, const htn_tensor* y) {
  if (!htn_same_structure(p->like, x) || !htn_same_structure(p->like, y))
    return p->ctx->fail(HTN_ERR_SHAPE, "heff_apply: x / y do not have the plan's block structure");
  if (x == y || x->d == y->d) return p->ctx->fail(HTN_ERR_INVALID, "heff_apply: x and y must be distinct tensors");
  return HTN_OK;
}

// stage mask: 1 = L (T = GL.x), 2 = W (U = mix T), 4 = R (partials = U.GR), 8 = final mix (y)
static int32_t run_stages(htn_plan* p, const htn_tensor* x, htn_tensor* y, int mask) {
  htn_ctx* ctx = p->ctx;
  Bases bs;
  bs.x = x->d;
  bs.y = y->d;
  if (mask & 1) launch_gemm(p->itemsL, p->segsL, p->gsrcs, p->nitemsL, bs, p->gridL, ctx->stream);
  if (mask & 2) launch_mix(p->mixT, p->mixS, p->mixC, p->nmixCU, bs, ctx->stream);
  if (mask & 4) launch_gemm(p->itemsR, p->segsR, p->gsrcs, p->nitemsR, bs, p->gridR, ctx->stream);
  if (mask & 8) launch_mix(p->mixT, p->mixS, p->mixC + p->nmixCU, p->nmixC - p->nmixCU, bs, ctx->stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_CUDA, std::string("heff_apply launch: ") + cudaGetErrorString(e));
  return HTN_OK;
}

int32_t htn_heff_apply(htn_plan* p, const htn_tensor* x, htn_tensor* y) {
  if (!p || !x || !y) return HTN_ERR_INVALID;
  std::lock_guard<std::mutex> g(p->ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(p->ctx->device);
  return run_stages(p, x, y, 15);
}

int32_t htn_heff_apply_host(htn_plan* p, const double* x_host, double* y_host, int64_t nelem) {
  if (!p || !x_host || !y_host) return HTN_ERR_INVALID;
  htn_ctx* ctx = p->ctx;
  int32_t rc;
  if (!p->hx) {
    if ((rc = htn_tensor_create_like(p->like, &p->hx))) return rc;
    if ((rc = htn_tensor_create_like(p->like, &p->hy))) return rc;
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if ((rc = htn_upload_locked(p->hx, x_host, nelem))) return rc;
  if ((rc = run_stages(p, p->hx, p->hy, 15))) return rc;
  return htn_download_locked(p->hy, y_host, nelem);
}

int32_t htn_heff_time(htn_plan* p, const htn_tensor* x, htn_tensor* y, int32_t reps, float* ms_total) {
  if (!p || !x || !y || !ms_total || reps <= 0) return HTN_ERR_INVALID;
  htn_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStreamSynchronize(ctx->stream);
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < reps && rc == HTN_OK; ++r) rc = run_stages(p, x, y, 15);
  cudaEventRecord(e1, ctx->stream);
  cudaError_t e = cudaEventSynchronize(e1);
  if (rc == HTN_OK && e != cudaSuccess) rc = ctx->fail(HTN_ERR_CUDA, std::string("heff_time: ") + cudaGetErrorString(e));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  *ms_total = t;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int32_t htn_plan_stats(const htn_plan* p, double* stats, int32_t n) {
  if (!p || !stats) return HTN_ERR_INVALID;
  for (int i = 0; i < n && i < 12; ++i) stats[i] = p->stats[i];
  return HTN_OK;
}

int32_t htn_plan_profile(htn_plan* p, const htn_tensor* x, htn_tensor* y, int32_t reps, float* ms) {
  if (!p || !x || !y || !ms || reps <= 0) return HTN_ERR_INVALID;
  htn_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> g(ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int masks[4] = {15, 1, 2 | 8, 4};
  for (int k = 0; k < 4; ++k) {
    cudaStreamSynchronize(ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    for (int r = 0; r < reps; ++r)
      if ((rc = run_stages(p, x, y, masks[k]))) break;
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    if (rc == HTN_OK && e != cudaSuccess) rc = ctx->fail(HTN_ERR_CUDA, std::string("profile: ") + cudaGetErrorString(e));
    if (rc) break;
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    ms[k] = t / reps;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc == HTN_OK) rc = run_stages(p, x, y, 15);  // leave y = H x behind
  return rc;
}

}  // extern "C"
