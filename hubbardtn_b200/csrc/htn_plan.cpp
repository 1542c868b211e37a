// Operator planners: lower the effective Hamiltonians and MPO transfers to contraction programs.
//
// Replaces MPSKit 0.13.1 `AC_hamiltonian`/`∂AC`, `C_hamiltonian`/`∂C` and `TransferMatrix`
// (reached from find_groundstate, /root/reference/src/HubbardFunctions.jl:1012,1017,1027)
// together with the TensorKit / TensorOperations machinery under them (permute + recouple +
// one BLAS gemm per coupled sector).  The fusion-tree bookkeeping is done
// ONCE per plan on the host; the staging is the one of oracle/heff.py:
//   H_AC   stage L : T[a,l',l,s,r]   = GL[a,l',l] . x[l,s,r]
//          stage W : U[b,l',s',r',r] = sum coef . T            (coef = w N / dim r')
//          stage R : y[l',s',r']     = sum_{b,r} U . GR[b,r,r']  (+ identity right level)
//   H_C    stage L : T[a,c',c] = GL[a,c',c] . C[c] ;  stage R : y[c'] = sum_{a,c} T . GR[a,c,c']
//   T_L    stages L, W as H_AC with x = A, then  GL'[b,r',r] = sum_{l',s'} A^T[l',s',r'] . U
//   T_R    T[b,l,s,r,r'] = A[l,s,r] . GR[b,r,r'] ; U[a,l,l',s',r'] = sum coef . T ;
//          GR'[a,l,l'] = sum_{s',r'} U . A^T[l',s',r']
// Identity environment levels (GL[first] = 1, GR[last] = 1) are elided exactly.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <new>
#include <tuple>

#include "htn_program.hpp"

using namespace htn;

extern "C" {
int32_t htn_upload_locked(htn_tensor* t, const double* host, int64_t nelem);
int32_t htn_download_locked(const htn_tensor* t, double* host, int64_t nelem);
bool htn_same_structure(const htn_tensor* x, const htn_tensor* y);
}

namespace htn {

// environment view: (level, i, j) -> block; a BOND tensor acts as a one-level environment
struct EnvView {
  const htn_tensor* t;
  int find(int a, int i, int j) const {
    if (t->kind == HTN_T_BOND) return (a == 0 && i == j) ? t->find(i, i, i) : -1;
    return t->find(a, i, j);
  }
  int identity_level() const { return t->kind == HTN_T_BOND ? -1 : t->identity_level; }
};

struct WsBlock {
  int rows, cols, ld;
  int64_t off;
};

struct Src {
  Opnd o;
  double coef;
};

// One term of an effective-Hamiltonian-like contraction:
//   Y[yi] += coef * GL[a,lp,l] . X[xi] . (right object of level b mapping r -> rp)
// X / Y blocks are referred to by their index in `like` (so the same back end serves one-site
// tensors (l,s,r) and two-site tensors (l,s1,m,s2,r)).
struct HTerm {
  int yi, xi, a, lp, l, b, r, rp;
  bool operator<(const HTerm& o) const {
    return std::tie(yi, a, l, xi, b, r) < std::tie(o.yi, o.a, o.l, o.xi, o.b, o.r);
  }
};

// Front end shared by H_AC, H_AC2 and T_L: stage L (T = GL . X, once per (a,lp,l,xi)) and the mix
// sources of every U block (b, yi, r).
//
// Stage L comes in two forms.  Classic: one grouped-GEMM task per T block.  Stacked (H_AC / H_AC2 with a real
// environment tensor, contracted multiplicity small enough for the shared-memory slab): for every x block all T
// blocks {T[a,lp,l,xi]}_(lp,a) are rows of ONE tall array that the stacked kernel (htn_stackl.cuh) fills from the GL
// panel of sector l; the left sectors lp are grouped into waves and the mix targets of a wave are formed by the
// mixer warps of the same launch as soon as the wave's jobs are done.
struct LeftFront {
  std::vector<std::tuple<int, int, int>> ukeys;  // (b, yi, r)
  std::vector<WsBlock> ub;
  std::vector<std::vector<Src>> usrc;
  int n_t = 0;
  bool stacked = false;
  std::vector<StackJobH> jobs;
  std::vector<int> wave_of_lp;
  int nwaves = 0;
};

static int stack_max_atoms(int K) {
  const int k4 = (K + 3) & ~3;
  int at = (STACK_SLAB_ELEMS / std::max(k4, 4) - 4) / 8;
  return std::min(at, 7);
}

static int tile_waves_mode = 0;

static LeftFront build_front(Program& pg, const htn_tensor* like, const EnvView& GL,
                             const std::map<HTerm, double>& terms, int slot_x, int slot_gl, bool allow_stack = false) {
  LeftFront F;
  const int idL = GL.identity_level();
  typedef std::tuple<int, int, int, int> TKey;  // (a,lp,l,xi)
  std::map<TKey, int> tindex;
  std::vector<TKey> tkeys;
  for (auto& kv : terms) {
    if (kv.second == 0.0) continue;
    const HTerm& t = kv.first;
    if (t.a == idL) continue;
    TKey key(t.a, t.lp, t.l, t.xi);
    if (!tindex.count(key)) {
      tindex[key] = (int)tkeys.size();
      tkeys.push_back(key);
    }
  }
  static int stack_env = -1;
  if (stack_env < 0) {
    const char* e = getenv("HTN_STACK");  // 0: classic stage L everywhere (A/B experiments)
    stack_env = e ? atoi(e) : 1;
  }
  const htn_tensor* G = GL.t;
  const bool can_stack = allow_stack && stack_env && G->kind == HTN_T_ENVL && !G->panels.empty();
  std::vector<WsBlock> tb(tkeys.size());
  std::vector<char> is_stacked(tkeys.size(), 0);
  std::vector<GemmTaskH> tasksL;
  // ---- which T blocks go through the stacked kernel: contracted multiplicity must leave >= 4 column atoms ----
  std::map<int, std::vector<int>> by_xi;  // xi -> T indices (stacked ones)
  for (size_t ti = 0; ti < tkeys.size(); ++ti) {
    int a, lp, l, xi;
    std::tie(a, lp, l, xi) = tkeys[ti];
    const Block& xb = like->blocks[xi];
    const Block& yb_rows = G->blocks[GL.find(a, lp, l)];
    tb[ti] = WsBlock{yb_rows.rows, xb.cols, even_up(xb.cols), 0};
    if (can_stack && stack_max_atoms(yb_rows.cols) >= std::min(4, (xb.cols + 7) / 8) && yb_rows.cols >= 1) {
      is_stacked[ti] = 1;
      by_xi[xi].push_back((int)ti);
    }
  }
  // ---- the stacked kernel pays off only when there are enough 64-row tiles to keep its 296 persistent CTAs busy with
  // multi-tile jobs (C4-synthetic, chi = 96: 16.5 k tiles).  On a small MPO -- the reference's polyacetylene model has
  // chi = 10: ~2 k tiles, one job per CTA -- the classic grouped GEMM with its static balanced schedule is 3.7x faster for
  // stage L (0.036 ms against 0.133 ms at D_red = 1024; gpurun_out/r2_real_c4_variants.txt), so those plans stay classic.
  {
    static long long min_tiles = -1;
    if (min_tiles < 0) {
      const char* e = getenv("HTN_STACK_MIN_TILES");
      min_tiles = e ? atoll(e) : 7000;
    }
    long long tiles = 0;
    for (auto& kv : by_xi) {
      const Block& xb = like->blocks[kv.first];
      long long rows = 0;
      for (int ti : kv.second) rows += tb[ti].rows;
      const int l = std::get<2>(tkeys[kv.second[0]]);
      const int atoms = (xb.cols + 7) / 8, maxat = std::max(1, stack_max_atoms(G->panels[l].cols));
      tiles += (rows + 63) / 64 * ((atoms + maxat - 1) / maxat);
    }
    if (tiles < min_tiles) {
      by_xi.clear();
      std::fill(is_stacked.begin(), is_stacked.end(), 0);
    }
  }
  // ---- waves.  Wave 0: the jobs of the LIGHT panels (contracted multiplicity <= 16: hardly any flops or bytes, but
  // many rows) -- one run per x block, every mix target waits for it.  Waves 1..: the left sectors lp grouped by the
  // bytes of heavy T they own (heavy sectors first); a target of sector lp waits for wave_of_lp[lp] (and wave 0).
  const int nlp = (int)like->s0.sec.size();
  F.wave_of_lp.assign(nlp, 0);
  auto light_panel = [&](int l) { return G->panels[l].cols <= 16; };
  if (!by_xi.empty()) {
    std::vector<double> bytes(nlp, 0.0);
    for (size_t ti = 0; ti < tkeys.size(); ++ti)
      if (is_stacked[ti] && !light_panel(std::get<2>(tkeys[ti]))) bytes[std::get<1>(tkeys[ti])] += 8.0 * tb[ti].rows * tb[ti].ld;
    static double wave_mb = -1.0, wave_taper = 1.0;
    if (wave_mb < 0) {
      const char* e = getenv("HTN_WAVE_MB");
      wave_mb = e ? atof(e) : 48.0;
      e = getenv("HTN_WAVE_TAPER");  // < 1: every wave holds this fraction of the bytes of the one before (short tail)
      if (e) wave_taper = atof(e);
    }
    std::vector<int> order;
    for (int lp = 0; lp < nlp; ++lp)
      if (bytes[lp] > 0) order.push_back(lp);
    static int wave_order = -1;
    if (wave_order < 0) {
      const char* e = getenv("HTN_WAVE_ORDER");  // 1: waves = consecutive left sectors (contiguous panel rows), 0: heavy sectors first
      wave_order = e ? atoi(e) : 0;
      e = getenv("HTN_TILE_WAVES");  // 1: jobs span waves, completion is reported tile by tile (implies HTN_WAVE_ORDER=1)
      tile_waves_mode = e ? atoi(e) : 0;
      if (tile_waves_mode) wave_order = 1;
    }
    if (!wave_order) std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return bytes[x] > bytes[y]; });
    // the LAST wave is kept small (its mix cannot hide behind any DMMA work): the lightest sectors up to tail_mb
    static double tail_mb = -1.0;
    if (tail_mb < 0) {
      const char* e = getenv("HTN_WAVE_TAIL_MB");
      tail_mb = e ? atof(e) : 0.0;
    }
    size_t ntail = 0;
    if (tail_mb > 0 && !wave_order) {
      double t = 0.0;
      while (ntail + 1 < order.size() && t + bytes[order[order.size() - 1 - ntail]] <= tail_mb * 1e6) {
        t += bytes[order[order.size() - 1 - ntail]];
        ++ntail;
      }
    }
    double acc = 0.0, cap = wave_mb * 1e6;
    int w = 1;
    for (size_t oi = 0; oi < order.size(); ++oi) {
      const int lp = order[oi];
      if (acc > 0 && (acc + bytes[lp] > cap || (ntail > 0 && oi == order.size() - ntail))) {
        ++w;
        acc = 0.0;
        cap = std::max(cap * wave_taper, 2e6);
      }
      F.wave_of_lp[lp] = w;
      acc += bytes[lp];
    }
    F.nwaves = w + 1;
    F.stacked = true;
    if (tile_waves_mode) {  // sectors without heavy T ride with the wave of their predecessor: waves stay monotone in lp
      int cur = 1;
      for (int lp = 0; lp < nlp; ++lp) {
        if (bytes[lp] > 0)
          cur = F.wave_of_lp[lp];
        else
          F.wave_of_lp[lp] = cur;
      }
    }
  }
  // ---- stacked T arrays and jobs ----
  for (auto& kv : by_xi) {
    const int xi = kv.first;
    const Block& xb = like->blocks[xi];
    const int l = std::get<2>(tkeys[kv.second[0]]);
    const htn_tensor::Panel& pn = G->panels[l];
    struct Row {
      int prow, rows, lp, ti;
    };
    std::vector<Row> rows;
    for (int ti : kv.second) {
      int a, lp, l2, x2;
      std::tie(a, lp, l2, x2) = tkeys[ti];
      rows.push_back(Row{G->block_prow[GL.find(a, lp, l2)], tb[ti].rows, lp, ti});
    }
    std::sort(rows.begin(), rows.end(), [](const Row& x, const Row& y) { return x.prow < y.prow; });
    const int span0 = rows.front().prow, span1 = rows.back().prow + rows.back().rows;
    const int ldT = even_up(xb.cols);
    const int64_t tarr = pg.ws_alloc((int64_t)(span1 - span0) * ldT);
    for (const Row& r : rows) tb[r.ti].off = tarr + (int64_t)(r.prow - span0) * ldT;
    // column pieces: the 8-column atoms split evenly into pieces of <= maxat atoms
    const int atoms = (xb.cols + 7) / 8, maxat = std::max(1, stack_max_atoms(pn.cols));
    const int npieces = (atoms + maxat - 1) / maxat;
    // row runs: consecutive needed blocks of one wave (gaps of < 32 unneeded rows are computed along)
    size_t i = 0;
    while (i < rows.size()) {
      size_t j = i;
      int r0 = rows[i].prow, r1 = rows[i].prow + rows[i].rows;
      const bool light = light_panel(l);
      const bool by_tile = tile_waves_mode && !light;
      const int wave = light ? 0 : (by_tile ? -1 : F.wave_of_lp[rows[i].lp]);
      while (j + 1 < rows.size() && (light || ((by_tile || F.wave_of_lp[rows[j + 1].lp] == wave) && rows[j + 1].prow - r1 < 32))) {
        ++j;
        r1 = rows[j].prow + rows[j].rows;
      }
      int c0 = 0;
      for (int pc = 0; pc < npieces; ++pc) {
        const int at = atoms / npieces + (pc < atoms % npieces ? 1 : 0);
        const int nt = std::min(at * 8, xb.cols - c0);
        // jobs of 6 tiles (measured best of 2..24: larger jobs delay the completion of their wave and with it the
        // mixers, smaller ones pay a slab reload each; HTN_STACK_FALL=1 cuts the end of every run finer)
        int m0 = r0;
        while (m0 < r1) {
          static int tpj = -1, fall = -1, tpj_tail = -1, tail_waves = 2, tpj_head = -1;
          if (tpj < 0) {
            const char* e = getenv("HTN_STACK_TPJ");  // tiles per job (tuning)
            tpj = e ? std::max(1, atoi(e)) : 6 * stack_gemm_groups();
            e = getenv("HTN_STACK_FALL");
            fall = e ? atoi(e) : 0;
            // a wave is complete one job duration after its last job was drawn: short jobs in the last waves bring the
            // end of the mix closer to the end of the DMMA work
            e = getenv("HTN_STACK_TPJ_TAIL");
            tpj_tail = e ? std::max(1, atoi(e)) : tpj;
            e = getenv("HTN_STACK_TAIL_WAVES");
            if (e) tail_waves = atoi(e);
            e = getenv("HTN_STACK_TPJ_HEAD");  // jobs of the first heavy wave: the mixers cannot start before it is complete
            tpj_head = e ? std::max(1, atoi(e)) : tpj;
          }
          const int rem_tiles = (r1 - m0 + 63) / 64;
          const int tpj_w = (light || by_tile) ? tpj : (wave <= 1 ? tpj_head : (wave >= F.nwaves - tail_waves ? tpj_tail : tpj));
          const int tiles = fall ? std::max(1, std::min(tpj_w, (rem_tiles + 2) / 3)) : tpj_w;
          const int M = std::min(tiles * 64, r1 - m0);
          StackJobH jb{};
          jb.A = Opnd{slot_gl, pn.off + (int64_t)m0 * pn.ld};
          jb.lda = pn.ld;
          jb.B = Opnd{slot_x, xb.off + c0};
          jb.ldb = xb.ld;
          jb.C = Opnd{SLOT_WS, tarr + (int64_t)(m0 - span0) * ldT + c0};
          jb.ldc = ldT;
          jb.K = pn.cols;
          jb.nt = nt;
          jb.nb = std::min(even_up(nt), xb.ld - c0);
          jb.M = M;
          jb.tmap = G->d_tmaps ? l : -1;
          jb.arow = m0;
          jb.wave = wave;
          if (by_tile) {  // waves of the T blocks under every 64-row tile of the job (rows i .. j of `rows`, sorted by prow)
            for (int t0 = m0; t0 < m0 + M; t0 += 64) {
              int lo = INT32_MAX, hi = -1;
              for (size_t q = i; q <= j; ++q)
                if (rows[q].prow < t0 + 64 && rows[q].prow + rows[q].rows > t0) {
                  lo = std::min(lo, F.wave_of_lp[rows[q].lp]);
                  hi = std::max(hi, F.wave_of_lp[rows[q].lp]);
                }
              if (hi < 0) lo = 0;  // a tile of gap rows only
              jb.tw.push_back({lo, hi});
            }
          }
          F.jobs.push_back(jb);
          m0 += M;
        }
        c0 += nt;
      }
      i = j + 1;
    }
  }
  // ---- classic tasks for the rest; algorithmic flops of all T blocks ----
  for (size_t ti = 0; ti < tkeys.size(); ++ti) {
    int a, lp, l, xi;
    std::tie(a, lp, l, xi) = tkeys[ti];
    const Block& xb = like->blocks[xi];
    const Block& gl = G->blocks[GL.find(a, lp, l)];
    if (is_stacked[ti]) {
      const double f = 2.0 * tb[ti].rows * tb[ti].cols * gl.cols;
      pg.flops += f;
      pg.flops_tag[TAG_L] += f;
      continue;
    }
    WsBlock& w = tb[ti];
    w.off = pg.ws_alloc((int64_t)w.rows * w.ld);
    GemmTaskH g;
    g.C = Opnd{SLOT_WS, w.off};
    g.ldc = w.ld;
    g.M = w.rows;
    g.N = w.cols;
    g.segs.push_back(GemmSegH{Opnd{slot_gl, gl.off}, gl.ld, Opnd{slot_x, xb.off}, xb.ld, gl.cols});
    tasksL.push_back(std::move(g));
  }
  // ---- mix sources of every U block ----
  std::map<std::tuple<int, int, int>, int> uindex;
  for (auto& kv : terms) {
    if (kv.second == 0.0) continue;
    const HTerm& t = kv.first;
    const Block& xb = like->blocks[t.xi];
    const Block& yb = like->blocks[t.yi];
    Src src;
    src.coef = kv.second;
    if (t.a == idL) {
      if (t.lp != t.l) continue;  // identity level is trivial
      src.o = Opnd{slot_x, xb.off};
    } else {
      src.o = Opnd{SLOT_WS, tb[tindex[TKey(t.a, t.lp, t.l, t.xi)]].off};
    }
    auto key = std::make_tuple(t.b, t.yi, t.r);
    auto it = uindex.find(key);
    int ui;
    if (it == uindex.end()) {
      ui = (int)F.ub.size();
      uindex[key] = ui;
      F.ukeys.push_back(key);
      F.ub.push_back(WsBlock{yb.rows, xb.cols, even_up(xb.cols), -1});
      F.usrc.emplace_back();
    } else
      ui = it->second;
    F.usrc[ui].push_back(src);
  }
  F.n_t = (int)tb.size();
  pg.add_gemm(tasksL, TAG_L);
  return F;
}

// terms of the one-site network  sum w N / dim(r') : H_AC (has_right = GR block) and T_L (has_right =
// output block)
static std::map<HTerm, double> one_site_terms(int sym, const htn_tensor* like, const EnvView& GL, const htn_mpo* W,
                                              const std::function<bool(int, int, int)>& has_right /*(b,r,rp)*/) {
  const auto& Vl = like->s0;
  const auto& Vr = like->s1;
  const auto& P = like->legs;
  std::map<std::pair<int, int>, std::vector<int>> pl;  // (a,l) -> l'
  const int nl = (int)Vl.sec.size(), nr = (int)Vr.sec.size();
  const int nlev_l = (int)W->Ml.sec.size(), nlev_r = (int)W->Mr.sec.size();
  for (int a = 0; a < nlev_l; ++a)
    for (int lp = 0; lp < nl; ++lp)
      for (int l = 0; l < nl; ++l)
        if (GL.find(a, lp, l) >= 0) pl[{a, l}].push_back(lp);
  std::map<std::pair<int, int>, std::vector<int>> pr;  // (b,r) -> r'
  for (int b = 0; b < nlev_r; ++b)
    for (int r = 0; r < nr; ++r)
      for (int rp = 0; rp < nr; ++rp)
        if (has_right(b, r, rp)) pr[{b, r}].push_back(rp);
  std::vector<std::vector<std::pair<int, int>>> xs(P.sec.size());
  for (const Block& b : like->blocks) xs[b.lab[1]].push_back({b.lab[0], b.lab[2]});
  std::map<HTerm, double> terms;
  for (const MpoEntry& e : W->entries) {
    for (auto& lr : xs[e.s]) {
      int l = lr.first, r = lr.second;
      auto itl = pl.find({e.a, l});
      auto itr = pr.find({e.b, r});
      if (itl == pl.end() || itr == pr.end()) continue;
      for (int lp : itl->second)
        for (int rp : itr->second) {
          const int yi = like->find(lp, e.sp, rp);
          if (yi < 0) continue;
          double n = network(sym, Vl.sec[lp], P.sec[e.sp], Vr.sec[rp], Vl.sec[l], P.sec[e.s], Vr.sec[r],
                             W->Ml.sec[e.a], W->Mr.sec[e.b], e.c);
          if (n == 0.0) continue;
          terms[HTerm{yi, like->find(l, e.s, r), e.a, lp, l, e.b, r, rp}] += e.w * n / sdim(sym, Vr.sec[rp]);
        }
    }
  }
  return terms;
}

// Multi-GPU sharding of ONE effective-Hamiltonian application: the apply is a sum of independent
// terms, so any partition of the term list gives partial results y_k with y = sum_k y_k (one allreduce).  Units of
// the partition are the LEFT SECTORS lp of the output: all three stages of a sector (GL rows, T, U, y blocks) stay on
// one device, nothing is computed twice, and a device only touches the GL panels rows of its sectors.  A sector that
// is heavier than a fair share (at D=1024 four sectors hold 80 % of the work) is split further by MPO level a of
// its GL blocks; stage R of such a sector then runs on each of its owners (on their partial U).
static void shard_terms(std::map<HTerm, double>& terms, const htn_tensor* like, int idL, int nshards, int shard) {
  if (nshards <= 1) return;
  const auto& Vl = like->s0;
  const auto& Vr = like->s1;
  const int nlp = (int)Vl.sec.size();
  auto rdim = [&](int yi) { return like->kind == HTN_T_MPS ? like->blocks[yi].lab[2] : like->blocks[yi].lab[4]; };
  std::map<std::tuple<int, int, int, int>, char> tseen;  // (a,lp,l,xi)
  std::map<std::tuple<int, int, int>, char> useen;       // (b,yi,r)
  std::vector<double> costR(nlp, 0.0), costL(nlp, 0.0);
  std::map<std::pair<int, int>, double> costLa;  // (lp,a)
  for (auto& kv : terms) {
    if (kv.second == 0.0) continue;
    const HTerm& t = kv.first;
    const double nlpv = Vl.mult[t.lp];
    if (t.a != idL && !tseen.count({t.a, t.lp, t.l, t.xi})) {
      tseen[{t.a, t.lp, t.l, t.xi}] = 1;
      const double c = nlpv * Vl.mult[t.l] * like->blocks[t.xi].cols;
      costL[t.lp] += c;
      costLa[{t.lp, t.a}] += c;
    }
    if (!useen.count({t.b, t.yi, t.r})) {
      useen[{t.b, t.yi, t.r}] = 1;
      costR[t.lp] += nlpv * like->blocks[t.xi].cols * Vr.mult[rdim(t.yi)];
    }
  }
  double total = 0.0;
  for (int lp = 0; lp < nlp; ++lp) total += costL[lp] + costR[lp];
  const double fair = total / nshards;
  struct Unit {
    int lp, part, nparts;
    double cost;
  };
  std::vector<Unit> units;
  std::map<std::pair<int, int>, int> part_of;  // (lp,a) -> part
  for (int lp = 0; lp < nlp; ++lp) {
    const double c = costL[lp] + costR[lp];
    if (c <= 0) continue;
    int k = c > 1.15 * fair ? (int)std::ceil(c / fair) : 1;
    std::vector<std::pair<double, int>> la;  // levels of this sector by falling cost
    for (auto& kv : costLa)
      if (kv.first.first == lp) la.push_back({kv.second, kv.first.second});
    k = std::max(1, std::min<int>(k, (int)la.size()));
    std::sort(la.begin(), la.end(), [](auto& x, auto& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
    std::vector<double> load(k, 0.0);
    for (auto& pr : la) {
      const int i = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      load[i] += pr.first;
      part_of[{lp, pr.second}] = i;
    }
    for (int i = 0; i < k; ++i) units.push_back(Unit{lp, i, k, load[i] + costR[lp]});
  }
  std::stable_sort(units.begin(), units.end(), [](const Unit& x, const Unit& y) { return x.cost > y.cost; });
  std::vector<double> sload(nshards, 0.0);
  std::map<std::pair<int, int>, int> owner;  // (lp,part) -> shard
  for (const Unit& u : units) {
    const int i = (int)(std::min_element(sload.begin(), sload.end()) - sload.begin());
    sload[i] += u.cost;
    owner[{u.lp, u.part}] = i;
  }
  for (auto it = terms.begin(); it != terms.end();) {
    const HTerm& t = it->first;
    auto pit = part_of.find({t.lp, t.a});
    const int part = pit == part_of.end() ? 0 : pit->second;  // identity-level terms ride with part 0
    auto oit = owner.find({t.lp, part});
    const int own = oit == owner.end() ? 0 : oit->second;
    if (own != shard)
      it = terms.erase(it);
    else
      ++it;
  }
}

// Back end shared by H_AC and H_AC2: U blocks (stage W), stage R with split-K, final mix into y.
// slots: 0 = x, 1 = y, 2 = GL, 3 = GR
static void build_heff_backend(Program& pg, const htn_tensor* like, const htn_tensor* GR, LeftFront& F, int* n_u_out,
                               int* n_mix_t, int* n_mix_s_out) {
  const int idR = GR->identity_level;
  std::vector<MixTaskH> mixU;
  std::vector<std::vector<MixSrcH>> yextra(like->blocks.size());
  std::vector<GemmTaskH> tasksR(like->blocks.size());
  int n_u = 0, n_mix_s = 0;
  for (size_t yi = 0; yi < like->blocks.size(); ++yi) {
    const Block& yb = like->blocks[yi];
    tasksR[yi].C = Opnd{1, yb.off};
    tasksR[yi].ldc = yb.ld;
    tasksR[yi].M = yb.rows;
    tasksR[yi].N = yb.cols;
  }
  std::vector<int> order(F.ukeys.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::sort(order.begin(), order.end(), [&](int i, int j) { return F.ukeys[i] < F.ukeys[j]; });
  std::map<int, MixTaskH> y0;  // stacked front: direct (identity right level) contributions, formed by the mixer warps
  for (int ui : order) {
    int b, yi, r;
    std::tie(b, yi, r) = F.ukeys[ui];
    n_mix_s += (int)F.usrc[ui].size();
    const int lp = like->blocks[yi].lab[0];
    if (b == idR) {
      if (!F.stacked) {
        for (const Src& s : F.usrc[ui]) yextra[yi].push_back(MixSrcH{s.o, s.coef});
      } else {
        MixTaskH& t = y0[yi];
        for (const Src& s : F.usrc[ui]) t.srcs.push_back(MixSrcH{s.o, s.coef});
      }
      continue;
    }
    WsBlock& w = F.ub[ui];
    w.off = pg.ws_alloc((int64_t)w.rows * w.ld);
    MixTaskH mt;
    mt.dst = Opnd{SLOT_WS, w.off};
    mt.nelem = w.rows * w.ld;
    mt.wave = F.stacked ? F.wave_of_lp[lp] : -1;  // (wave 0 = light panels only: the mixers always wait for it too)
    for (const Src& s : F.usrc[ui]) mt.srcs.push_back(MixSrcH{s.o, s.coef});
    mixU.push_back(std::move(mt));
    const int rp = like->kind == HTN_T_MPS ? like->blocks[yi].lab[2] : like->blocks[yi].lab[4];
    const Block& gr = GR->blocks[GR->find(b, r, rp)];
    tasksR[yi].segs.push_back(GemmSegH{Opnd{SLOT_WS, w.off}, w.ld, Opnd{3, gr.off}, gr.ld, gr.rows});
    ++n_u;
  }
  // direct (identity right level) contributions: ~60 sources per y block.  A mix chunk is one warp's serial chain over
  // its sources, so the list is cut into partial sums of <= Y0_GROUP sources; the final reduce of stage R adds them up.
  static int y0_group = -1;
  if (y0_group < 0) {
    const char* e = getenv("HTN_Y0_GROUP");
    y0_group = e ? std::max(1, atoi(e)) : 16;
  }
  for (auto& kv : y0) {
    const Block& yb = like->blocks[kv.first];
    MixTaskH& all = kv.second;
    for (size_t s0 = 0; s0 < all.srcs.size(); s0 += (size_t)y0_group) {
      MixTaskH t;
      t.srcs.assign(all.srcs.begin() + s0, all.srcs.begin() + std::min(all.srcs.size(), s0 + (size_t)y0_group));
      const int64_t off = pg.ws_alloc((int64_t)yb.rows * yb.ld);
      t.dst = Opnd{SLOT_WS, off};
      t.nelem = yb.rows * yb.ld;
      t.wave = F.wave_of_lp[yb.lab[0]];
      yextra[kv.first].push_back(MixSrcH{t.dst, 1.0});
      mixU.push_back(std::move(t));
    }
  }
  *n_mix_t = (int)(mixU.size() + like->blocks.size());
  if (F.stacked) {
    pg.add_stack(F.jobs, mixU, F.nwaves, 2, TAG_L);
  } else {
    // launch order of the U mixes: a T block feeds ~1.7 U blocks (other MPO levels b, other s'); walking the targets
    // in the order of their first source puts those readers next to each other in time, so the repeat reads of T hit L2
    if (!getenv("HTN_MIX_ORDER_BY_TARGET")) {
      auto first_src = [](const MixTaskH& t) {
        int64_t m = INT64_MAX;
        for (const MixSrcH& q : t.srcs)
          if (q.src.slot == SLOT_WS) m = std::min<int64_t>(m, q.src.off);
        return m;
      };
      std::stable_sort(mixU.begin(), mixU.end(), [&](const MixTaskH& a, const MixTaskH& b) { return first_src(a) < first_src(b); });
    }
    pg.add_mix(mixU, TAG_W);
  }
  pg.add_gemm_reduce(tasksR, yextra, TAG_R, TAG_Y, true);
  *n_u_out = n_u;
  *n_mix_s_out = n_mix_s;
}

}  // namespace htn

extern "C" {

int32_t htn_plan_destroy(htn_plan* p) {
  if (!p) return HTN_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  p->prog.destroy();
  if (p->like_in) htn_tensor_destroy(p->like_in);
  if (p->like_out) htn_tensor_destroy(p->like_out);
  if (p->hx) htn_tensor_destroy(p->hx);
  if (p->hy) htn_tensor_destroy(p->hy);
  delete p;
  return HTN_OK;
}

static void fill_stats(htn_plan* p, int n_t, int n_u, int n_mix_t, int n_mix_s) {
  const Program& pg = p->prog;
  p->stats[0] = pg.flops;
  p->stats[1] = pg.flops_tag[TAG_L];
  p->stats[2] = pg.flops_tag[TAG_R];
  p->stats[3] = n_t;
  p->stats[4] = n_u;
  p->stats[5] = n_mix_t;
  p->stats[6] = n_mix_s;
  p->stats[7] = (double)pg.ws_elems * sizeof(double);
  p->stats[8] = pg.n_gemm_tiles_tag[TAG_L];
  p->stats[9] = pg.n_gemm_tiles_tag[TAG_R];
  p->stats[10] = pg.padded_flops;
  p->stats[11] = pg.launches();
}

// slots: 0 = x, 1 = y, 2 = GL, 3 = GR
int32_t htn_plan_heff_ac(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W, const htn_tensor* GR,
                         const htn_tensor* like, htn_plan** out) {
  return htn_plan_heff_ac_sharded(ctx, GL, W, GR, like, 1, 0, out);
}

// shard `shard` of `nshards`: this plan computes a partial y; the sum over the shards (one allreduce) is H_AC x
int32_t htn_plan_heff_ac_sharded(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W, const htn_tensor* GR,
                                 const htn_tensor* like, int32_t nshards, int32_t shard, htn_plan** out) {
  if (!ctx || !GL || !W || !GR || !like || !out) return HTN_ERR_INVALID;
  *out = nullptr;
  if (nshards < 1 || shard < 0 || shard >= nshards) return ctx->fail(HTN_ERR_INVALID, "plan_heff_ac: shard index out of range");
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

  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  htn_plan* p = nullptr;
  try {
    p = new htn_plan();
    p->ctx = ctx;
    p->kind = HTN_PLAN_HEFF_AC;
    p->like_in = like_copy;
    p->bound[2] = GL;
    p->bound[3] = GR;
    Program& pg = p->prog;
    EnvView gl{GL};
    auto terms = one_site_terms(sym, like, gl, W, [&](int b, int r, int rp) { return GR->find(b, r, rp) >= 0; });
    shard_terms(terms, like, gl.identity_level(), nshards, shard);
    LeftFront F = build_front(pg, like, gl, terms, 0, 2, true);
    int n_u = 0, n_mix_t = 0, n_mix_s = 0;
    build_heff_backend(pg, like, GR, F, &n_u, &n_mix_t, &n_mix_s);
    if ((rc = pg.finalize(ctx, 4))) {
      htn_plan_destroy(p);
      return rc;
    }
    fill_stats(p, F.n_t, n_u, n_mix_t, n_mix_s);
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

// Two-site effective Hamiltonian: two nested copies of the one-site recoupling network
// (oracle/twosite.py:HeffAC2Plan).  slots: 0 = x2, 1 = y2, 2 = GL, 3 = GR
int32_t htn_plan_heff_ac2(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W1, const htn_mpo* W2,
                          const htn_tensor* GR, const htn_tensor* like, htn_plan** out) {
  if (!ctx || !GL || !W1 || !W2 || !GR || !like || !out) return HTN_ERR_INVALID;
  *out = nullptr;
  if (GL->kind != HTN_T_ENVL || GR->kind != HTN_T_ENVR || like->kind != HTN_T_MPS2)
    return ctx->fail(HTN_ERR_INVALID, "plan_heff_ac2: wrong tensor kinds");
  const int sym = like->sym;
  if (GL->s0.sec != like->s0.sec || GL->s0.mult != like->s0.mult || GR->s0.sec != like->s1.sec ||
      GR->s0.mult != like->s1.mult)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_ac2: environment bond spaces do not match x2");
  if (GL->legs.sec != W1->Ml.sec || W1->Mr.sec != W2->Ml.sec || GR->legs.sec != W2->Mr.sec ||
      like->legs.sec != W1->P.sec || like->legs2.sec != W2->P.sec)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_ac2: MPO legs do not match");
  htn_tensor* like_copy = nullptr;
  int32_t rc = htn_tensor_create_like(like, &like_copy);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  htn_plan* p = nullptr;
  try {
    p = new htn_plan();
    p->ctx = ctx;
    p->kind = HTN_PLAN_HEFF_AC2;
    p->like_in = like_copy;
    p->bound[2] = GL;
    p->bound[3] = GR;
    Program& pg = p->prog;
    EnvView gl{GL};
    const auto& Vl = like->s0;
    const auto& Vr = like->s1;
    const auto& P1 = like->legs;
    const auto& P2 = like->legs2;
    std::map<std::pair<int, int>, std::vector<int>> pl, pr;
    for (const Block& b : GL->blocks) pl[{b.lab[0], b.lab[2]}].push_back(b.lab[1]);
    for (const Block& b : GR->blocks) pr[{b.lab[0], b.lab[1]}].push_back(b.lab[2]);
    std::vector<std::vector<const MpoEntry*>> w1_by_s(P1.sec.size());
    for (const MpoEntry& e : W1->entries) w1_by_s[e.s].push_back(&e);
    std::map<std::pair<int, int>, std::vector<const MpoEntry*>> w2_by_bs;
    for (const MpoEntry& e : W2->entries) w2_by_bs[{e.a, e.s}].push_back(&e);
    std::map<std::tuple<int, int, int>, int> mid_index;
    for (size_t i = 0; i < like->mid.size(); ++i) mid_index[{like->mid[i].p, like->mid[i].q, like->mid[i].n}] = (int)i;
    std::map<HTerm, double> terms;
    for (size_t xi = 0; xi < like->blocks.size(); ++xi) {
      const Block& xb = like->blocks[xi];
      const int l = xb.lab[0], s1 = xb.lab[1], m = xb.lab[2], s2 = xb.lab[3], r = xb.lab[4];
      const Sector cm = like->mid[m];
      for (const MpoEntry* e1 : w1_by_s[s1]) {
        auto itl = pl.find({e1->a, l});
        if (itl == pl.end()) continue;
        auto it2 = w2_by_bs.find({e1->b, s2});
        if (it2 == w2_by_bs.end()) continue;
        for (int lp : itl->second) {
          const Sector clp = Vl.sec[lp], csp = P1.sec[e1->sp];
          const int p2 = (clp.p + csp.p) & 1, n2 = clp.n + csp.n;
          const int qlo = sym == HTN_SYM_SU2U1 ? std::abs(clp.q - csp.q) : clp.q + csp.q;
          const int qhi = clp.q + csp.q;
          for (int q = qlo; q <= qhi; q += 2) {
            const Sector cmp{p2, q, n2};
            auto itm = mid_index.find({cmp.p, cmp.q, cmp.n});
            if (itm == mid_index.end()) continue;
            const double na = network(sym, clp, csp, cmp, Vl.sec[l], P1.sec[s1], cm, W1->Ml.sec[e1->a], W1->Mr.sec[e1->b], e1->c);
            if (na == 0.0) continue;
            for (const MpoEntry* e2 : it2->second) {
              auto itr = pr.find({e2->b, r});
              if (itr == pr.end()) continue;
              for (int rp : itr->second) {
                const int yi = like->find5(lp, e1->sp, itm->second, e2->sp, rp);
                if (yi < 0) continue;
                const double nb = network(sym, cmp, P2.sec[e2->sp], Vr.sec[rp], cm, P2.sec[s2], Vr.sec[r], W2->Ml.sec[e2->a],
                                          W2->Mr.sec[e2->b], e2->c);
                if (nb == 0.0) continue;
                terms[HTerm{yi, (int)xi, e1->a, lp, l, e2->b, r, rp}] +=
                    (e1->w * na / sdim(sym, cmp)) * (e2->w * nb / sdim(sym, Vr.sec[rp]));
              }
            }
          }
        }
      }
    }
    LeftFront F = build_front(pg, like, gl, terms, 0, 2, true);
    int n_u = 0, n_mix_t = 0, n_mix_s = 0;
    build_heff_backend(pg, like, GR, F, &n_u, &n_mix_t, &n_mix_s);
    if ((rc = pg.finalize(ctx, 4))) {
      htn_plan_destroy(p);
      return rc;
    }
    fill_stats(p, F.n_t, n_u, n_mix_t, n_mix_s);
  } catch (const std::bad_alloc&) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_OOM, "plan_heff_ac2: host allocation failed");
  } catch (const std::exception& e) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_INVALID, std::string("plan_heff_ac2: ") + e.what());
  }
  *out = p;
  return HTN_OK;
}

// x2[l,s1,m,s2,r] = A1[l,s1,m] . A2[m,s2,r]
int32_t htn_contract_two_site(const htn_tensor* A1, const htn_tensor* A2, htn_tensor* x2) {
  if (!A1 || !A2 || !x2) return HTN_ERR_INVALID;
  htn_ctx* ctx = A1->ctx;
  if (A1->kind != HTN_T_MPS || A2->kind != HTN_T_MPS || x2->kind != HTN_T_MPS2)
    return ctx->fail(HTN_ERR_INVALID, "contract_two_site: wrong tensor kinds");
  if (A1->s1.sec != A2->s0.sec || A1->s1.mult != A2->s0.mult || A1->s0.sec != x2->s0.sec || A1->s0.mult != x2->s0.mult ||
      A2->s1.sec != x2->s1.sec || A2->s1.mult != x2->s1.mult || A1->legs.sec != x2->legs.sec || A2->legs.sec != x2->legs2.sec)
    return ctx->fail(HTN_ERR_SHAPE, "contract_two_site: spaces do not match");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  try {
    Program pg;
    std::vector<GemmTaskH> tasks;
    for (const Block& b : x2->blocks) {
      GemmTaskH t;
      t.C = Opnd{2, b.off};
      t.ldc = b.ld;
      t.M = b.rows;
      t.N = b.cols;
      const Sector cm = x2->mid[b.lab[2]];
      int mi = -1;
      for (size_t i = 0; i < A1->s1.sec.size(); ++i)
        if (A1->s1.sec[i] == cm) mi = (int)i;
      if (mi >= 0) {
        const Block& a1 = A1->blocks[A1->find(b.lab[0], b.lab[1], mi)];
        const Block& a2 = A2->blocks[A2->find(mi, b.lab[3], b.lab[4])];
        t.segs.push_back(GemmSegH{Opnd{0, a1.off}, a1.ld, Opnd{1, a2.off}, a2.ld, a1.cols});
      }
      tasks.push_back(std::move(t));
    }
    pg.add_gemm(tasks, TAG_L);
    int32_t rc = pg.finalize(ctx, 3);
    if (rc == HTN_OK) {
      const double* slots[3] = {A1->d, A2->d, x2->d};
      rc = pg.run(slots);
    }
    cudaStreamSynchronize(ctx->stream);
    pg.destroy();
    return rc;
  } catch (const std::exception& e) {
    return ctx->fail(HTN_ERR_INVALID, std::string("contract_two_site: ") + e.what());
  }
}

// slots: 0 = x (C), 1 = y, 2 = GL, 3 = GR.   GL lives on the bond of C (left env of the next
// site), GR on the same bond (right env of this site).
int32_t htn_plan_heff_c(htn_ctx* ctx, const htn_tensor* GL, const htn_tensor* GR, const htn_tensor* like,
                        htn_plan** out) {
  if (!ctx || !GL || !GR || !like || !out) return HTN_ERR_INVALID;
  *out = nullptr;
  if (GL->kind != HTN_T_ENVL || GR->kind != HTN_T_ENVR || like->kind != HTN_T_BOND)
    return ctx->fail(HTN_ERR_INVALID, "plan_heff_c: wrong tensor kinds");
  if (GL->s0.sec != like->s0.sec || GL->s0.mult != like->s0.mult || GR->s0.sec != like->s0.sec ||
      GR->s0.mult != like->s0.mult || GL->legs.sec != GR->legs.sec)
    return ctx->fail(HTN_ERR_SHAPE, "plan_heff_c: spaces / MPO levels of GL, GR and C do not match");
  htn_tensor* like_copy = nullptr;
  int32_t rc = htn_tensor_create_like(like, &like_copy);
  if (rc) return rc;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  htn_plan* p = nullptr;
  try {
    p = new htn_plan();
    p->ctx = ctx;
    p->kind = HTN_PLAN_HEFF_C;
    p->like_in = like_copy;
    p->bound[2] = GL;
    p->bound[3] = GR;
    Program& pg = p->prog;
    const int idL = GL->identity_level, idR = GR->identity_level;
    const int nc = (int)like->blocks.size();
    std::vector<GemmTaskH> tasksL, tasksR(nc);
    std::vector<std::vector<MixSrcH>> yextra(nc);
    for (int c = 0; c < nc; ++c) {
      const Block& yb = like->blocks[c];
      tasksR[c].C = Opnd{1, yb.off};
      tasksR[c].ldc = yb.ld;
      tasksR[c].M = yb.rows;
      tasksR[c].N = yb.cols;
    }
    int n_t = 0, n_u = 0;
    // y[c'] = sum_{a,c} GL[a,c',c] C[c] GR[a,c,c']   (all recoupling coefficients are 1)
    for (const Block& gl : GL->blocks) {
      const int a = gl.lab[0], cp = gl.lab[1], c = gl.lab[2];
      const int gri = GR->find(a, c, cp);
      if (gri < 0) continue;
      const Block& xb = like->blocks[c];
      Opnd t;
      int tld;
      if (a == idL) {
        t = Opnd{0, xb.off};
        tld = xb.ld;
      } else {
        const int ld = even_up(xb.cols);
        const int64_t off = pg.ws_alloc((int64_t)gl.rows * ld);
        GemmTaskH tk;
        tk.C = Opnd{SLOT_WS, off};
        tk.ldc = ld;
        tk.M = gl.rows;
        tk.N = xb.cols;
        tk.segs.push_back(GemmSegH{Opnd{2, gl.off}, gl.ld, Opnd{0, xb.off}, xb.ld, gl.cols});
        tasksL.push_back(std::move(tk));
        t = Opnd{SLOT_WS, off};
        tld = ld;
        ++n_t;
      }
      if (a == idR) {
        yextra[cp].push_back(MixSrcH{t, 1.0});  // c == c' for the trivial level
      } else {
        const Block& gr = GR->blocks[gri];
        tasksR[cp].segs.push_back(GemmSegH{t, tld, Opnd{3, gr.off}, gr.ld, gr.rows});
        ++n_u;
      }
    }
    pg.add_gemm(tasksL, TAG_L);
    pg.add_gemm_reduce(tasksR, yextra, TAG_R, TAG_Y);
    if ((rc = pg.finalize(ctx, 4))) {
      htn_plan_destroy(p);
      return rc;
    }
    fill_stats(p, n_t, n_u, nc, 0);
  } catch (const std::bad_alloc&) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_OOM, "plan_heff_c: host allocation failed");
  } catch (const std::exception& e) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_INVALID, std::string("plan_heff_c: ") + e.what());
  }
  *out = p;
  return HTN_OK;
}

// Transfer plans.  slots: 0 = A, 1 = A^T (plain blockwise transpose, kind MPST), 2 = env in, 3 = env out.
// `env_in` / `env_out` fix the structure (ENVL/ENVR, or BOND tensors acting as one-level environments
// for the MPO-free transfer matrix of the gauge / GMRES steps).
int32_t htn_plan_transfer(htn_ctx* ctx, int32_t side, const htn_mpo* W, const htn_tensor* A, const htn_tensor* At,
                          const htn_tensor* env_in, const htn_tensor* env_out, htn_plan** out) {
  if (!ctx || !W || !A || !At || !env_in || !env_out || !out) return HTN_ERR_INVALID;
  *out = nullptr;
  if (A->kind != HTN_T_MPS || At->kind != HTN_T_MPST) return ctx->fail(HTN_ERR_INVALID, "plan_transfer: A / A^T kinds");
  const int sym = A->sym;
  const bool left = side == HTN_SIDE_LEFT;
  const int want = left ? HTN_T_ENVL : HTN_T_ENVR;
  for (const htn_tensor* e : {env_in, env_out})
    if (e->kind != want && e->kind != HTN_T_BOND) return ctx->fail(HTN_ERR_INVALID, "plan_transfer: environment kind");
  const htn_space& Vin = left ? A->s0 : A->s1;
  const htn_space& Vout = left ? A->s1 : A->s0;
  if (env_in->s0.sec != Vin.sec || env_in->s0.mult != Vin.mult || env_out->s0.sec != Vout.sec ||
      env_out->s0.mult != Vout.mult)
    return ctx->fail(HTN_ERR_SHAPE, "plan_transfer: environment bond spaces do not match A");
  const htn_legs& Min = left ? W->Ml : W->Mr;
  const htn_legs& Mout = left ? W->Mr : W->Ml;
  if ((env_in->kind != HTN_T_BOND && env_in->legs.sec != Min.sec) ||
      (env_out->kind != HTN_T_BOND && env_out->legs.sec != Mout.sec) ||
      (env_in->kind == HTN_T_BOND && Min.sec.size() != 1) || (env_out->kind == HTN_T_BOND && Mout.sec.size() != 1) ||
      A->legs.sec != W->P.sec)
    return ctx->fail(HTN_ERR_SHAPE, "plan_transfer: MPO legs do not match");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  htn_plan* p = nullptr;
  int32_t rc;
  try {
    p = new htn_plan();
    p->ctx = ctx;
    p->kind = left ? HTN_PLAN_TRANSFER_L : HTN_PLAN_TRANSFER_R;
    Program& pg = p->prog;
    EnvView ein{env_in}, eout{env_out};
    const auto& Vl = A->s0;
    const auto& Vr = A->s1;
    const auto& P = A->legs;
    int n_t = 0, n_u = 0, n_mix_s = 0;
    if (left) {
      // front end identical to H_AC with x = A; the "right partner" test is the output block
      auto terms = one_site_terms(sym, A, ein, W, [&](int b, int r, int rp) { return eout.find(b, rp, r) >= 0; });
      LeftFront F = build_front(pg, A, ein, terms, 0, 2);
      n_t = F.n_t;
      std::vector<MixTaskH> mixU;
      std::map<int, GemmTaskH> tasks;  // by output block index
      std::vector<int> order(F.ukeys.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
      std::sort(order.begin(), order.end(), [&](int i, int j) { return F.ukeys[i] < F.ukeys[j]; });
      for (int ui : order) {
        int b, yi, r;
        std::tie(b, yi, r) = F.ukeys[ui];
        const int lp = A->blocks[yi].lab[0], sp = A->blocks[yi].lab[1], rp = A->blocks[yi].lab[2];
        WsBlock& w = F.ub[ui];
        w.off = pg.ws_alloc((int64_t)w.rows * w.ld);
        MixTaskH mt;
        mt.dst = Opnd{SLOT_WS, w.off};
        mt.nelem = w.rows * w.ld;
        for (const Src& s : F.usrc[ui]) mt.srcs.push_back(MixSrcH{s.o, s.coef});
        n_mix_s += (int)mt.srcs.size();
        mixU.push_back(std::move(mt));
        ++n_u;
        const int oi = eout.find(b, rp, r);
        const Block& at = At->blocks[At->find(lp, sp, rp)];  // [n_rp x n_lp]
        tasks[oi].segs.push_back(GemmSegH{Opnd{1, at.off}, at.ld, Opnd{SLOT_WS, w.off}, w.ld, at.cols});
      }
      pg.add_mix(mixU, TAG_W);
      std::vector<GemmTaskH> tasksR;
      for (size_t oi = 0; oi < env_out->blocks.size(); ++oi) {
        const Block& ob = env_out->blocks[oi];
        GemmTaskH t = tasks.count((int)oi) ? tasks[(int)oi] : GemmTaskH{};
        t.C = Opnd{3, ob.off};
        t.ldc = ob.ld;
        t.M = ob.rows;
        t.N = ob.cols;
        tasksR.push_back(std::move(t));
      }
      pg.add_gemm(tasksR, TAG_R);
    } else {
      // T[b,l,s,r,rp] = A[l,s,r] . GR[b,r,rp]   (n_l x n_rp);  b = identity level -> T = A block
      const int idR = ein.identity_level();
      std::map<std::tuple<int, int, int, int, int>, int> tindex;
      std::vector<WsBlock> tb;
      std::vector<GemmTaskH> tasks1;
      typedef std::tuple<int, int, int, int, int> UKey;  // (a,l,lp,sp,rp)
      std::map<UKey, std::vector<Src>> usrc;
      std::map<std::pair<int, int>, std::vector<int>> pr;  // (b,r) -> rp
      const int nr = (int)Vr.sec.size(), nlv = (int)Vl.sec.size();
      for (int b = 0; b < (int)W->Mr.sec.size(); ++b)
        for (int r = 0; r < nr; ++r)
          for (int rp = 0; rp < nr; ++rp)
            if (ein.find(b, r, rp) >= 0) pr[{b, r}].push_back(rp);
      std::vector<std::vector<std::pair<int, int>>> xs(P.sec.size());
      for (const Block& b : A->blocks) xs[b.lab[1]].push_back({b.lab[0], b.lab[2]});
      std::map<std::tuple<int, int, int, int, int, int, int, int>, double> terms;  // a,l,lp,sp,rp | b,s,r
      for (const MpoEntry& e : W->entries)
        for (auto& lr : xs[e.s]) {
          const int l = lr.first, r = lr.second;
          auto itr = pr.find({e.b, r});
          if (itr == pr.end()) continue;
          for (int rp : itr->second)
            for (int lp = 0; lp < nlv; ++lp) {
              if (A->find(lp, e.sp, rp) < 0 || eout.find(e.a, l, lp) < 0) continue;
              double n = network(sym, Vl.sec[lp], P.sec[e.sp], Vr.sec[rp], Vl.sec[l], P.sec[e.s], Vr.sec[r],
                                 W->Ml.sec[e.a], W->Mr.sec[e.b], e.c);
              if (n == 0.0) continue;
              terms[std::make_tuple(e.a, l, lp, e.sp, rp, e.b, e.s, r)] += e.w * n / sdim(sym, Vl.sec[lp]);
            }
        }
      for (auto& kv : terms) {
        if (kv.second == 0.0) continue;
        int a, l, lp, sp, rp, b, s, r;
        std::tie(a, l, lp, sp, rp, b, s, r) = kv.first;
        Src src;
        src.coef = kv.second;
        const Block& ab = A->blocks[A->find(l, s, r)];
        if (b == idR) {
          if (r != rp) continue;
          src.o = Opnd{0, ab.off};
        } else {
          auto key = std::make_tuple(b, l, s, r, rp);
          auto it = tindex.find(key);
          int ti;
          if (it == tindex.end()) {
            ti = (int)tb.size();
            tindex[key] = ti;
            const Block& gr = env_in->blocks[ein.find(b, r, rp)];
            WsBlock w{ab.rows, gr.cols, even_up(gr.cols), 0};
            w.off = pg.ws_alloc((int64_t)w.rows * w.ld);
            tb.push_back(w);
            GemmTaskH t;
            t.C = Opnd{SLOT_WS, w.off};
            t.ldc = w.ld;
            t.M = w.rows;
            t.N = w.cols;
            t.segs.push_back(GemmSegH{Opnd{0, ab.off}, ab.ld, Opnd{2, gr.off}, gr.ld, ab.cols});
            tasks1.push_back(std::move(t));
          } else
            ti = it->second;
          src.o = Opnd{SLOT_WS, tb[ti].off};
        }
        usrc[UKey(a, l, lp, sp, rp)].push_back(src);
      }
      n_t = (int)tb.size();
      pg.add_gemm(tasks1, TAG_L);
      std::vector<MixTaskH> mixU;
      std::map<int, GemmTaskH> tasks;
      for (auto& kv : usrc) {
        int a, l, lp, sp, rp;
        std::tie(a, l, lp, sp, rp) = kv.first;
        const int rows = Vl.mult[l], cols = Vr.mult[rp], ld = even_up(cols);
        const int64_t off = pg.ws_alloc((int64_t)rows * ld);
        MixTaskH mt;
        mt.dst = Opnd{SLOT_WS, off};
        mt.nelem = rows * ld;
        for (const Src& s : kv.second) mt.srcs.push_back(MixSrcH{s.o, s.coef});
        n_mix_s += (int)mt.srcs.size();
        mixU.push_back(std::move(mt));
        ++n_u;
        const int oi = eout.find(a, l, lp);
        const Block& at = At->blocks[At->find(lp, sp, rp)];  // [n_rp x n_lp]
        tasks[oi].segs.push_back(GemmSegH{Opnd{SLOT_WS, off}, ld, Opnd{1, at.off}, at.ld, at.rows});
      }
      pg.add_mix(mixU, TAG_W);
      std::vector<GemmTaskH> tasksR;
      for (size_t oi = 0; oi < env_out->blocks.size(); ++oi) {
        const Block& ob = env_out->blocks[oi];
        GemmTaskH t = tasks.count((int)oi) ? tasks[(int)oi] : GemmTaskH{};
        t.C = Opnd{3, ob.off};
        t.ldc = ob.ld;
        t.M = ob.rows;
        t.N = ob.cols;
        tasksR.push_back(std::move(t));
      }
      pg.add_gemm(tasksR, TAG_R);
    }
    if ((rc = pg.finalize(ctx, 4))) {
      htn_plan_destroy(p);
      return rc;
    }
    fill_stats(p, n_t, n_u, n_u, n_mix_s);
  } catch (const std::bad_alloc&) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_OOM, "plan_transfer: host allocation failed");
  } catch (const std::exception& e) {
    if (p) htn_plan_destroy(p);
    return ctx->fail(HTN_ERR_INVALID, std::string("plan_transfer: ") + e.what());
  }
  *out = p;
  return HTN_OK;
}

int32_t htn_transfer_apply(htn_plan* p, const htn_tensor* A, const htn_tensor* At, const htn_tensor* env_in,
                           htn_tensor* env_out) {
  if (!p || !A || !At || !env_in || !env_out) return HTN_ERR_INVALID;
  if (p->kind != HTN_PLAN_TRANSFER_L && p->kind != HTN_PLAN_TRANSFER_R)
    return p->ctx->fail(HTN_ERR_INVALID, "transfer_apply: not a transfer plan");
  if (env_in->d == env_out->d) return p->ctx->fail(HTN_ERR_INVALID, "transfer_apply: in and out must differ");
  std::lock_guard<std::recursive_mutex> g(p->ctx->mu);
  cudaSetDevice(p->ctx->device);
  const double* slots[4] = {A->d, At->d, env_in->d, env_out->d};
  return p->prog.run(slots);
}

static int32_t check_xy(htn_plan* p, const htn_tensor* x, const htn_tensor* y) {
  if (p->kind != HTN_PLAN_HEFF_AC && p->kind != HTN_PLAN_HEFF_C && p->kind != HTN_PLAN_HEFF_AC2)
    return p->ctx->fail(HTN_ERR_INVALID, "heff_apply: not an effective-Hamiltonian plan");
  if (!htn_same_structure(p->like_in, x) || !htn_same_structure(p->like_in, y))
    return p->ctx->fail(HTN_ERR_SHAPE, "heff_apply: x / y do not have the plan's block structure");
  if (x == y || x->d == y->d) return p->ctx->fail(HTN_ERR_INVALID, "heff_apply: x and y must be distinct tensors");
  return HTN_OK;
}

int32_t htn_heff_run(htn_plan* p, const double* x, double* y, int mask) {
  const double* slots[4] = {x, y, p->bound[2] ? p->bound[2]->d : nullptr, p->bound[3] ? p->bound[3]->d : nullptr};
  const unsigned char* tmaps[4] = {nullptr, nullptr, p->bound[2] ? p->bound[2]->d_tmaps : nullptr,
                                   p->bound[3] ? p->bound[3]->d_tmaps : nullptr};
  return p->prog.run(slots, mask, tmaps);
}

int32_t htn_heff_apply(htn_plan* p, const htn_tensor* x, htn_tensor* y) {
  if (!p || !x || !y) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(p->ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(p->ctx->device);
  return htn_heff_run(p, x->d, y->d, 0xF);
}

int32_t htn_heff_apply_host(htn_plan* p, const double* x_host, double* y_host, int64_t nelem) {
  if (!p || !x_host || !y_host) return HTN_ERR_INVALID;
  htn_ctx* ctx = p->ctx;
  int32_t rc;
  if (!p->hx) {
    if ((rc = htn_tensor_create_like(p->like_in, &p->hx))) return rc;
    if ((rc = htn_tensor_create_like(p->like_in, &p->hy))) return rc;
  }
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  cudaSetDevice(ctx->device);
  if ((rc = htn_upload_locked(p->hx, x_host, nelem))) return rc;
  if ((rc = htn_heff_run(p, p->hx->d, p->hy->d, 0xF))) return rc;
  return htn_download_locked(p->hy, y_host, nelem);
}

int32_t htn_heff_time(htn_plan* p, const htn_tensor* x, htn_tensor* y, int32_t reps, float* ms_total) {
  if (!p || !x || !y || !ms_total || reps <= 0) return HTN_ERR_INVALID;
  htn_ctx* ctx = p->ctx;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaStreamSynchronize(ctx->stream);
  cudaEventRecord(e0, ctx->stream);
  for (int r = 0; r < reps && rc == HTN_OK; ++r) rc = htn_heff_run(p, x->d, y->d, 0xF);
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
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  int32_t rc = check_xy(p, x, y);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int masks[4] = {15, TAG_L, TAG_W | TAG_Y, TAG_R};
  for (int k = 0; k < 4; ++k) {
    cudaStreamSynchronize(ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    for (int r = 0; r < reps; ++r)
      if ((rc = htn_heff_run(p, x->d, y->d, masks[k]))) break;
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
  if (rc == HTN_OK) rc = htn_heff_run(p, x->d, y->d, 0xF);  // leave y = H x behind
  return rc;
}

}  // extern "C"
