// Contraction-program builder and executor (see htn_program.hpp).
#include "htn_program.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <queue>
#include <tuple>

namespace htn {

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
  if (M <= 0 || N <= 0) return best;
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

inline void enc(const Opnd& o, long long& off, int& base) {
  // workspace references stay offsets until finalize() turns them into absolute pointers
  off = o.off;
  base = o.slot == SLOT_WS ? -1 : o.slot + 1;
}

double padded_tile_flops(int mn, int nn, int K) {
  return 2.0 * ((mn + 7) / 8 * 8) * ((nn + 7) / 8 * 8) * ((K + 3) / 4 * 4.0);
}

template <class T>
int32_t to_device(htn_ctx* ctx, const std::vector<T>& v, T** out) {
  *out = nullptr;
  if (v.empty()) return HTN_OK;
  cudaError_t e = cudaMalloc(out, v.size() * sizeof(T));
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_OOM, std::string("cudaMalloc(program table): ") + cudaGetErrorString(e));
  e = h2d_on_stream(*out, v.data(), v.size() * sizeof(T), ctx->stream);
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_CUDA, std::string("cudaMemcpy(program table): ") + cudaGetErrorString(e));
  return HTN_OK;
}

}  // namespace

int64_t Program::ws_alloc(int64_t elems) {
  int64_t off = ws_elems;
  ws_elems = align_up(ws_elems + std::max<int64_t>(elems, 0), 16);
  return off;
}

typedef std::map<std::tuple<int, int64_t, int>, int> GroupMap;
static int group_id(GroupMap& g, const Opnd& o, int tile_off) {
  auto key = std::make_tuple(o.slot, o.off, tile_off);
  auto it = g.find(key);
  if (it != g.end()) return it->second;
  const int id = (int)g.size();
  g[key] = id;
  return id;
}

static void emit_seg(Stage& st, const GemmSegH& s, const TileSpec& ts) {
  GemmSeg sg{};
  enc(s.A, sg.a_off, sg.a_base);
  sg.a_off += (int64_t)ts.mo * s.lda;
  sg.lda = s.lda;
  enc(s.B, sg.b_off, sg.b_base);
  sg.b_off += ts.no;
  sg.ldb = s.ldb;
  sg.K = s.K;
  st.segs.push_back(sg);
}

void Program::add_gemm(std::vector<GemmTaskH>& tasks, int tag) {
  Stage st;
  st.kind = 0;
  st.tag = tag;
  GroupMap groups;
  for (const GemmTaskH& t : tasks) {
    for (const GemmSegH& s : t.segs) {
      flops += 2.0 * t.M * t.N * s.K;
      flops_tag[tag & 15] += 2.0 * t.M * t.N * s.K;
    }
    for (const TileSpec& ts : tile_block(t.M, t.N)) {
      GemmItem it{};
      enc(t.C, it.c_off, it.c_base);
      it.c_off += (int64_t)ts.mo * t.ldc + ts.no;
      it.ldc = t.ldc;
      it.mt = ts.mn;
      it.nt = ts.nn;
      it.layout = ts.layout;
      it.beta = 0;
      it.seg_begin = (int)st.segs.size();
      it.nchunks = 0;
      it.group = -1;
      for (const GemmSegH& s : t.segs) {
        if (s.K <= 0) continue;
        if (it.group < 0) it.group = group_id(groups, s.A, ts.mo);   // stage L: the environment block is the big shared operand
        emit_seg(st, s, ts);
        it.nchunks += (s.K + GEMM_BK - 1) / GEMM_BK;
        padded_flops += padded_tile_flops(ts.mn, ts.nn, s.K);
      }
      it.seg_end = (int)st.segs.size();
      st.items.push_back(it);
    }
  }
  auto cost = [](const GemmItem& it) { return (double)it.mt * it.nt * it.nchunks; };
  std::stable_sort(st.items.begin(), st.items.end(), [&](const GemmItem& a, const GemmItem& b) { return cost(a) > cost(b); });
  n_gemm_tiles_tag[tag & 15] += (int)st.items.size();
  if (!st.items.empty()) stages.push_back(std::move(st));
}

void Program::add_gemm_reduce(std::vector<GemmTaskH>& tasks, std::vector<std::vector<MixSrcH>>& extra, int tag_gemm,
                              int tag_mix, bool adaptive_split) {
  // split-K: a reduce task has few tiles but a K loop over every (level, sector) pair; cut the
  // segment list into nsplit parts of ~SPLIT_CHUNKS chunks, each writing its own partial copy of
  // the output block (summed in fixed order by the mix => deterministic, no atomics)
  static int split_k = 0;
  if (split_k == 0) {
    const char* e = getenv("HTN_SPLIT_K");  // tuning knob: K extent of one split-K part
    split_k = e ? std::max(GEMM_BK, atoi(e)) : 640;
  }
  int SPLIT_CHUNKS = split_k / GEMM_BK;
  const int SPLIT_MAX = 32;
  if (adaptive_split && !getenv("HTN_SPLIT_K")) {
    // small reduce stages (a chi = 10 MPO at D_red = 1024: 282 tile-parts of K = 640 for 444 persistent CTAs) are cut finer,
    // down to K = 128 per part, until there are ~3 parts per CTA
    long long chunk_tiles = 0;
    for (const GemmTaskH& t : tasks) {
      long long ch = 0;
      for (const GemmSegH& sg : t.segs)
        if (sg.K > 0) ch += (sg.K + GEMM_BK - 1) / GEMM_BK;
      if (t.M > 0 && t.N > 0) chunk_tiles += ch * (long long)tile_block(t.M, t.N).size();
    }
    const long long target_items = 3LL * 148 * 3;
    SPLIT_CHUNKS = (int)std::max<long long>(8, std::min<long long>(SPLIT_CHUNKS, chunk_tiles / target_items));
  }
  Stage st;
  st.kind = 0;
  st.tag = tag_gemm;
  GroupMap groups;
  std::vector<MixTaskH> mixes;
  for (size_t ti = 0; ti < tasks.size(); ++ti) {
    const GemmTaskH& t = tasks[ti];
    MixTaskH mx;
    mx.dst = t.C;
    mx.nelem = t.M * t.ldc;
    if (ti < extra.size()) mx.srcs = extra[ti];
    std::vector<const GemmSegH*> segs;
    int total_chunks = 0;
    for (const GemmSegH& s : t.segs)
      if (s.K > 0) {
        segs.push_back(&s);
        total_chunks += (s.K + GEMM_BK - 1) / GEMM_BK;
        flops += 2.0 * t.M * t.N * s.K;
        flops_tag[tag_gemm & 15] += 2.0 * t.M * t.N * s.K;
      }
    if (!segs.empty() && t.M > 0 && t.N > 0) {
      int nsplit = std::max(1, std::min({SPLIT_MAX, (total_chunks + SPLIT_CHUNKS / 2) / SPLIT_CHUNKS, (int)segs.size()}));
      std::vector<int> cut(nsplit + 1, 0);
      {
        int acc = 0, sidx = 1;
        for (size_t q = 0; q < segs.size(); ++q) {
          acc += (segs[q]->K + GEMM_BK - 1) / GEMM_BK;
          while (sidx < nsplit && acc >= (long long)total_chunks * sidx / nsplit) cut[sidx++] = (int)q + 1;
        }
        for (; sidx <= nsplit; ++sidx) cut[sidx] = (int)segs.size();
        for (int q = 1; q <= nsplit; ++q) cut[q] = std::max(cut[q], cut[q - 1]);
      }
      const int64_t blk = (int64_t)t.M * t.ldc;
      std::vector<int64_t> poff(nsplit, -1);
      for (int sp = 0; sp < nsplit; ++sp)
        if (cut[sp + 1] > cut[sp]) {
          poff[sp] = ws_alloc(blk);
          mx.srcs.push_back(MixSrcH{Opnd{SLOT_WS, poff[sp]}, 1.0});
        }
      for (const TileSpec& ts : tile_block(t.M, t.N))
        for (int sp = 0; sp < nsplit; ++sp) {
          if (poff[sp] < 0) continue;
          GemmItem it{};
          it.c_base = -1;
          it.c_off = poff[sp] + (int64_t)ts.mo * t.ldc + ts.no;
          it.ldc = t.ldc;
          it.mt = ts.mn;
          it.nt = ts.nn;
          it.layout = ts.layout;
          it.beta = 0;
          it.seg_begin = (int)st.segs.size();
          it.nchunks = 0;
          it.group = group_id(groups, segs[cut[sp]]->B, ts.no);   // stage R: parts that start on the same GR block walk the same GR blocks
          for (int q = cut[sp]; q < cut[sp + 1]; ++q) {
            emit_seg(st, *segs[q], ts);
            it.nchunks += (segs[q]->K + GEMM_BK - 1) / GEMM_BK;
            padded_flops += padded_tile_flops(ts.mn, ts.nn, segs[q]->K);
          }
          it.seg_end = (int)st.segs.size();
          st.items.push_back(it);
        }
    }
    if (mx.nelem > 0) mixes.push_back(std::move(mx));
  }
  auto cost = [](const GemmItem& it) { return (double)it.mt * it.nt * it.nchunks; };
  std::stable_sort(st.items.begin(), st.items.end(), [&](const GemmItem& a, const GemmItem& b) { return cost(a) > cost(b); });
  n_gemm_tiles_tag[tag_gemm & 15] += (int)st.items.size();
  if (!st.items.empty()) stages.push_back(std::move(st));
  add_mix(mixes, tag_mix);
}

static void fill_mix_tables(Stage& st, const std::vector<MixTaskH>& tasks);

void Program::add_mix(std::vector<MixTaskH>& tasks, int tag) {
  Stage st;
  st.kind = 1;
  st.tag = tag;
  fill_mix_tables(st, tasks);
  if (!st.mc.empty()) stages.push_back(std::move(st));
}

void Program::add_stack(std::vector<StackJobH>& jobs, std::vector<MixTaskH>& mixes, int nwaves, int tmap_slot, int tag) {
  Stage st;
  st.kind = 2;
  st.tag = tag;
  st.nwaves = std::max(nwaves, 0);
  st.tmap_slot = tmap_slot;
  st.wave_need.assign(std::max(nwaves, 1), 0);
  // table order = ticket order: wave by wave, heavy jobs first
  auto cost = [](const StackJobH& j) {
    return (double)((j.M + 63) / 64) * (((j.K + 15) / 16) * (((j.nt + 7) / 8) + 0.75) + 6.0) + 20.0;
  };
  auto first_wave = [](const StackJobH& j) { return j.wave >= 0 ? j.wave : (j.tw.empty() ? 0 : j.tw.front().first); };
  std::stable_sort(jobs.begin(), jobs.end(), [&](const StackJobH& a, const StackJobH& b) {
    if (first_wave(a) != first_wave(b)) return first_wave(a) < first_wave(b);
    return cost(a) > cost(b);
  });
  for (const StackJobH& j : jobs) {
    StackJob d{};
    enc(j.A, d.a_off, d.a_base);
    enc(j.B, d.b_off, d.b_base);
    enc(j.C, d.c_off, d.c_base);
    d.lda = j.lda;
    d.ldb = j.ldb;
    d.ldc = j.ldc;
    d.K = j.K;
    d.nt = j.nt;
    d.nb = j.nb;
    d.M = j.M;
    d.tmap = j.tmap;
    d.arow = j.arow;
    d.wave = j.wave;
    d.tile0 = -1;
    if (j.wave < 0) {
      d.tile0 = (int)st.tile_waves.size();
      for (auto& pr : j.tw) {
        st.tile_waves.push_back(make_int2(pr.first, pr.second));
        for (int w = pr.first; w <= pr.second; ++w)
          if (w >= 0 && w < (int)st.wave_need.size()) st.wave_need[w] += 4;  // the four warps of the group that computes the tile
      }
    }
    st.sjobs.push_back(d);
    if (j.wave >= 0 && j.wave < (int)st.wave_need.size()) st.wave_need[j.wave] += stack_gemm_cons_warps();  // every consumer warp signals per job
    padded_flops += 2.0 * ((j.M + 7) / 8 * 8) * ((j.nt + 7) / 8 * 8) * ((j.K + 3) / 4 * 4.0);
  }
  // ticket order of the mix: wave by wave; inside a wave the targets with the longest source lists first (a chunk is one
  // warp's serial chain of source pieces: the direct-y targets with ~60 sources must not be the last tickets of a wave)
  static int mix_lpt = -1;
  if (mix_lpt < 0) {
    const char* e = getenv("HTN_MIX_LPT");
    mix_lpt = e ? atoi(e) : 1;
  }
  std::stable_sort(mixes.begin(), mixes.end(), [](const MixTaskH& a, const MixTaskH& b) {
    if (a.wave != b.wave) return a.wave < b.wave;
    return mix_lpt ? a.srcs.size() > b.srcs.size() : false;
  });
  fill_mix_tables(st, mixes);
  {  // mix chunks per wave, behind the job counts
    const size_t nw = st.wave_need.size();
    st.wave_need.resize(2 * nw, 0);
    for (const MixChunk& c : st.mc)
      if (c.pad_ >= 0 && c.pad_ < (int)nw) st.wave_need[nw + c.pad_] += 1;
  }
  n_gemm_tiles_tag[TAG_L] += (int)st.sjobs.size();
  if (getenv("HTN_PLAN_DEBUG")) {
    long long tiles = 0;
    std::vector<int> per_wave(st.wave_need.size() / 2, 0);
    for (const StackJob& j : st.sjobs) {
      tiles += (j.M + 63) / 64;
      if (j.wave >= 0 && j.wave < (int)per_wave.size()) per_wave[j.wave]++;
    }
    fprintf(stderr, "[htn] stacked stage: %zu jobs, %lld tiles, %d waves, %zu mix targets, %zu mix chunks; jobs per wave:", st.sjobs.size(),
            tiles, st.nwaves, st.mt.size(), st.mc.size());
    for (int v : per_wave) fprintf(stderr, " %d", v);
    fprintf(stderr, "\n");
  }
  if (!st.sjobs.empty() || !st.mc.empty()) stages.push_back(std::move(st));
}

static void fill_mix_tables(Stage& st, const std::vector<MixTaskH>& tasks) {
  int last_wave = -1;
  for (const MixTaskH& t : tasks) last_wave = std::max(last_wave, t.wave);
  static int tail_div = -1;
  if (tail_div < 0) {
    const char* e = getenv("HTN_MIX_TAIL_DIV");  // chunks of the last wave are this many times smaller (they run with nothing to hide behind)
    tail_div = e ? std::max(1, atoi(e)) : 1;
  }
  for (const MixTaskH& t : tasks) {
    if (t.nelem <= 0) continue;
    MixTarget mt{};
    enc(t.dst, mt.off, mt.base);
    mt.nelem = t.nelem;
    mt.src_begin = (int)st.ms.size();
    for (const MixSrcH& s : t.srcs) {
      MixSrc ms{};
      enc(s.src, ms.off, ms.base);
      ms.coef = s.coef;
      st.ms.push_back(ms);
    }
    mt.src_end = (int)st.ms.size();
    int ti = (int)st.mt.size();
    st.mt.push_back(mt);
    // ~16k element-sources per CTA, chunk a multiple of 512 elements (256 threads x double2)
    int per = 16384 / std::max<int>(1, (int)t.srcs.size());
    if (st.kind == 2 && last_wave > 0 && t.wave == last_wave) per /= tail_div;
    per = std::max(512, std::min(8192, per / 512 * 512));
    for (int e = 0; e < t.nelem; e += per) st.mc.push_back(MixChunk{ti, e, std::min(per, t.nelem - e), t.wave});
  }
}

// Static schedule of one GEMM stage: the kernel's CTA b walks items b, b + G, b + 2G, ...  The time a
// tile holds its CTA is set by the flex extent (the DMMA sequence every consumer warp issues per
// chunk) and the number of K chunks, not by its area, plus a fixed charge per chunk (ring
// hand-over) and per item (epilogue, table reads).  Longest-processing-time-first over G CTAs with
// that model; lists are padded to equal length with empty items (mt = 0) the kernel skips.
static void balance_items(std::vector<GemmItem>& items, int cap, int tag, int nsm) {
  const int n = (int)items.size();
  const int G = std::max(1, std::min(n, cap));
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("HTN_GEMM_BALANCE");
    mode = e ? atoi(e) : 1;
  }
  if (mode == 0 || n <= G) return;
  auto cost = [](const GemmItem& it) {
    const int flex = ((it.layout ? it.nt : it.mt) + 7) >> 3;
    return 6.0 + (double)it.nchunks * (flex + 0.75);
  };
  std::vector<double> c(n);
  double total = 0.0;
  for (int i = 0; i < n; ++i) total += (c[i] = cost(items[i]));
  // units of scheduling: single items.  Experiment HTN_GEMM_BALANCE=3: the items of one group (same big operand
  // block) stay together on one CTA, back to back, so that the repeat reads of that block hit L2 (groups heavier
  // than a quarter of a CTA's share are cut).  Measured SLOWER (stage L +4 %, stage R +18 %): a CTA then runs
  // look-alike tiles in a row and the three CTAs of an SM fall into the same load/DMMA/store phases.
  std::vector<std::vector<int>> units;
  {
    std::map<int, std::vector<int>> by_group;
    for (int i = 0; i < n; ++i) {
      if (items[i].group < 0 || mode != 3)
        units.push_back({i});
      else
        by_group[items[i].group].push_back(i);
    }
    const double cap_unit = 0.25 * total / G;
    for (auto& kv : by_group) {
      std::vector<int> cur;
      double acc = 0.0;
      for (int i : kv.second) {
        if (!cur.empty() && acc + c[i] > cap_unit) {
          units.push_back(cur);
          cur.clear();
          acc = 0.0;
        }
        cur.push_back(i);
        acc += c[i];
      }
      if (!cur.empty()) units.push_back(cur);
    }
  }
  const int nu = (int)units.size();
  std::vector<double> uc(nu, 0.0);
  for (int u = 0; u < nu; ++u)
    for (int i : units[u]) uc[u] += c[i];
  std::vector<int> order(nu);
  for (int u = 0; u < nu; ++u) order[u] = u;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return uc[a] > uc[b]; });
  typedef std::pair<double, int> Load;  // (load, cta)
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int b = 0; b < G; ++b) heap.push(Load(0.0, b));
  std::vector<std::vector<int>> lists(G);
  for (int u : order) {
    Load l = heap.top();
    heap.pop();
    for (int i : units[u]) lists[l.second].push_back(i);
    heap.push(Load(l.first + uc[u], l.second));
  }
  if (mode == 4) {  // experiment: random order inside every CTA's list (fixed seed)
    uint64_t rs = 0x9E3779B97F4A7C15ull;
    for (auto& l : lists)
      for (size_t k = l.size(); k > 1; --k) {
        rs = rs * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(l[k - 1], l[(size_t)((rs >> 33) % k)]);
      }
  }
  // Order inside a CTA's list (LPT leaves it sorted by falling cost).  The three CTAs of an SM overlap best when
  // long and short tiles alternate: 1 = zigzag (largest, smallest, 2nd largest, ...), 2 = interleave the two
  // halves (largest, median, 2nd largest, ...), 0 = falling cost.  Stage L (many short-K tiles per CTA) gains
  // 6 % from the zigzag; stage R (a handful of long split-K parts per CTA) loses 8 % and keeps falling cost.
  static int order_l = -1, order_r = -1;
  if (order_l < 0) {
    const char* e = getenv("HTN_ORDER_L");
    order_l = e ? atoi(e) : 1;
    e = getenv("HTN_ORDER_R");
    order_r = e ? atoi(e) : 0;
  }
  const int ord = mode == 5 ? 1 : (tag == TAG_L ? order_l : order_r);
  if (ord == 3 || ord == 4) {  // experiments: zigzag over the K length (3) or the flex extent (4) instead of the cost
    for (auto& l : lists)
      std::stable_sort(l.begin(), l.end(), [&](int a, int b) {
        if (ord == 3) return items[a].nchunks > items[b].nchunks;
        const int fa = items[a].layout ? items[a].nt : items[a].mt, fb = items[b].layout ? items[b].nt : items[b].mt;
        return fa > fb;
      });
  }
  if (ord >= 1 && ord <= 4) {
    for (auto& l : lists) {
      std::vector<int> z;
      if (ord != 2) {
        size_t lo = 0, hi = l.size();
        while (lo < hi) {
          z.push_back(l[lo++]);
          if (lo < hi) z.push_back(l[--hi]);
        }
      } else {
        const size_t h = (l.size() + 1) / 2;
        for (size_t k = 0; k < h; ++k) {
          z.push_back(l[k]);
          if (k + h < l.size()) z.push_back(l[k + h]);
        }
      }
      l.swap(z);
    }
  }
  // De-phasing of the CTAs that share an SM (blockIdx = sm, sm + nsm, sm + 2 nsm): 0 = forward / backward / from
  // the middle, 1 = cyclic shifts by thirds of the list
  static int wave_mode = -1;
  if (wave_mode < 0) {
    const char* e = getenv("HTN_WAVE");
    wave_mode = e ? atoi(e) : 0;
  }
  for (int b = 0; b < G; ++b) {
    std::vector<int>& l = lists[b];
    const int w = (b / std::max(1, nsm)) % 3, nl = (int)l.size();
    if (w == 0 || nl < 2) continue;
    if (wave_mode == 0) {
      if (w == 1)
        std::reverse(l.begin(), l.end());
      else
        std::rotate(l.begin(), l.begin() + nl / 2, l.end());
    } else if (wave_mode == 1) {
      std::rotate(l.begin(), l.begin() + (w * nl) / 3, l.end());
    }
  }
  size_t nmax = 0;
  for (const auto& l : lists) nmax = std::max(nmax, l.size());
  if (getenv("HTN_PLAN_DEBUG")) {
    double tot = 0, mx = 0, mx0 = 0;
    std::vector<double> rr(G, 0.0);
    for (int i = 0; i < n; ++i) rr[i % G] += c[i], tot += c[i];
    fprintf(stderr, "[htn] gemm stage: %d scheduling units\n", nu);  // the unbalanced deal (items arrive sorted by area)
    for (int b = 0; b < G; ++b) mx0 = std::max(mx0, rr[b]);
    while (!heap.empty()) mx = std::max(mx, heap.top().first), heap.pop();
    fprintf(stderr, "[htn] gemm stage: %d items on %d CTAs, model load max/mean %.3f (round-robin deal %.3f), list length %zu\n",
            n, G, mx / (tot / G), mx0 / (tot / G), nmax);
  }
  std::vector<GemmItem> out(nmax * (size_t)G, GemmItem{});
  for (int b = 0; b < G; ++b)
    for (size_t k = 0; k < lists[b].size(); ++k) out[k * G + b] = items[lists[b][k]];
  items.swap(out);
}

int32_t Program::finalize(htn_ctx* c, int nslots_) {
  ctx = c;
  nslots = nslots_;
  cudaSetDevice(ctx->device);
  ws_elems = std::max<int64_t>(ws_elems, 16);
  if (cudaMalloc(&ws, ws_elems * sizeof(double)) != cudaSuccess) {
    ws = nullptr;
    return ctx->fail(HTN_ERR_OOM, "program workspace allocation failed");
  }
  cudaMemsetAsync(ws, 0, ws_elems * sizeof(double), ctx->stream);  // ordered before the first run on this stream
  auto fix = [&](long long& off, int& base) {
    if (base == -1) {
      off = reinterpret_cast<long long>(ws + off);
      base = 0;
    }
  };
  int per_sm = gemm_max_ctas_per_sm();
  if (const char* e = getenv("HTN_GEMM_CTAS")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // occupancy experiments
  const int cap = ctx->sm_count * per_sm;
  int32_t rc;
  for (Stage& st : stages) {
    if (st.kind == 0) {
      for (GemmSeg& sg : st.segs) {
        fix(sg.a_off, sg.a_base);
        fix(sg.b_off, sg.b_base);
      }
      for (GemmItem& it : st.items) fix(it.c_off, it.c_base);
      if (st.segs.empty()) st.segs.push_back(GemmSeg{});  // keep the table pointer valid
      balance_items(st.items, cap, st.tag, ctx->sm_count);
      if ((rc = to_device(ctx, st.items, &st.d_items)) || (rc = to_device(ctx, st.segs, &st.d_segs))) return rc;
      st.n = (int)st.items.size();
      st.grid = std::max(1, std::min(st.n, cap));
    } else {
      for (MixTarget& t : st.mt) fix(t.off, t.base);
      for (MixSrc& s : st.ms) fix(s.off, s.base);
      if (st.ms.empty()) st.ms.push_back(MixSrc{});
      if (st.kind == 2) {
        if (st.mt.empty()) st.mt.push_back(MixTarget{});
        if (st.mc.empty()) st.mc.push_back(MixChunk{});
        st.n = (int)st.mc.size();
        if (st.mc.size() == 1 && st.mc[0].nelem == 0) st.n = 0;
      }
      if ((rc = to_device(ctx, st.mt, &st.d_mt)) || (rc = to_device(ctx, st.ms, &st.d_ms)) ||
          (rc = to_device(ctx, st.mc, &st.d_mc)))
        return rc;
      if (st.kind == 2) {
        std::vector<MixChunkX> mcx;
        for (const MixChunk& c : st.mc) {
          const MixTarget& t = st.mt[c.target];
          mcx.push_back(MixChunkX{t.off, t.base, t.src_begin, t.src_end - t.src_begin, c.elem0, c.nelem, c.pad_});
        }
        if ((rc = to_device(ctx, mcx, &st.d_mcx))) return rc;
      }
      if (st.kind == 1) st.n = (int)st.mc.size();
      if (st.kind == 2) {
        for (StackJob& j : st.sjobs) {
          fix(j.a_off, j.a_base);
          fix(j.b_off, j.b_base);
          fix(j.c_off, j.c_base);
        }
        st.n_sjobs = (int)st.sjobs.size();
        if (st.sjobs.empty()) st.sjobs.push_back(StackJob{});
        if (st.tile_waves.empty()) st.tile_waves.push_back(make_int2(0, -1));
        if ((rc = to_device(ctx, st.sjobs, &st.d_sjobs)) || (rc = to_device(ctx, st.wave_need, &st.d_wave_need)) ||
            (rc = to_device(ctx, st.tile_waves, &st.d_tile_waves)))
          return rc;
        const size_t nctr = 4 + st.wave_need.size() + 8;
        if (cudaMalloc(&st.d_ctr, nctr * sizeof(unsigned long long)) != cudaSuccess)
          return ctx->fail(HTN_ERR_OOM, "program counters allocation failed");
        cudaMemsetAsync(st.d_ctr, 0, nctr * sizeof(unsigned long long), ctx->stream);
        st.grid = ctx->sm_count * stack_gemm_ctas_per_sm();
        std::vector<StackJob>().swap(st.sjobs);
      }
    }
    // host tables are no longer needed
    std::vector<GemmItem>().swap(st.items);
    std::vector<GemmSeg>().swap(st.segs);
    std::vector<MixTarget>().swap(st.mt);
    std::vector<MixSrc>().swap(st.ms);
    std::vector<MixChunk>().swap(st.mc);
  }
  finalized = true;
  return HTN_OK;
}

int32_t Program::run(const double* const* slots, int mask, const unsigned char* const* slot_tmaps) const {
  Bases bs;
  for (int i = 0; i < MAX_SLOTS; ++i) bs.p[i] = i < nslots ? slots[i] : nullptr;
  static int stack_dbg = -1, mix_lag = 0;
  if (stack_dbg < 0) {
    const char* e = getenv("HTN_STACK_DEBUG");  // timing experiments only (results are wrong when bits 1, 2, 4 are set)
    stack_dbg = e ? atoi(e) : 0;
    e = getenv("HTN_MIX_LAG");  // > 0: back-pressure on the stacked jobs (measured slower: profiles/r2_stack_mix_analysis.md)
    if (e) mix_lag = atoi(e);
  }
  for (const Stage& st : stages) {
    if (!(st.tag & mask)) continue;
    if (st.kind == 0)
      launch_gemm(st.d_items, st.d_segs, st.n, bs, st.grid, ctx->stream);
    else if (st.kind == 1)
      launch_mix(st.d_mt, st.d_ms, st.d_mc, st.n, bs, ctx->stream);
    else {
      StackArgs a{};
      a.jobs = st.d_sjobs;
      a.njobs = st.n_sjobs;
      a.tmaps = (st.tmap_slot >= 0 && slot_tmaps) ? slot_tmaps[st.tmap_slot] : nullptr;
      if (st.tmap_slot >= 0 && !a.tmaps) return ctx->fail(HTN_ERR_INVALID, "program: the stacked stage needs the tensor maps of its environment operand");
      a.mt = st.d_mt;
      a.ms = st.d_ms;
      a.mc = st.d_mc;
      a.mcx = st.d_mcx;
      a.nmix = (stack_dbg & 16) ? 0 : st.n;  // experiment 16: the mix as a separate launch
      a.wave_need = st.d_wave_need;
      a.tile_waves = st.d_tile_waves;
      a.nwaves = std::max(st.nwaves, 1);
      a.mix_lag = mix_lag;
      a.ctr = st.d_ctr;
      a.epoch = ++st.epoch;
      a.dbg = stack_dbg;
      static unsigned long long* d_ts = nullptr;
      static int ts_cap = 0;
      if (stack_dbg & 32) {
        cudaMemsetAsync(st.d_ctr + 4 + 2 * a.nwaves, 0, 8 * sizeof(unsigned long long), ctx->stream);
        if (ts_cap < st.n) {
          cudaFree(d_ts);
          cudaMalloc(&d_ts, ((size_t)2 * st.n + 256) * sizeof(unsigned long long));
          ts_cap = st.n;
        }
        cudaMemsetAsync(d_ts, 0, ((size_t)2 * st.n + 256) * sizeof(unsigned long long), ctx->stream);
        a.dbg_ts = d_ts;
      }
      launch_stack_gemm(a, bs, st.grid, ctx->stream);
      if ((stack_dbg & 16) && st.n > 0) launch_mix(st.d_mt, st.d_ms, st.d_mc, st.n, bs, ctx->stream);
      if (stack_dbg & 32) {  // timeline probe: when did the DMMA warps and the mixers finish?
        unsigned long long t[8];
        cudaMemcpyAsync(t, st.d_ctr + 4 + 2 * a.nwaves, sizeof(t), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        static int printed = 0;
        if (printed == 3 && d_ts) {  // chunk-level view of one launch: what runs after the DMMA warps are done?
          std::vector<unsigned long long> ts((size_t)2 * st.n + 256);
          cudaMemcpy(ts.data(), d_ts, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
          const unsigned long long t_start = ~t[0], t_jobs = t[1];
          int n_after = 0, n_started_after = 0;
          double sum_dur_after = 0, max_dur = 0, sum_dur_before = 0;
          int n_before = 0;
          std::vector<int> hist(12, 0);
          for (int c = 0; c < st.n; ++c) {
            const unsigned long long b = ts[2 * c], e = ts[2 * c + 1];
            if (!b || !e) continue;
            const double dur = (e - b) * 1e-3;
            if (e > t_jobs) {
              ++n_after;
              sum_dur_after += dur;
              max_dur = std::max(max_dur, dur);
              if (b > t_jobs) ++n_started_after;
              const int bin = std::min(11, (int)((e - t_jobs) * 1e-3 / 10.0));
              hist[bin]++;
            } else {
              ++n_before;
              sum_dur_before += dur;
            }
          }
          fprintf(stderr, "[htn] mix chunks: %d ended before the jobs were done (mean %.1f us each), %d after (%d of them also started after; mean %.1f us, max %.1f us); "
                          "ends per 10 us after jobs-done:", n_before, sum_dur_before / std::max(n_before, 1), n_after, n_started_after,
                  sum_dur_after / std::max(n_after, 1), max_dur);
          for (int v : hist) fprintf(stderr, " %d", v);
          fprintf(stderr, "\n[htn] waves complete at (us):");
          for (int w = 0; w < std::min(a.nwaves, 256); ++w) fprintf(stderr, " %.0f", ts[(size_t)2 * st.n + w] ? (ts[(size_t)2 * st.n + w] - t_start) * 1e-3 : -1.0);
          fprintf(stderr, "\n");
        }
        if (printed++ < 6)
          fprintf(stderr,
                  "[htn] stack timeline: jobs done at %.1f us (%llu of %d mix chunks drawn by then), mix done at %.1f us after the first warp "
                  "started; mixer warps: %llu chunks, per chunk %.0f clk ticket+record, %.0f clk wave wait, %.0f clk data\n",
                  (t[1] - (~t[0])) * 1e-3, t[3], st.n, (t[2] - (~t[0])) * 1e-3, t[7], (double)t[4] / std::max<unsigned long long>(t[7], 1),
                  (double)t[5] / std::max<unsigned long long>(t[7], 1), (double)t[6] / std::max<unsigned long long>(t[7], 1));
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return ctx->fail(HTN_ERR_CUDA, std::string("program launch: ") + cudaGetErrorString(e));
  return HTN_OK;
}

int Program::launches(int mask) const {
  int n = 0;
  for (const Stage& st : stages)
    if (st.tag & mask) ++n;
  return n;
}

void Program::destroy() {
  for (Stage& st : stages) {
    cudaFree(st.d_items);
    cudaFree(st.d_segs);
    cudaFree(st.d_mt);
    cudaFree(st.d_ms);
    cudaFree(st.d_mc);
    cudaFree(st.d_mcx);
    cudaFree(st.d_sjobs);
    cudaFree(st.d_wave_need);
    cudaFree(st.d_tile_waves);
    cudaFree(st.d_ctr);
  }
  stages.clear();
  cudaFree(ws);
  ws = nullptr;
}

}  // namespace htn
