// C-ABI entry points: context, spaces, tensors, MPO, vector algebra (see include/htn.h).
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <new>
#include <numeric>

#include "htn_linalg.hpp"

using namespace htn;

#define HTN_TRY try {
#define HTN_CATCH(ctxptr)                                                                      \
  }                                                                                            \
  catch (const std::bad_alloc&) {                                                              \
    if (ctxptr) (ctxptr)->err = "host allocation failed";                                      \
    return HTN_ERR_OOM;                                                                        \
  }                                                                                            \
  catch (const std::exception& e) {                                                            \
    if (ctxptr) (ctxptr)->err = e.what();                                                      \
    return HTN_ERR_INVALID;                                                                    \
  }                                                                                            \
  catch (...) {                                                                                \
    if (ctxptr) (ctxptr)->err = "unknown error";                                               \
    return HTN_ERR_INVALID;                                                                    \
  }

static thread_local std::string g_noctx_err = "";

static int32_t cuda_fail(htn_ctx* ctx, cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  if (ctx) ctx->err = m;
  return e == cudaErrorMemoryAllocation ? HTN_ERR_OOM : HTN_ERR_CUDA;
}
#define CU(ctx, call)                                          \
  do {                                                         \
    cudaError_t e_ = (call);                                   \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
  } while (0)

extern "C" {

int32_t htn_version(void) { return 100; }

int32_t htn_ctx_create(int32_t device, htn_ctx** out) {
  if (!out) return HTN_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_noctx_err = "no CUDA device available (libhtn has no CPU fallback)";
    return HTN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) {
    g_noctx_err = "device index out of range";
    return HTN_ERR_INVALID;
  }
  htn_ctx* c = new (std::nothrow) htn_ctx();
  if (!c) return HTN_ERR_OOM;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_noctx_err = "cudaSetDevice/cudaStreamCreate failed";
    delete c;
    return HTN_ERR_CUDA;
  }
  cudaDeviceProp prop;
  const bool have_prop = cudaGetDeviceProperties(&prop, device) == cudaSuccess;
  if (have_prop) c->sm_count = prop.multiProcessorCount;
  // the library holds one sm_100a cubin and no PTX: any other device would fail later with "no kernel image"
  if (!have_prop || prop.major != 10 || prop.minor != 0) {
    g_noctx_err = "libhtn is built for sm_100a (B200) only";
    cudaStreamDestroy(c->stream);
    delete c;
    return HTN_ERR_NO_DEVICE;
  }
  if (cudaMallocHost(&c->red_host, 64 * sizeof(double)) != cudaSuccess || cudaMalloc(&c->kry_scal, 8192 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&c->kry_scal_host, 8192 * sizeof(double)) != cudaSuccess || cudaMalloc(&c->d_status, sizeof(int)) != cudaSuccess) {
    g_noctx_err = "context allocation failed";
    htn_ctx_destroy(c);
    return HTN_ERR_OOM;
  }
  cudaMemsetAsync(c->d_status, 0, sizeof(int), c->stream);
  *out = c;
  return HTN_OK;
}

int32_t htn_ctx_destroy(htn_ctx* ctx) {
  if (!ctx) return HTN_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  htn_drop_krylov_graphs(ctx);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->red) cudaFree(ctx->red);
  if (ctx->red_host) cudaFreeHost(ctx->red_host);
  if (ctx->kry_V) cudaFree(ctx->kry_V);
  if (ctx->kry_scal) cudaFree(ctx->kry_scal);
  if (ctx->kry_scal_host) cudaFreeHost(ctx->kry_scal_host);
  if (ctx->kry_partial) cudaFree(ctx->kry_partial);
  if (ctx->d_status) cudaFree(ctx->d_status);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return HTN_OK;
}

const char* htn_last_error_string(htn_ctx* ctx) { return ctx ? ctx->err.c_str() : g_noctx_err.c_str(); }

int32_t htn_ctx_stream(htn_ctx* ctx, void** stream) {
  if (!ctx || !stream) return HTN_ERR_INVALID;
  *stream = reinterpret_cast<void*>(ctx->stream);
  return HTN_OK;
}

int32_t htn_ctx_synchronize(htn_ctx* ctx) {
  if (!ctx) return HTN_ERR_INVALID;
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return HTN_OK;
}

// ---- spaces ---------------------------------------------------------------------------
int32_t htn_space_create(htn_ctx* ctx, int32_t sym, int32_t nsec, const int32_t* labels, const int32_t* mult,
                         htn_space** out) {
  if (!ctx || !out || nsec < 0 || (nsec > 0 && (!labels || !mult))) return HTN_ERR_INVALID;
  if (sym != HTN_SYM_SU2U1 && sym != HTN_SYM_U1U1) return ctx->fail(HTN_ERR_INVALID, "unknown symmetry kind");
  HTN_TRY
  std::vector<std::pair<Sector, int32_t>> v;
  for (int i = 0; i < nsec; ++i) {
    Sector s{labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]};
    if ((s.p != 0 && s.p != 1) || (sym == HTN_SYM_SU2U1 && s.q < 0))
      return ctx->fail(HTN_ERR_INVALID, "invalid sector label");
    if (mult[i] < 0) return ctx->fail(HTN_ERR_INVALID, "negative multiplicity");
    if (mult[i] == 0) continue;
    for (auto& pr : v)
      if (pr.first == s) return ctx->fail(HTN_ERR_INVALID, "duplicate sector in space");
    v.push_back({s, mult[i]});
  }
  std::sort(v.begin(), v.end(), [sym](auto& a, auto& b) { return canonical_less(sym, a.first, b.first); });
  htn_space* sp = new htn_space();
  sp->ctx = ctx;
  sp->sym = sym;
  for (auto& pr : v) {
    sp->sec.push_back(pr.first);
    sp->mult.push_back(pr.second);
  }
  *out = sp;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_space_destroy(htn_space* s) {
  delete s;
  return HTN_OK;
}

int32_t htn_space_info(const htn_space* s, int32_t* nsec, int32_t* labels, int32_t* mult) {
  if (!s) return HTN_ERR_INVALID;
  if (nsec) *nsec = (int32_t)s->sec.size();
  for (size_t i = 0; i < s->sec.size(); ++i) {
    if (labels) {
      labels[3 * i] = s->sec[i].p;
      labels[3 * i + 1] = s->sec[i].q;
      labels[3 * i + 2] = s->sec[i].n;
    }
    if (mult) mult[i] = s->mult[i];
  }
  return HTN_OK;
}

int32_t htn_legs_create(htn_ctx* ctx, int32_t sym, int32_t n, const int32_t* labels, htn_legs** out) {
  // host-only object: ctx may be NULL (used by the CPU-side tests of the MPO projection)
  if (!out || n < 0 || (n > 0 && !labels)) return HTN_ERR_INVALID;
  if (sym != HTN_SYM_SU2U1 && sym != HTN_SYM_U1U1) return HTN_ERR_INVALID;
  HTN_TRY
  htn_legs* l = new htn_legs();
  l->ctx = ctx;
  l->sym = sym;
  for (int i = 0; i < n; ++i) l->sec.push_back(Sector{labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]});
  *out = l;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_legs_destroy(htn_legs* l) {
  delete l;
  return HTN_OK;
}

// ---- tensors --------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (libhtn links no driver library)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int32_t finalize_tensor(htn_tensor* t) {
  htn_ctx* ctx = t->ctx;
  // Blocks of one coupled sector (MPS: same r; transposed MPS: same l) are laid out back to back
  // so that they form ONE row-major panel -- the matrix TensorKit stores per coupled sector and
  // the unit the QR / LQ gauge kernels work on in place.  Panels start 128-byte aligned.
  int64_t off = 0, hoff = 0;
  const bool env = t->kind == HTN_T_ENVL || t->kind == HTN_T_ENVR;
  if (env) {
    // stacked panels (htn_internal.hpp): device order (lab[2]; identity level last; lab[1]; level), table order kept
    const int nv = (int)t->s0.sec.size();
    std::vector<int> order(t->blocks.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    const int idl = t->identity_level;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      const Block &a = t->blocks[x], &b = t->blocks[y];
      const int ia = a.lab[0] == idl, ib = b.lab[0] == idl;
      return std::make_tuple(a.lab[2], ia, a.lab[1], a.lab[0]) < std::make_tuple(b.lab[2], ib, b.lab[1], b.lab[0]);
    });
    t->panels.assign(nv, htn_tensor::Panel{0, 0, 0, 0, 0});
    t->block_prow.assign(t->blocks.size(), 0);
    int prev = -1;
    for (int i : order) {
      Block& b = t->blocks[i];
      b.ld = even_up(b.cols);
      const int g = b.lab[2];
      if (g != prev) {
        off = align_up(off, 16);
        t->panels[g].off = off;
        t->panels[g].cols = b.cols;
        t->panels[g].ld = b.ld;
        prev = g;
      }
      b.off = off;
      t->block_prow[i] = t->panels[g].rows;
      t->panels[g].rows += b.rows;
      if (b.lab[0] != idl) t->panels[g].rows_active += b.rows;
      off += (int64_t)b.rows * b.ld;
    }
  }
  int prev_group = -1;
  for (size_t i = 0; i < t->blocks.size(); ++i) {
    Block& b = t->blocks[i];
    if (!env) {
      b.ld = even_up(b.cols);
      const int group = t->kind == HTN_T_MPS ? b.lab[2] : (t->kind == HTN_T_MPST ? b.lab[0] : -2 - (int)i);
      if (group != prev_group) off = align_up(off, 16);
      prev_group = group;
      b.off = off;
      off += (int64_t)b.rows * b.ld;
    }
    b.hoff = hoff;
    hoff += (int64_t)b.rows * b.cols;
    if (t->kind == HTN_T_MPS2)
      t->index5[std::array<int, 5>{b.lab[0], b.lab[1], b.lab[2], b.lab[3], b.lab[4]}] = (int)i;
    else
      t->index[std::make_tuple(b.lab[0], b.lab[1], b.lab[2])] = (int)i;
  }
  off = align_up(off, 16);
  t->dsize = std::max<int64_t>(off, 16);
  t->hsize = hoff;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMalloc(&t->d, t->dsize * sizeof(double)));
  CU(ctx, cudaMemsetAsync(t->d, 0, t->dsize * sizeof(double), ctx->stream));
  if (env && !t->panels.empty() && tensor_map_encoder()) {
    // one 2-D tensor map per panel: dims (cols, rows), row pitch ld * 8 B, box 16 x 64, 128-byte swizzle; columns
    // beyond `cols` (K tails, the pad column) and rows beyond the panel read as zeros
    std::vector<CUtensorMap> maps(t->panels.size());
    bool ok = true;
    for (size_t g = 0; g < t->panels.size() && ok; ++g) {
      const htn_tensor::Panel& pn = t->panels[g];
      memset(&maps[g], 0, sizeof(CUtensorMap));
      if (pn.rows == 0 || pn.cols == 0) continue;
      cuuint64_t dims[2] = {(cuuint64_t)pn.cols, (cuuint64_t)pn.rows};
      cuuint64_t strides[1] = {(cuuint64_t)pn.ld * 8};
      cuuint32_t box[2] = {16, 64};
      cuuint32_t es[2] = {1, 1};
      ok = tensor_map_encoder()(&maps[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, t->d + pn.off, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (ok) {
      CU(ctx, cudaMalloc(&t->d_tmaps, maps.size() * sizeof(CUtensorMap)));
      CU(ctx, cudaMemcpyAsync(t->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream));
      CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  // device block table + row-chunk table
  std::vector<DevBlock> db(t->blocks.size());
  std::vector<int> chunks;
  for (size_t i = 0; i < t->blocks.size(); ++i) {
    const Block& b = t->blocks[i];
    db[i] = DevBlock{b.off, b.hoff, b.rows, b.cols, b.ld, b.weight};
    int rows_per = std::max(1, 1024 / std::max(1, b.ld));  // ~1k elements per CTA: enough CTAs to fill 148 SMs at 3 MB vectors
    for (int r0 = 0; r0 < b.rows; r0 += rows_per) {
      chunks.push_back((int)i);
      chunks.push_back(r0);
      chunks.push_back(std::min(rows_per, b.rows - r0));
    }
  }
  t->nchunks = (int)chunks.size() / 3;
  if (!db.empty()) {
    CU(ctx, cudaMalloc(&t->dblocks, db.size() * sizeof(DevBlock)));
    CU(ctx, cudaMemcpyAsync(t->dblocks, db.data(), db.size() * sizeof(DevBlock), cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMalloc(&t->dchunks, chunks.size() * sizeof(int)));
    CU(ctx, cudaMemcpyAsync(t->dchunks, chunks.data(), chunks.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
  return HTN_OK;
}

static htn_tensor* new_tensor(htn_ctx* ctx, int kind, int sym) {
  static std::atomic<uint64_t> next_uid{0};
  htn_tensor* t = new htn_tensor();
  t->ctx = ctx;
  t->uid = ++next_uid;
  t->kind = kind;
  t->sym = sym;
  t->s0.ctx = t->s1.ctx = ctx;
  t->s0.sym = t->s1.sym = sym;
  t->legs.ctx = ctx;
  t->legs.sym = sym;
  return t;
}

int32_t htn_tensor_create_mps(htn_ctx* ctx, const htn_space* Vl, const htn_legs* P, const htn_space* Vr,
                              htn_tensor** out) {
  if (!ctx || !Vl || !P || !Vr || !out) return HTN_ERR_INVALID;
  if (Vl->sym != P->sym || Vr->sym != P->sym) return ctx->fail(HTN_ERR_INVALID, "symmetry kinds differ");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, HTN_T_MPS, P->sym);
  t->s0 = *Vl;
  t->s1 = *Vr;
  t->legs = *P;
  // canonical block order: coupled sector r, then s, then l (oracle/tensors.py:mps_keys)
  for (int r = 0; r < (int)Vr->sec.size(); ++r)
    for (int s = 0; s < (int)P->sec.size(); ++s)
      for (int l = 0; l < (int)Vl->sec.size(); ++l)
        if (allowed(t->sym, Vl->sec[l], P->sec[s], Vr->sec[r])) {
          Block b{};
          b.lab[0] = l;
          b.lab[1] = s;
          b.lab[2] = r;
          b.rows = Vl->mult[l];
          b.cols = Vr->mult[r];
          b.weight = sdim(t->sym, Vr->sec[r]);
          t->blocks.push_back(b);
        }
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_tensor_create_bond(htn_ctx* ctx, const htn_space* V, htn_tensor** out) {
  if (!ctx || !V || !out) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, HTN_T_BOND, V->sym);
  t->s0 = *V;
  t->s1 = *V;
  for (int c = 0; c < (int)V->sec.size(); ++c) {
    Block b{};
    b.lab[0] = b.lab[1] = b.lab[2] = c;
    b.rows = b.cols = V->mult[c];
    b.weight = sdim(t->sym, V->sec[c]);
    t->blocks.push_back(b);
  }
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_tensor_create_env(htn_ctx* ctx, int32_t side, const htn_space* V, const htn_legs* M,
                              int32_t identity_level, htn_tensor** out) {
  if (!ctx || !V || !M || !out) return HTN_ERR_INVALID;
  if (side != HTN_SIDE_LEFT && side != HTN_SIDE_RIGHT) return ctx->fail(HTN_ERR_INVALID, "bad side");
  if (V->sym != M->sym) return ctx->fail(HTN_ERR_INVALID, "symmetry kinds differ");
  if (identity_level >= (int)M->sec.size()) return ctx->fail(HTN_ERR_INVALID, "identity_level out of range");
  if (identity_level >= 0) {
    Sector s = M->sec[identity_level];
    if (s.p != 0 || s.q != 0 || s.n != 0) return ctx->fail(HTN_ERR_INVALID, "identity level must carry the trivial sector");
  }
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, side == HTN_SIDE_LEFT ? HTN_T_ENVL : HTN_T_ENVR, V->sym);
  t->s0 = *V;
  t->s1 = *V;
  t->legs = *M;
  t->identity_level = identity_level;
  const int nv = (int)V->sec.size();
  // canonical order: level, then first bond index, then second (oracle/tensors.py:envl_keys/envr_keys)
  for (int a = 0; a < (int)M->sec.size(); ++a)
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) {
        // left : (a, l'=i, l=j), l' in a(x)l ; right: (b, r=i, r'=j), r' in b(x)r
        bool ok = side == HTN_SIDE_LEFT ? allowed(t->sym, M->sec[a], V->sec[j], V->sec[i])
                                        : allowed(t->sym, M->sec[a], V->sec[i], V->sec[j]);
        if (!ok) continue;
        Block b{};
        b.lab[0] = a;
        b.lab[1] = i;
        b.lab[2] = j;
        b.rows = V->mult[i];
        b.cols = V->mult[j];
        b.weight = sdim(t->sym, V->sec[side == HTN_SIDE_LEFT ? i : j]);
        t->blocks.push_back(b);
      }
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_tensor_create_like(const htn_tensor* src, htn_tensor** out) {
  if (!src || !out) return HTN_ERR_INVALID;
  htn_ctx* ctx = src->ctx;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, src->kind, src->sym);
  t->s0 = src->s0;
  t->s1 = src->s1;
  t->legs = src->legs;
  t->legs2 = src->legs2;
  t->mid = src->mid;
  t->identity_level = src->identity_level;
  t->blocks = src->blocks;
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

// two-site tensor in the fusion-tree basis (l,s1 -> m), (m,s2 -> r); block order = oracle/twosite.py:
// coupled sector r, then s2, then m (canonical sector order), then s1, then l
int32_t htn_tensor_create_mps2(htn_ctx* ctx, const htn_space* Vl, const htn_legs* P1, const htn_legs* P2,
                               const htn_space* Vr, htn_tensor** out) {
  if (!ctx || !Vl || !P1 || !P2 || !Vr || !out) return HTN_ERR_INVALID;
  if (Vl->sym != P1->sym || Vr->sym != P1->sym || P2->sym != P1->sym) return ctx->fail(HTN_ERR_INVALID, "symmetry kinds differ");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, HTN_T_MPS2, P1->sym);
  const int sym = t->sym;
  t->s0 = *Vl;
  t->s1 = *Vr;
  t->legs = *P1;
  t->legs2 = *P2;
  t->legs2.ctx = ctx;
  const int nl = (int)Vl->sec.size(), n1 = (int)P1->sec.size(), n2 = (int)P2->sec.size(), nr = (int)Vr->sec.size();
  auto fuse = [&](Sector a, Sector b) {
    std::vector<Sector> out;
    const int p = (a.p + b.p) & 1, n = a.n + b.n;
    if (sym == HTN_SYM_SU2U1)
      for (int q = std::abs(a.q - b.q); q <= a.q + b.q; q += 2) out.push_back(Sector{p, q, n});
    else
      out.push_back(Sector{p, a.q + b.q, n});
    return out;
  };
  // all intermediate sectors that occur in some block
  std::vector<Sector> mids;
  for (int l = 0; l < nl; ++l)
    for (int s1 = 0; s1 < n1; ++s1)
      for (Sector m : fuse(Vl->sec[l], P1->sec[s1])) {
        bool used = false;
        for (int s2 = 0; s2 < n2 && !used; ++s2)
          for (int r = 0; r < nr && !used; ++r) used = allowed(sym, m, P2->sec[s2], Vr->sec[r]);
        if (used && std::find(mids.begin(), mids.end(), m) == mids.end()) mids.push_back(m);
      }
  std::sort(mids.begin(), mids.end(), [sym](Sector a, Sector b) { return canonical_less(sym, a, b); });
  t->mid = mids;
  for (int r = 0; r < nr; ++r)
    for (int s2 = 0; s2 < n2; ++s2)
      for (int m = 0; m < (int)mids.size(); ++m) {
        if (!allowed(sym, mids[m], P2->sec[s2], Vr->sec[r])) continue;
        for (int s1 = 0; s1 < n1; ++s1)
          for (int l = 0; l < nl; ++l)
            if (allowed(sym, Vl->sec[l], P1->sec[s1], mids[m])) {
              Block b{};
              b.lab[0] = l;
              b.lab[1] = s1;
              b.lab[2] = m;
              b.lab[3] = s2;
              b.lab[4] = r;
              b.rows = Vl->mult[l];
              b.cols = Vr->mult[r];
              b.weight = sdim(sym, Vr->sec[r]);
              t->blocks.push_back(b);
            }
      }
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_tensor_mid_sectors(const htn_tensor* t, int32_t* n, int32_t* labels) {
  if (!t) return HTN_ERR_INVALID;
  if (n) *n = (int32_t)t->mid.size();
  if (labels)
    for (size_t i = 0; i < t->mid.size(); ++i) {
      labels[3 * i] = t->mid[i].p;
      labels[3 * i + 1] = t->mid[i].q;
      labels[3 * i + 2] = t->mid[i].n;
    }
  return HTN_OK;
}

int32_t htn_tensor_blocktable5(const htn_tensor* t, int32_t* nblocks, int64_t* nelem, int32_t* labels, int32_t* rows,
                               int32_t* cols, int64_t* offsets) {
  if (!t) return HTN_ERR_INVALID;
  if (nblocks) *nblocks = (int32_t)t->blocks.size();
  if (nelem) *nelem = t->hsize;
  for (size_t i = 0; i < t->blocks.size(); ++i) {
    const Block& b = t->blocks[i];
    if (labels)
      for (int k = 0; k < 5; ++k) labels[5 * i + k] = b.lab[k];
    if (rows) rows[i] = b.rows;
    if (cols) cols[i] = b.cols;
    if (offsets) offsets[i] = b.hoff;
  }
  return HTN_OK;
}

// blockwise transposed companion of an MPS tensor: blocks (l,s,r) stored as [n_r x n_l], ordered and
// laid out by LEFT sector (one row-major panel per l) -- the operand layout of the transfer GEMMs
// and the panel layout of the LQ gauge step
int32_t htn_tensor_create_transposed(const htn_tensor* src, htn_tensor** out) {
  if (!src || !out) return HTN_ERR_INVALID;
  htn_ctx* ctx = src->ctx;
  if (src->kind != HTN_T_MPS) return ctx->fail(HTN_ERR_INVALID, "create_transposed: source must be an MPS tensor");
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  HTN_TRY
  htn_tensor* t = new_tensor(ctx, HTN_T_MPST, src->sym);
  t->s0 = src->s0;
  t->s1 = src->s1;
  t->legs = src->legs;
  const int nl = (int)src->s0.sec.size(), ns = (int)src->legs.sec.size(), nr = (int)src->s1.sec.size();
  for (int l = 0; l < nl; ++l)
    for (int s = 0; s < ns; ++s)
      for (int r = 0; r < nr; ++r)
        if (allowed(t->sym, src->s0.sec[l], src->legs.sec[s], src->s1.sec[r])) {
          Block b{};
          b.lab[0] = l;
          b.lab[1] = s;
          b.lab[2] = r;
          b.rows = src->s1.mult[r];
          b.cols = src->s0.mult[l];
          b.weight = sdim(t->sym, src->s1.sec[r]);
          t->blocks.push_back(b);
        }
  int32_t rc = finalize_tensor(t);
  if (rc != HTN_OK) {
    htn_tensor_destroy(t);
    return rc;
  }
  *out = t;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_tensor_destroy(htn_tensor* t) {
  if (!t) return HTN_OK;
  cudaSetDevice(t->ctx->device);
  cudaStreamSynchronize(t->ctx->stream);
  if (t->d) cudaFree(t->d);
  if (t->dblocks) cudaFree(t->dblocks);
  if (t->dchunks) cudaFree(t->dchunks);
  if (t->d_tmaps) cudaFree(t->d_tmaps);
  for (auto& kv : t->devtables) cudaFree(kv.second.first);
  delete t;
  return HTN_OK;
}

int32_t htn_tensor_kind(const htn_tensor* t) { return t ? t->kind : HTN_ERR_INVALID; }

int32_t htn_tensor_device_ptr(const htn_tensor* t, void** ptr, int64_t* nelem_padded) {
  if (!t || !ptr) return HTN_ERR_INVALID;
  *ptr = t->d;
  if (nelem_padded) *nelem_padded = t->dsize;
  return HTN_OK;
}

int32_t htn_tensor_space(const htn_tensor* t, int32_t which, htn_space** out) {
  if (!t || !out) return HTN_ERR_INVALID;
  HTN_TRY
  *out = new htn_space(which == 0 ? t->s0 : t->s1);
  return HTN_OK;
  HTN_CATCH(t->ctx)
}

int32_t htn_tensor_blocktable(const htn_tensor* t, int32_t* nblocks, int64_t* nelem, int32_t* labels,
                              int32_t* rows, int32_t* cols, int64_t* offsets) {
  if (!t) return HTN_ERR_INVALID;
  if (nblocks) *nblocks = (int32_t)t->blocks.size();
  if (nelem) *nelem = t->hsize;
  for (size_t i = 0; i < t->blocks.size(); ++i) {
    const Block& b = t->blocks[i];
    if (labels) {
      labels[3 * i] = b.lab[0];
      labels[3 * i + 1] = b.lab[1];
      labels[3 * i + 2] = b.lab[2];
    }
    if (rows) rows[i] = b.rows;
    if (cols) cols[i] = b.cols;
    if (offsets) offsets[i] = b.hoff;
  }
  return HTN_OK;
}

static int32_t ensure_stage(htn_ctx* ctx, int64_t n) {
  if (ctx->stage_cap >= n) return HTN_OK;
  if (ctx->stage) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_cap = 0;
  }
  int64_t cap = std::max<int64_t>(n, 1 << 20);
  CU(ctx, cudaMalloc(&ctx->stage, cap * sizeof(double)));
  ctx->stage_cap = cap;
  return HTN_OK;
}

int32_t htn_upload_locked(htn_tensor* t, const double* host, int64_t nelem) {
  htn_ctx* ctx = t->ctx;
  if (nelem != t->hsize) return ctx->fail(HTN_ERR_SHAPE, "upload: element count does not match the block table");
  if (nelem == 0) return HTN_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  int32_t rc = ensure_stage(ctx, nelem);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(ctx->stage, host, nelem * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  launch_pack(t->dblocks, t->dchunks, t->nchunks, ctx->stage, t->d, ctx->stream);
  CU(ctx, cudaGetLastError());
  return HTN_OK;
}

int32_t htn_download_locked(const htn_tensor* t, double* host, int64_t nelem) {
  htn_ctx* ctx = t->ctx;
  if (nelem != t->hsize) return ctx->fail(HTN_ERR_SHAPE, "download: element count does not match the block table");
  if (nelem == 0) return HTN_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  int32_t rc = ensure_stage(ctx, nelem);
  if (rc) return rc;
  launch_unpack(t->dblocks, t->dchunks, t->nchunks, t->d, ctx->stage, ctx->stream);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(host, ctx->stage, nelem * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return HTN_OK;
}

int32_t htn_tensor_upload(htn_tensor* t, const double* host, int64_t nelem) {
  if (!t || (!host && nelem > 0)) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(t->ctx->mu);
  int32_t rc = htn_upload_locked(t, host, nelem);
  if (rc) return rc;
  // the host buffer belongs to the caller: do not return before the copy has read it
  CU(t->ctx, cudaStreamSynchronize(t->ctx->stream));
  return HTN_OK;
}

int32_t htn_tensor_download(const htn_tensor* t, double* host, int64_t nelem) {
  if (!t || (!host && nelem > 0)) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(t->ctx->mu);
  return htn_download_locked(t, host, nelem);
}

// ---- MPO ------------------------------------------------------------------------------
int32_t htn_mpo_create(htn_ctx* ctx, const htn_legs* Ml, const htn_legs* P, const htn_legs* Mr, int32_t nnz,
                       const int32_t* idx, const int32_t* clabel, const double* val, htn_mpo** out) {
  if (!ctx || !Ml || !P || !Mr || !out || nnz < 0 || (nnz > 0 && (!idx || !clabel || !val))) return HTN_ERR_INVALID;
  if (Ml->sym != P->sym || Mr->sym != P->sym) return ctx->fail(HTN_ERR_INVALID, "symmetry kinds differ");
  HTN_TRY
  htn_mpo* w = new htn_mpo();
  w->ctx = ctx;
  w->sym = P->sym;
  w->Ml = *Ml;
  w->P = *P;
  w->Mr = *Mr;
  for (int i = 0; i < nnz; ++i) {
    MpoEntry e{idx[4 * i], idx[4 * i + 1], idx[4 * i + 2], idx[4 * i + 3],
               Sector{clabel[3 * i], clabel[3 * i + 1], clabel[3 * i + 2]}, val[i]};
    bool in_range = e.a >= 0 && e.a < (int)Ml->sec.size() && e.b >= 0 && e.b < (int)Mr->sec.size() && e.sp >= 0 &&
                    e.sp < (int)P->sec.size() && e.s >= 0 && e.s < (int)P->sec.size();
    if (!in_range) {
      delete w;
      return ctx->fail(HTN_ERR_INVALID, "MPO entry index out of range");
    }
    if (!allowed(w->sym, Ml->sec[e.a], P->sec[e.sp], e.c) || !allowed(w->sym, P->sec[e.s], Mr->sec[e.b], e.c)) {
      delete w;
      return ctx->fail(HTN_ERR_INVALID, "MPO entry violates the fusion rules");
    }
    w->entries.push_back(e);
  }
  *out = w;
  return HTN_OK;
  HTN_CATCH(ctx)
}

// Wigner-Eckart projection of an invariant dense MPO tensor onto its reduced entries.
// dense[a m_a][s' m'][s m][b m_b], row-major, every multiplet expanded with m = -j .. +j.
int32_t htn_mpo_create_dense(htn_ctx* ctx, const htn_legs* Ml, const htn_legs* P, const htn_legs* Mr,
                             const double* dense, double tol, htn_mpo** out) {
  // host-only: ctx may be NULL
  if (!Ml || !P || !Mr || !dense || !out) return HTN_ERR_INVALID;
  if (Ml->sym != P->sym || Mr->sym != P->sym) return HTN_ERR_INVALID;
  HTN_TRY
  const int sym = P->sym;
  auto offsets = [&](const htn_legs* L, std::vector<int>& off) {
    int acc = 0;
    for (const Sector& s : L->sec) {
      off.push_back(acc);
      acc += sdim(sym, s);
    }
    return acc;
  };
  std::vector<int> ol, op, orr;
  const int Dl = offsets(Ml, ol), d = offsets(P, op), Dr = offsets(Mr, orr);
  auto at = [&](int a, int sp, int s, int b) -> double { return dense[(((int64_t)a * d + sp) * d + s) * Dr + b]; };
  // coupling coefficient <a ma; b mb | c mc> with state index i <-> m = -j + i (abelian: 1)
  auto cg = [&](Sector A, int ia, Sector B, int ib, Sector Cc, int ic) -> double {
    if (sym != HTN_SYM_SU2U1) return 1.0;
    return cg_su2(A.q, -A.q + 2 * ia, B.q, -B.q + 2 * ib, Cc.q, -Cc.q + 2 * ic);
  };
  std::vector<MpoEntry> entries;
  std::vector<double> rec((size_t)Dl * d * d * Dr, 0.0);
  for (int a = 0; a < (int)Ml->sec.size(); ++a)
    for (int sp = 0; sp < (int)P->sec.size(); ++sp) {
      const Sector ca = Ml->sec[a], csp = P->sec[sp];
      const int p2 = (ca.p + csp.p) & 1, n2 = ca.n + csp.n;
      const int qlo = sym == HTN_SYM_SU2U1 ? std::abs(ca.q - csp.q) : ca.q + csp.q, qhi = ca.q + csp.q;
      for (int q = qlo; q <= qhi; q += 2) {
        const Sector c{p2, q, n2};
        const int dc = sdim(sym, c);
        for (int s = 0; s < (int)P->sec.size(); ++s)
          for (int b = 0; b < (int)Mr->sec.size(); ++b) {
            const Sector cs = P->sec[s], cb = Mr->sec[b];
            if (!allowed(sym, cs, cb, c)) continue;
            const int da = sdim(sym, ca), dsp = sdim(sym, csp), ds = sdim(sym, cs), db = sdim(sym, cb);
            double w = 0.0;
            for (int x = 0; x < da; ++x)
              for (int y = 0; y < dsp; ++y)
                for (int z = 0; z < ds; ++z)
                  for (int v = 0; v < db; ++v) {
                    const double val = at(ol[a] + x, op[sp] + y, op[s] + z, orr[b] + v);
                    if (val == 0.0) continue;
                    for (int mc = 0; mc < dc; ++mc) w += val * cg(ca, x, csp, y, c, mc) * cg(cs, z, cb, v, c, mc);
                  }
            w /= dc;
            if (std::fabs(w) <= tol) continue;
            entries.push_back(MpoEntry{a, sp, s, b, c, w});
            for (int x = 0; x < da; ++x)
              for (int y = 0; y < dsp; ++y)
                for (int z = 0; z < ds; ++z)
                  for (int v = 0; v < db; ++v) {
                    double g = 0.0;
                    for (int mc = 0; mc < dc; ++mc) g += cg(ca, x, csp, y, c, mc) * cg(cs, z, cb, v, c, mc);
                    rec[((((int64_t)(ol[a] + x)) * d + op[sp] + y) * d + op[s] + z) * Dr + orr[b] + v] += w * g;
                  }
          }
      }
    }
  double err = 0.0;
  for (size_t i = 0; i < rec.size(); ++i) err = std::max(err, std::fabs(rec[i] - dense[i]));
  if (err > 1e-10) {
    const char* msg = "mpo_create_dense: tensor is not invariant under the symmetry (re-expansion mismatch)";
    if (ctx) ctx->err = msg; else g_noctx_err = msg;
    return HTN_ERR_INVALID;
  }
  htn_mpo* w = new htn_mpo();
  w->ctx = ctx;
  w->sym = sym;
  w->Ml = *Ml;
  w->P = *P;
  w->Mr = *Mr;
  w->entries = entries;
  *out = w;
  return HTN_OK;
  HTN_CATCH(ctx)
}

int32_t htn_mpo_entries(const htn_mpo* w, int32_t* nnz, int32_t* idx, int32_t* clabel, double* val) {
  if (!w) return HTN_ERR_INVALID;
  if (nnz) *nnz = (int32_t)w->entries.size();
  for (size_t i = 0; i < w->entries.size(); ++i) {
    const MpoEntry& e = w->entries[i];
    if (idx) {
      idx[4 * i] = e.a;
      idx[4 * i + 1] = e.sp;
      idx[4 * i + 2] = e.s;
      idx[4 * i + 3] = e.b;
    }
    if (clabel) {
      clabel[3 * i] = e.c.p;
      clabel[3 * i + 1] = e.c.q;
      clabel[3 * i + 2] = e.c.n;
    }
    if (val) val[i] = e.w;
  }
  return HTN_OK;
}

int32_t htn_mpo_destroy(htn_mpo* w) {
  delete w;
  return HTN_OK;
}

// ---- vector algebra -------------------------------------------------------------------
static bool same_structure(const htn_tensor* x, const htn_tensor* y) {
  if (x->kind != y->kind || x->sym != y->sym || x->blocks.size() != y->blocks.size() || x->dsize != y->dsize)
    return false;
  // same layout is not enough: the tensors must live on the same graded spaces
  if (!(x->s0.sec == y->s0.sec) || x->s0.mult != y->s0.mult || !(x->s1.sec == y->s1.sec) || x->s1.mult != y->s1.mult ||
      !(x->legs.sec == y->legs.sec))
    return false;
  for (size_t i = 0; i < x->blocks.size(); ++i) {
    const Block &a = x->blocks[i], &b = y->blocks[i];
    if (a.lab[0] != b.lab[0] || a.lab[1] != b.lab[1] || a.lab[2] != b.lab[2] || a.lab[3] != b.lab[3] ||
        a.lab[4] != b.lab[4] || a.rows != b.rows || a.cols != b.cols)
      return false;
  }
  return true;
}

bool htn_same_structure(const htn_tensor* x, const htn_tensor* y) { return same_structure(x, y); }

int32_t htn_tensor_dot(const htn_tensor* x, const htn_tensor* y, double* out) {
  if (!x || !y || !out) return HTN_ERR_INVALID;
  htn_ctx* ctx = x->ctx;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!same_structure(x, y)) return ctx->fail(HTN_ERR_SHAPE, "dot: tensors differ in structure");
  CU(ctx, cudaSetDevice(ctx->device));
  if (x->nchunks == 0) {
    *out = 0.0;
    return HTN_OK;
  }
  if (ctx->red_cap < x->nchunks + 1) {
    if (ctx->red) cudaFree(ctx->red);
    ctx->red_cap = std::max<int64_t>(x->nchunks + 1, 4096);
    CU(ctx, cudaMalloc(&ctx->red, ctx->red_cap * sizeof(double)));
  }
  launch_dot(x->dblocks, x->dchunks, x->nchunks, x->d, y->d, ctx->red + 1, ctx->red, ctx->stream);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(ctx->red_host, ctx->red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  *out = ctx->red_host[0];
  return HTN_OK;
}

int32_t htn_tensor_axpby(double alpha, const htn_tensor* x, double beta, htn_tensor* y) {
  if (!x || !y) return HTN_ERR_INVALID;
  htn_ctx* ctx = x->ctx;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  if (!same_structure(x, y)) return ctx->fail(HTN_ERR_SHAPE, "axpby: tensors differ in structure");
  CU(ctx, cudaSetDevice(ctx->device));
  launch_axpby(alpha, x->d, beta, y->d, x->dsize, ctx->stream);
  CU(ctx, cudaGetLastError());
  return HTN_OK;
}

int32_t htn_tensor_transpose(const htn_tensor* src, htn_tensor* dst, int32_t weighted) {
  if (!src || !dst) return HTN_ERR_INVALID;
  htn_ctx* ctx = src->ctx;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  const int mode = !weighted ? 0 : (src->kind == HTN_T_MPS ? 1 : 2);
  return htn::t_transpose(src, dst, mode);
}

// ---- test hooks -----------------------------------------------------------------------
int32_t htn_network_coefficient(int32_t sym, const int32_t* L, double* out) {
  if (!L || !out) return HTN_ERR_INVALID;
  auto S = [&](int i) { return Sector{L[3 * i], L[3 * i + 1], L[3 * i + 2]}; };
  *out = network(sym, S(0), S(1), S(2), S(3), S(4), S(5), S(6), S(7), S(8));
  return HTN_OK;
}

int32_t htn_probe_fp64_peak(htn_ctx* ctx, int32_t which, double* tflops) {
  if (!ctx || !tflops) return HTN_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> g(ctx->mu);
  CU(ctx, cudaSetDevice(ctx->device));
  *tflops = probe_fp64(which, ctx->sm_count, ctx->stream);
  CU(ctx, cudaGetLastError());
  return HTN_OK;
}

}  // extern "C"
