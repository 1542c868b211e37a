// Internal structures of libhtn (host-side sector bookkeeping + device arena handles).
// Host side replaces TensorKit's GradedSpace / fusion-tree bookkeeping (TensorKit 0.14.6,
// /root/reference/Manifest.toml:1156; not vendored); device side replaces its flat
// Vector{T} block store with a padded HBM arena + block tables.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "htn.h"

namespace htn {

struct Sector {
  int32_t p, q, n;  // parity, 2j (SU2U1) or 2Sz (U1U1), U(1) charge
  bool operator==(const Sector& o) const { return p == o.p && q == o.q && n == o.n; }
  bool operator<(const Sector& o) const { return std::tie(p, q, n) < std::tie(o.p, o.q, o.n); }
};

int sdim(int sym, Sector s);
bool allowed(int sym, Sector a, Sector b, Sector c);
bool canonical_less(int sym, Sector a, Sector b);
// <j1 m1; j2 m2 | j3 m3>, all arguments doubled
double cg_su2(int tj1, int tm1, int tj2, int tm2, int tj3, int tm3);
// N(l',s',r'; l,s,r; a,b,c): full contraction of the six coupling tensors of the H_AC network
double network(int sym, Sector lp, Sector sp, Sector rp, Sector l, Sector s, Sector r, Sector a,
               Sector b, Sector c);

inline int even_up(int x) { return (x + 1) & ~1; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Block {
  int32_t lab[3];
  int32_t rows, cols, ld;
  int64_t off;   // device offset (elements) inside the tensor's arena slice
  int64_t hoff;  // packed host offset (elements)
  int32_t weight;  // quantum dimension of the coupled sector (inner-product weight)
};

// device-side copy of a block record (pack / unpack / dot kernels)
struct DevBlock {
  long long off, hoff;
  int rows, cols, ld, weight;
};

}  // namespace htn

struct htn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;
  double* stage = nullptr;  // device staging buffer for packed host <-> padded device copies
  int64_t stage_cap = 0;
  double* red = nullptr;  // reduction scratch
  int64_t red_cap = 0;
  double* red_host = nullptr;  // pinned
  int sm_count = 148;
  int32_t fail(int32_t code, const std::string& msg) {
    err = msg;
    return code;
  }
};

struct htn_space {
  htn_ctx* ctx;
  int sym;
  std::vector<htn::Sector> sec;
  std::vector<int32_t> mult;
};

struct htn_legs {
  htn_ctx* ctx;
  int sym;
  std::vector<htn::Sector> sec;
};

struct htn_tensor {
  htn_ctx* ctx;
  int kind;
  int sym;
  htn_space s0, s1;  // MPS: Vl, Vr; BOND: V,V; ENV: V,V
  htn_legs legs;     // MPS: P; ENV: M
  int identity_level = -1;
  std::vector<htn::Block> blocks;
  std::map<std::tuple<int, int, int>, int> index;
  double* d = nullptr;
  int64_t dsize = 0;  // padded device elements
  int64_t hsize = 0;  // packed host elements
  htn::DevBlock* dblocks = nullptr;
  // row-chunk table for pack/unpack/dot kernels: (block, row0, nrows)
  int* dchunks = nullptr;
  int nchunks = 0;
  int find(int a, int b, int c) const {
    auto it = index.find(std::make_tuple(a, b, c));
    return it == index.end() ? -1 : it->second;
  }
};

struct MpoEntry {
  int32_t a, sp, s, b;
  htn::Sector c;
  double w;
};

struct htn_mpo {
  htn_ctx* ctx;
  int sym;
  htn_legs Ml, P, Mr;
  std::vector<MpoEntry> entries;
};

// ---- device work tables (shared between planner and kernels) ---------------------------
namespace htn {

enum Base { B_GL = 0, B_GR = 1, B_X = 2, B_Y = 3, B_T = 4, B_U = 5, B_P = 6, B_COUNT = 7 };

// References in the device tables are resolved by a post-pass of the planner: arrays that
// live as long as the plan (GL, GR, T, U) become absolute pointers (REF_ABS); the apply's
// input / output vectors stay relative (REF_X / REF_Y) and are passed per launch.
enum Ref { REF_ABS = 0, REF_X = 1, REF_Y = 2 };
struct Bases {
  const double* x;
  double* y;
};
__host__ __device__ inline const double* resolve(long long v, int base, const Bases& b) {
  return base == REF_ABS ? reinterpret_cast<const double*>(v) : (base == REF_X ? b.x + v : b.y + v);
}

// one K-segment of a grouped GEMM work item:  C_tile += coef * A_seg[mt x K] * B_seg[K x nt]
// If nsrc > 0 the A operand is not read from one array but assembled on the fly as
// sum_j srcs[src_begin + j].coef * (source block j) -- the stage-W recoupling mix fused into the
// operand load of stage R (all sources share the shape and leading dimension lda).
struct GemmSeg {
  long long a_off, b_off;  // element offsets from the bases (or absolute pointers, see Ref)
  int a_base, b_base;
  int lda, ldb;
  int K;
  int nsrc, src_begin;
  int pad_;
};

struct GemmItem {
  long long c_off;
  int c_base, ldc;
  int mt, nt;  // tile extent (<= BM, BN)
  int seg_begin, seg_end;
  int nchunks;  // sum over segments of ceil(K / BK)
  int beta;     // 0: overwrite, 1: accumulate
  int layout;   // 0: flex = M, strips over N ; 1: flex = N, strips over M (htn_kernels.cu)
  int pad_;
};

struct MixSrc {
  long long off;
  int base, pad_;
  double coef;
};

struct MixTarget {
  long long off;
  int base;
  int nelem;  // rows * ld (flat, padded)
  int src_begin, src_end;
};

struct MixChunk {
  int target, elem0, nelem, pad_;
};

void launch_gemm(const GemmItem* items, const GemmSeg* segs, const MixSrc* srcs, int nitems, Bases bases,
                 int grid, cudaStream_t st);
void launch_mix(const MixTarget* tg, const MixSrc* src, const MixChunk* chunks, int nchunks, Bases bases,
                cudaStream_t st);
void launch_pack(const DevBlock* blocks, const int* chunks, int nchunks, const double* packed, double* padded,
                 cudaStream_t st);
void launch_unpack(const DevBlock* blocks, const int* chunks, int nchunks, const double* padded, double* packed,
                   cudaStream_t st);
void launch_dot(const DevBlock* blocks, const int* chunks, int nchunks, const double* x, const double* y,
                double* partial, double* out, cudaStream_t st);
void launch_axpby(double alpha, const double* x, double beta, double* y, long long n, cudaStream_t st);
int gemm_max_ctas_per_sm();
double probe_fp64(int which, int sm_count, cudaStream_t st);

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16;

}  // namespace htn

struct htn_plan {
  htn_ctx* ctx;
  const htn_tensor* GL;
  const htn_tensor* GR;
  htn_tensor* like;  // private structural copy (no data use)
  double* T = nullptr;
  htn::MixSrc* gsrcs = nullptr;  // fused stage-W sources of the stage-R A operand (HTN_FUSE_W=1)
  double* U = nullptr;
  double* Pp = nullptr;  // split-K partial outputs of stage R: nsplit_max copies of the y layout
  int64_t t_elems = 0, u_elems = 0, p_elems = 0;  // u_elems: size U would have (never materialised)
  // device tables
  htn::GemmItem* itemsL = nullptr;
  htn::GemmSeg* segsL = nullptr;
  int nitemsL = 0, nsegsL = 0;
  htn::GemmItem* itemsR = nullptr;
  htn::GemmSeg* segsR = nullptr;
  int nitemsR = 0, nsegsR = 0;
  htn::MixTarget* mixT = nullptr;
  htn::MixSrc* mixS = nullptr;
  htn::MixChunk* mixC = nullptr;   // chunks [0, nmixCU) -> U targets, [nmixCU, nmixC) -> y targets
  int nmixT = 0, nmixS = 0, nmixC = 0, nmixCU = 0;
  int gridL = 0, gridR = 0;
  double stats[12] = {0};
  // host staging tensors for htn_heff_apply_host
  htn_tensor* hx = nullptr;
  htn_tensor* hy = nullptr;
};
