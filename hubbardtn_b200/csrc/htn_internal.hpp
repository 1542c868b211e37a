// Internal structures of libhtn (host-side sector bookkeeping + device arena handles).
// Host side replaces TensorKit's GradedSpace / fusion-tree bookkeeping (TensorKit 0.14.6,
// /root/reference/Manifest.toml:1156; not vendored); device side replaces its flat
// Vector{T} block store with a padded HBM arena + block tables.
#pragma once
#include <cuda_runtime.h>

#include <array>
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

// Host -> device table upload ORDERED ON THE LIBRARY STREAM.  A plain cudaMemcpy runs on the legacy
// default stream, which does not synchronise with the (non-blocking) library stream, and returns
// before the DMA of a small pageable buffer has landed: a kernel launched right after could read a
// half-written table.  Returns after the copy has completed (the host buffer may be freed).
inline cudaError_t h2d_on_stream(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

inline int even_up(int x) { return (x + 1) & ~1; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Block {
  int32_t lab[5];  // kinds with 3 labels leave lab[3], lab[4] = 0
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
  std::recursive_mutex mu;  // public entry points may nest (drivers call the primitives)
  std::string err;
  double* stage = nullptr;  // device staging buffer for packed host <-> padded device copies
  int64_t stage_cap = 0;
  double* red = nullptr;  // reduction scratch
  int64_t red_cap = 0;
  double* red_host = nullptr;  // pinned
  // Krylov workspace: basis vectors, device scalars (+ pinned mirror), multidot partials
  double* kry_V = nullptr;
  int64_t kry_cap = 0;
  double* kry_scal = nullptr;       // 512 doubles on the device
  double* kry_scal_host = nullptr;  // pinned mirror
  double* kry_partial = nullptr;
  int64_t kry_partial_cap = 0;
  // instantiated CUDA graphs of the Lanczos orthogonalisation step, keyed by (block table, basis, length, step, chunks,
  // variant); dropped whenever the Krylov buffers move (htn_linalg.cpp)
  std::map<std::array<long long, 6>, void*> kry_graphs;
  int* d_status = nullptr;  // device-side failure flags (QR rank deficiency, ...)
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
  uint64_t uid = 0;  // unique per created tensor (keys of cached device tables)
  int kind;
  int sym;
  htn_space s0, s1;  // MPS: Vl, Vr; BOND: V,V; ENV: V,V
  htn_legs legs;     // MPS: P; ENV: M; MPS2: P of the first site
  htn_legs legs2;    // MPS2: P of the second site
  std::vector<htn::Sector> mid;  // MPS2: intermediate sectors m (canonical order); lab[2] indexes this list
  std::map<std::array<int, 5>, int> index5;
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
  // Environment tensors are stored as STACKED PANELS: all blocks with the same last label (ENVL: the contracted
  // sector l of GL[a,l',l]; ENVR: the output sector r' of GR[b,r,r']) lie back to back as one row-major matrix
  // (ENVL rows ordered (l', a), ENVR rows ordered (r, b); blocks of the identity level last), so that a GEMM over
  // the panel sees one tall operand (csrc/htn_stackl.cuh).  The block table / host order is unaffected.
  struct Panel {
    int64_t off;  // device offset of the panel
    int rows, rows_active, cols, ld;  // total rows, rows before the identity-level blocks, columns, leading dimension
  };
  std::vector<Panel> panels;      // by sector index lab[2] (empty for non-environment kinds)
  std::vector<int> block_prow;    // first row of block i inside its panel
  unsigned char* d_tmaps = nullptr;  // ENVL/ENVR: one CUtensorMap (128 B) per panel: box 16 (cols) x 64 (rows), SWIZZLE_128B
  // cached device tables (transpose tiles, level fills, QR panels) keyed by (partner, mode)
  std::map<std::pair<const void*, int>, std::pair<void*, int>> devtables;
  int find5(int l, int s1, int m, int s2, int r) const {
    auto it = index5.find(std::array<int, 5>{l, s1, m, s2, r});
    return it == index5.end() ? -1 : it->second;
  }
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

// A contraction program addresses memory through "references": base 0 = absolute device
// pointer (program workspace, resolved when the program is finalised), base k>0 = offset into
// the k-th tensor bound at launch time (slot k-1).  The same program therefore serves any set
// of tensors with the planned block structure (x/y of a Krylov iteration, AL or AR, ...).
constexpr int MAX_SLOTS = 6;
struct Bases {
  const double* p[MAX_SLOTS];
};
__host__ __device__ inline const double* resolve(long long v, int base, const Bases& b) {
  return base == 0 ? reinterpret_cast<const double*>(v) : b.p[base - 1] + v;
}

// one K-segment of a grouped GEMM work item:  C_tile += A_seg[mt x K] * B_seg[K x nt]
struct GemmSeg {
  long long a_off, b_off;
  int a_base, b_base;
  int lda, ldb;
  int K;
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
  int group;    // host scheduling only: items that read the same large operand block share a group (>= 0)
};

// one job of the stacked stage-L GEMM (htn_stackl.cuh):  C[M x nt] = A[M x K] . B[K x nt], B resident in shared memory
struct StackJob {
  long long a_off, b_off, c_off;
  int a_base, b_base, c_base;
  int lda, ldb, ldc;
  int K;    // contraction length (multiplicity of the contracted sector)
  int nt;   // valid output columns (<= 64)
  int nb;   // doubles copied per slab row: even, >= nt, <= 8 ceil(nt / 8)
  int M;    // rows of the run (tiles of 64)
  int tmap; // >= 0: A is read through the 2-D tensor map tmaps[tmap] (TMA), rows arow .. arow + M of that panel
  int arow;
  int wave; // >= 0: one arrival per consumer warp on that wave's counter when the job is done; < 0: tile-level (tile0)
  int tile0; // first entry of this job's tiles in StackArgs::tile_waves
  int pad_;
};
constexpr int STACK_SLAB_ELEMS = 10080;
inline bool stack_job_fits(int K, int nt) { return ((K + 3) & ~3) * (((nt + 7) & ~7) + 4) <= STACK_SLAB_ELEMS && nt <= 64; }
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
  int target, elem0, nelem, pad_;  // pad_: wave the chunk waits for in the fused stage (-1: none)
};
// self-contained chunk record of the fused stage (one 32-byte load instead of chunk -> target): same order as MixChunk
struct MixChunkX {
  long long off;  // destination block
  int base;
  int src_begin, nsrc;
  int elem0, nelem;
  int wave;
};

// Arguments of the fused stage L + W launch.  Counters (64-bit, never reset: `epoch` = 1-based launch number):
//   ctr[0] job ticket, ctr[1] mix ticket, ctr[2] watchdog flag, ctr[4 + w] consumer-warp arrivals of wave w,
//   ctr[4 + nwaves + w] finished mix chunks of wave w (back-pressure: jobs of wave w + lag wait for them),
//   ctr[4 + 2 nwaves ..] timeline probe (HTN_STACK_DEBUG & 32): first start, last job end, last mix end (globaltimer ns).
struct StackArgs {
  const StackJob* jobs;
  int njobs;
  const unsigned char* tmaps;
  const MixTarget* mt;
  const MixSrc* ms;
  const MixChunk* mc;  // sorted by wave; pad_ = wave the chunk waits for (-1: none)
  const MixChunkX* mcx;
  int nmix;
  const int2* tile_waves;  // per tile of the tile-level jobs: first and last wave it reports to
  const int* wave_need;  // [w]: arrivals that complete wave w (= SL_NCONS * jobs of the wave); [nwaves + w]: mix chunks of wave w
  int nwaves;
  int mix_lag;           // > 0: the jobs of wave w are held back until the mix of wave w - mix_lag is done (T stays in L2)
  unsigned long long* ctr;
  unsigned long long epoch;
  int dbg;
  unsigned long long* dbg_ts;  // HTN_STACK_DEBUG & 32: per mix chunk (start, end) globaltimer stamps
};

void launch_stack_gemm(const StackArgs& args, const Bases& bases, int grid, cudaStream_t st);

int stack_gemm_ctas_per_sm();
int stack_gemm_groups();      // consumer groups per CTA (HTN_STACK_NG)
int stack_gemm_cons_warps();  // consumer warps per CTA: arrivals per job on the wave counters

void launch_gemm(const GemmItem* items, const GemmSeg* segs, int nitems, const Bases& bases, int grid,
                 cudaStream_t st);
void launch_mix(const MixTarget* tg, const MixSrc* src, const MixChunk* chunks, int nchunks, const Bases& bases,
                cudaStream_t st);
void launch_pack(const DevBlock* blocks, const int* chunks, int nchunks, const double* packed, double* padded,
                 cudaStream_t st);
void launch_unpack(const DevBlock* blocks, const int* chunks, int nchunks, const double* padded, double* packed,
                   cudaStream_t st);
void launch_dot(const DevBlock* blocks, const int* chunks, int nchunks, const double* x, const double* y,
                double* partial, double* out, cudaStream_t st);
void launch_axpby(double alpha, const double* x, double beta, double* y, long long n, cudaStream_t st);
// out[j] = <V_j, w> (weighted) for j < nvec, V_j = V + j * stride; deterministic two-pass
void launch_multidot(const DevBlock* blocks, const int* chunks, int nchunks, const double* V, long long stride,
                     int nvec, const double* w, double* partial, double* out, cudaStream_t st);
// w += sum_j sign * h[j] V_j   (h on the device)
void launch_multiaxpy(const double* V, long long stride, int nvec, const double* h, double sign, double* w,
                      long long n, cudaStream_t st);
// y = x * (*scal) or x / (*scal) with the scalar on the device
void launch_scale_dev(const double* x, const double* scal, int invert, double* y, long long n, cudaStream_t st);
// blockwise transposes: dst block = scale[b] * (src block)^T ; tables give (src off, dst off, rows, cols, lds, ldd)
struct TrBlock {
  long long soff, doff;
  int rows, cols, lds, ldd;
  double scale;
};
void launch_transpose(const TrBlock* blocks, int nblocks, const double* src, double* dst, cudaStream_t st);
void launch_copy2d(const TrBlock* blocks, int nblocks, const double* src, double* dst, cudaStream_t st);
// one-sided Jacobi SVD of row panels: G [k x len] (ldg), Q [k x k] (ldq); k <= 1024
struct SvdPanel {
  long long offG, offQ, offS;
  int k, len, ldg, ldq, u_in_g, pad_;
};
// returns 0 ok, 1 sweeps did not converge, < 0 CUDA / allocation failure
int launch_svd(const SvdPanel* panels_dev, const SvdPanel* panels_host, int npanels, double* G, double* Q, double* G2,
               double* Q2, double* sig, cudaStream_t st);
// set blocks to 0 (mode 0) or to the unit matrix (mode 1): table entries (off, rows, cols, ld)
struct FillBlock {
  long long off;
  int rows, cols, ld, mode;
};
void launch_fill(const FillBlock* blocks, int nblocks, double* dst, cudaStream_t st);
void launch_diag_inv(const FillBlock* blocks, int nblocks, const double* src, double* dst, cudaStream_t st);
// in-place QR with positive diagonal (iterated classical Gram-Schmidt) of row-major [m x n] panels:
// panel p: Q overwrites A (off_a, m, n, lda); R (n x n upper, row-major ldr) written at off_r
struct QrPanel {
  long long off_a, off_r;
  int m, n, lda, ldr;
};
void launch_qr(const QrPanel* panels, int npanels, int max_m, double* A, double* R, int* status, cudaStream_t st);
int gemm_max_ctas_per_sm();
double probe_fp64(int which, int sm_count, cudaStream_t st);

#ifndef HTN_BK
#define HTN_BK 16
#endif
constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = HTN_BK;  // K extent of one staged chunk (16 or 32)

}  // namespace htn
