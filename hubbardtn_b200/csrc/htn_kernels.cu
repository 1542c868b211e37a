// sm_100a device kernels of the H_eff hot path.
//
//  * grouped_gemm_kernel : persistent grouped FP64 tensor-core GEMM with K-segments.  Every
//    work item is one <=64x64 output tile of one symmetry-sector block; its K loop runs over
//    a list of segments (A_seg, B_seg, K_seg, coef), i.e. over (MPO level, sector) pairs.
//    Replaces the per-coupled-sector BLAS `zgemm` calls TensorKit issues for `∂AC`
//    (MKL_jll 2025.0.1 zgemm, Manifest.toml:716).  FP64 has no
//    tcgen05/UMMA kind on sm_100a; the tensor path is DMMA.8x8x4 (mma.sync.m8n8k4.f64),
//    operands staged global -> shared with 16-byte cp.async in a 3-stage ring.
//  * mix_kernel : stage W, block-wise linear combinations U = sum coef * T with the SU(2)
//    recoupling coefficients (the O(D^2) part of the apply).
//  * pack / unpack / dot / axpby : arena <-> packed host layout and Krylov vector algebra
//    with the TensorKit inner product (weights = quantum dimension of the coupled sector).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "htn_internal.hpp"
#include "htn_stackl.cuh"

namespace htn {

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------
// grouped GEMM
// ------------------------------------------------------------------------------------
constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
constexpr int STAGES = 64 / BK;  // 64 k-columns (64 KB) in flight per CTA: 4 x 16 or 2 x 32
#ifndef HTN_NPROD
#define HTN_NPROD 1
#endif
constexpr int NCONS_WARPS = 4, NPROD_WARPS = HTN_NPROD;  // 2: one warp stages B, the other A (experiment)
constexpr int NPROD = NPROD_WARPS * 32;
constexpr int NTHREADS = (NCONS_WARPS + NPROD_WARPS) * 32;
// Shared-memory tiles are UNPADDED and XOR-swizzled in units of 4 doubles (32 B):
//   A chunk [64 rows][16 k]:  element (row, k)   at row * 16 + (k   ^ ((row  & 3) << 2))
//   B chunk [16 k][64 cols]:  element (krow, c)  at krow * 64 + (c   ^ ((krow & 3) << 2))
// so that the 16 lanes of a half warp (4 rows x 4 k, or 4 k x 4 cols) hit 16 distinct doubles of one
// 128-byte line for every DMMA fragment load, and the 16-byte cp.async stores stay aligned.  Without
// padding a stage is 16 KB, which lets 4 stages x 3 CTAs fit one SM (a 3-stage ring starved the
// consumers ~22 % of the time: profiles/r1b_*).
constexpr int LDAS = BK;
constexpr int LDBS = BN;
constexpr int A_STAGE = BM * LDAS;
constexpr int B_STAGE = BK * LDBS;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double) + 2 * STAGES * 8 + STAGES * 4 + 64;

// ---- mbarrier helpers (producer/consumer ring; CTA scope) ----------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// producer-side wait: back off between polls so that the (mostly idle) producer warps do not compete for issue
// slots with the DMMA warps of their SM sub-partition
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, unsigned parity) {
  unsigned ok;
  while (true) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(128);
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// Work distribution inside a CTA.  Warps 0-3 are CONSUMERS (DMMA only), warps 4-7 PRODUCERS
// (operand staging).  A tile is MA x NA DMMA atoms (8x8 each, <= 8x8 atoms).  One dimension is
// "fixed": it is cut into 4 strips of 2 atoms, one strip per consumer role; the other ("flex")
// dimension has 1..8 atoms and every consumer covers all of it.  All consumers therefore do the
// same work, block extents are padded only to the atom size 8, and the inner loop is
// straight-line code specialised on the flex extent (no predication).
//   layout 0 (LAY_A): flex = M (rows), fixed = N: role w owns col atoms 2w, 2w+1
//   layout 1 (LAY_B): flex = N (cols), fixed = M: role w owns row atoms 2w, 2w+1
// acc[flex atom][fixed atom f][2 values of the DMMA C fragment].
// per-lane fragment addressing (see the swizzle comment above)
struct FragAddr {
  int a_base;    // A: row part + t
  int koff[4];   // A: swizzled k offset of k4-step kk
  int b_even;    // B: krow t, swizzled column base for even 8-column atoms
  int b_odd;     //    ... for odd atoms
  int g3;        // A: row & 3 (swizzle key)
};

template <int FLEX, bool LAYB>
__device__ __forceinline__ void mma_k4(double (&acc)[8][2][2], const double* __restrict__ as,
                                       const double* __restrict__ bs, const FragAddr& fa, int kk) {
  double fx[FLEX], ff[2];
  const double* ap = as + fa.a_base + fa.koff[kk & 3] + (kk >> 2) * 16;
  const double* bp = bs + kk * 4 * LDBS;
  if (!LAYB) {
#pragma unroll
    for (int i = 0; i < FLEX; ++i) fx[i] = ap[i * 8 * LDAS];
    ff[0] = bp[fa.b_even];
    ff[1] = bp[fa.b_odd + 8];
#pragma unroll
    for (int i = 0; i < FLEX; ++i) {
      dmma884(acc[i][0][0], acc[i][0][1], fx[i], ff[0]);
      dmma884(acc[i][1][0], acc[i][1][1], fx[i], ff[1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < FLEX; ++j) fx[j] = bp[((j & 1) ? fa.b_odd : fa.b_even) + j * 8];
    ff[0] = ap[0];
    ff[1] = ap[8 * LDAS];
#pragma unroll
    for (int j = 0; j < FLEX; ++j) {
      dmma884(acc[j][0][0], acc[j][0][1], ff[0], fx[j]);
      dmma884(acc[j][1][0], acc[j][1][1], ff[1], fx[j]);
    }
  }
}

#ifndef HTN_KUNROLL
#define HTN_KUNROLL 4
#endif

// fragments of k4-step kk: fx = the FLEX atoms of the flex operand, ff = the two atoms of this role's strip
template <int FLEX, bool LAYB>
__device__ __forceinline__ void load_frags(double (&fx)[FLEX], double (&ff)[2], const double* __restrict__ as,
                                           const double* __restrict__ bs, const FragAddr& fa, int kk) {
  const double* ap = as + fa.a_base + (((kk ^ fa.g3) & 3) << 2) + (kk >> 2) * 16;
  const double* bp = bs + kk * 4 * LDBS;
  if (!LAYB) {
#pragma unroll
    for (int i = 0; i < FLEX; ++i) fx[i] = ap[i * 8 * LDAS];
    ff[0] = bp[fa.b_even];
    ff[1] = bp[fa.b_odd + 8];
  } else {
#pragma unroll
    for (int j = 0; j < FLEX; ++j) fx[j] = bp[((j & 1) ? fa.b_odd : fa.b_even) + j * 8];
    ff[0] = ap[0];
    ff[1] = ap[8 * LDAS];
  }
}
template <int FLEX, bool LAYB>
__device__ __forceinline__ void issue_dmmas(double (&acc)[8][2][2], const double (&fx)[FLEX], const double (&ff)[2]) {
#pragma unroll
  for (int i = 0; i < FLEX; ++i) {
    if (!LAYB) {
      dmma884(acc[i][0][0], acc[i][0][1], fx[i], ff[0]);
      dmma884(acc[i][1][0], acc[i][1][1], fx[i], ff[1]);
    } else {
      dmma884(acc[i][0][0], acc[i][0][1], ff[0], fx[i]);
      dmma884(acc[i][1][0], acc[i][1][1], ff[1], fx[i]);
    }
  }
}

template <int FLEX, bool LAYB>
__device__ __forceinline__ void mma_chunk(double (&acc)[8][2][2], const double* __restrict__ as,
                                          const double* __restrict__ bs, const FragAddr& fa, int nk4) {
#if HTN_KUNROLL == 4
  if (nk4 == BK / 4) {
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) mma_k4<FLEX, LAYB>(acc, as, bs, fa, kk);
  } else {  // K tail of a segment (a rolled loop here is smaller but measured 1 % slower)
    for (int kk = 0; kk < BK / 4; ++kk)
      if (kk < nk4) mma_k4<FLEX, LAYB>(acc, as, bs, fa, kk);
  }
#else
  // rolled K loop (small code: sixteen tile-shape specialisations share the instruction cache of an SM
  // whose three CTAs run different ones): two k4-steps per trip, the fragments of the next step are
  // loaded into the other register set before the DMMAs of the current one issue
#if HTN_KUNROLL == 1
  double fx[FLEX], ff[2];
#pragma unroll 1
  for (int kk = 0; kk < nk4; ++kk) {
    load_frags<FLEX, LAYB>(fx, ff, as, bs, fa, kk);
    issue_dmmas<FLEX, LAYB>(acc, fx, ff);
  }
  return;
#endif
  double fx0[FLEX], ff0[2], fx1[FLEX], ff1[2];
  load_frags<FLEX, LAYB>(fx0, ff0, as, bs, fa, 0);
#pragma unroll 1
  for (int kk = 0; kk < nk4; kk += 2) {
    if (kk + 1 < nk4) load_frags<FLEX, LAYB>(fx1, ff1, as, bs, fa, kk + 1);
    issue_dmmas<FLEX, LAYB>(acc, fx0, ff0);
    if (kk + 1 >= nk4) break;
    if (kk + 2 < nk4) load_frags<FLEX, LAYB>(fx0, ff0, as, bs, fa, kk + 2);
    issue_dmmas<FLEX, LAYB>(acc, fx1, ff1);
  }
#endif
}

struct Ring {
  double* As;
  double* Bs;
  uint64_t* full;
  uint64_t* empty;
  int* meta;
  int stage;
  unsigned phase;
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// Consumer side of one work item, fully specialised on the tile shape: waits for the staged
// chunks, issues the DMMAs, releases the stages, then stores the accumulators.
template <int FLEX, bool LAYB>
__device__ __forceinline__ void consume_item(const GemmItem& item, Ring& rg, int role, int lane, const Bases& bases,
                                             int dbg) {
  const int g = lane >> 2, t = lane & 3;
  const int mt = item.mt, nt = item.nt;
  const bool active = role * 16 < (LAYB ? mt : nt);  // warp-uniform: this strip holds data
  FragAddr fa;
  fa.a_base = (LAYB ? role * 16 * LDAS : 0) + g * LDAS + t;
  fa.g3 = g & 3;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) fa.koff[kk] = (kk ^ (g & 3)) << 2;
  {
    const int bsw = (t & 2) << 2, gsw = g ^ ((t & 1) << 2), cbase = LAYB ? 0 : role * 16;
    fa.b_even = t * LDBS + cbase + bsw + gsw;
    fa.b_odd = t * LDBS + cbase - bsw + gsw;
  }

  double acc[8][2][2];
#pragma unroll
  for (int i = 0; i < FLEX; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;

  const int nchunks = item.nchunks;
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    mbar_wait(&rg.full[rg.stage], rg.phase);
    if (active && !(dbg & 1)) {  // timing experiment HTN_GEMM_DEBUG=1: no DMMA
#ifdef HTN_FULLCHUNK
      // K tails are zero-filled by the producer: always run the full chunk (no meta read / branch)
      mma_chunk<FLEX, LAYB>(acc, rg.As + rg.stage * A_STAGE, rg.Bs + rg.stage * B_STAGE, fa, BK / 4);
#else
      const int nk4 = rg.meta[rg.stage];
      mma_chunk<FLEX, LAYB>(acc, rg.As + rg.stage * A_STAGE, rg.Bs + rg.stage * B_STAGE, fa, nk4);
#endif
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&rg.empty[rg.stage]);
    rg.advance();
  }

  if (active) {
    double* C = const_cast<double*>(resolve(item.c_off, item.c_base, bases));
    const long long ldc = item.ldc;
    // only the last flex atom and the strip's own edge can cross the tile extent: everything else is
    // stored unpredicated (keeps the sixteen specialised epilogues small)
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int fi = (role * 2 + f) * 8;
      if (!LAYB) {
        const int col = fi + 2 * t;
        if (col < nt) {
          const bool pair = col + 1 < nt;
          double* p = C + g * ldc + col;
#pragma unroll
          for (int x = 0; x < FLEX; ++x) {
            if (x < FLEX - 1 || x * 8 + g < mt) {
              if (pair)
                *reinterpret_cast<double2*>(p + x * 8 * ldc) = make_double2(acc[x][f][0], acc[x][f][1]);
              else
                p[x * 8 * ldc] = acc[x][f][0];
            }
          }
        }
      } else {
        const int row = fi + g;
        if (row < mt) {
          double* p = C + row * ldc + 2 * t;
#pragma unroll
          for (int x = 0; x < FLEX; ++x) {
            if (x < FLEX - 1 || x * 8 + 2 * t + 1 < nt)
              *reinterpret_cast<double2*>(p + x * 8) = make_double2(acc[x][f][0], acc[x][f][1]);
            else if (x * 8 + 2 * t < nt)
              p[x * 8] = acc[x][f][0];
          }
        }
      }
    }
  }
}

template <bool LAYB>
__device__ __forceinline__ void consume_dispatch(int flex, const GemmItem& item, Ring& rg, int role, int lane,
                                                 const Bases& bases, int dbg) {
  switch (flex) {
    case 1: consume_item<1, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 2: consume_item<2, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 3: consume_item<3, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 4: consume_item<4, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 5: consume_item<5, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 6: consume_item<6, LAYB>(item, rg, role, lane, bases, dbg); break;
    case 7: consume_item<7, LAYB>(item, rg, role, lane, bases, dbg); break;
    default: consume_item<8, LAYB>(item, rg, role, lane, bases, dbg); break;
  }
}

__global__ void __launch_bounds__(NTHREADS, NPROD_WARPS == 1 ? 3 : 2)
grouped_gemm_kernel(const GemmItem* __restrict__ items, const GemmSeg* __restrict__ segs, int nitems,
                    const __grid_constant__ Bases bases, int dbg) {
  extern __shared__ __align__(16) double smem[];
  Ring rg;
  rg.As = smem;
  rg.Bs = smem + STAGES * A_STAGE;
  rg.full = reinterpret_cast<uint64_t*>(rg.Bs + STAGES * B_STAGE);
  rg.empty = rg.full + STAGES;
  rg.meta = reinterpret_cast<int*>(rg.empty + STAGES);  // nk4 of the chunk held by each stage
  rg.stage = 0;
  rg.phase = 0;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&rg.full[s], NPROD + 1);     // every producer lane: one cp.async (noinc) arrive; lane 0: one plain arrive
      mbar_init(&rg.empty[s], NCONS_WARPS);  // one elected lane per consumer warp
    }
  }
  __syncthreads();
  if (blockIdx.x >= nitems) return;
  // CTA b walks items b, b + G, b + 2G, ...: the host lays the table out as one list per CTA (balanced by a cost
  // model, ordered so that the CTAs sharing an SM are out of phase; htn_program.cpp:balance_items).
  const int n_mine = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_index = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };

  if (warp >= NCONS_WARPS) {
    // =========================== PRODUCER (one warp) ================================
    int it = item_index(0);
    GemmItem item = items[it];
    for (int j = 0; j < n_mine; ++j) {
      const int itn = j + 1 < n_mine ? item_index(j + 1) : nitems;
      GemmItem next_item = item;
      if (itn < nitems) next_item = items[itn];  // prefetch: consumed at the next iteration
      const int mt = item.mt, nt = item.nt;
      GemmSeg sg{};
      if (item.seg_begin < item.seg_end) sg = segs[item.seg_begin];
      // (padding items of the balanced schedule have mt = 0 and no segments: nothing to load)
      for (int si = item.seg_begin; si < item.seg_end; ++si) {
        GemmSeg sg_next = sg;
        if (si + 1 < item.seg_end) sg_next = segs[si + 1];
        const int K = sg.K;
        const double* Bg = resolve(sg.b_off, sg.b_base, bases);
        const double* Ag = resolve(sg.a_off, sg.a_base, bases);
        for (int k0 = 0; k0 < K; k0 += BK) {
          mbar_wait_backoff(&rg.empty[rg.stage], rg.phase ^ 1u);
          double* as = rg.As + rg.stage * A_STAGE;
          double* bs = rg.Bs + rg.stage * B_STAGE;
          if (!(dbg & 2)) {  // timing experiment HTN_GEMM_DEBUG=2: no operand loads
          // ---- B operand: 16 x 64 chunk: lane = 16-byte segment of a row, q = row ---------
          // (columns >= nt feed only discarded outputs: loaded only up to the tile extent)
          if (NPROD_WARPS == 1 || warp == NCONS_WARPS) {
            int bytes_row = (nt - lane * 2) * 8;
            bytes_row = bytes_row < 0 ? 0 : (bytes_row > 16 ? 16 : bytes_row);
            const char* src = reinterpret_cast<const char*>(Bg + (long long)k0 * sg.ldb + lane * 2);
            const long long step = (long long)sg.ldb * 8;
            // swizzled destination of this lane's 16-byte segment in k-row q: (2 lane) ^ ((q & 3) << 2)
            double* dst0 = bs + ((lane * 2) ^ 0);
            double* dst1 = bs + ((lane * 2) ^ 4);
            double* dst2 = bs + ((lane * 2) ^ 8);
            double* dst3 = bs + ((lane * 2) ^ 12);
#define HTN_BDST(q) (((q) & 3) == 0 ? dst0 : ((q) & 3) == 1 ? dst1 : ((q) & 3) == 2 ? dst2 : dst3) + (q) * LDBS
            if (k0 + BK <= K) {
              if (bytes_row == 16) {
#pragma unroll
                for (int q = 0; q < BK; ++q) cp_async16(HTN_BDST(q), src + q * step, 16);
              } else if (bytes_row > 0) {
#pragma unroll
                for (int q = 0; q < BK; ++q) cp_async16(HTN_BDST(q), src + q * step, bytes_row);
              }
            } else {  // K tail: rows >= K must be zero (they meet the zero-filled A columns)
#pragma unroll
              for (int q = 0; q < BK; ++q) {
                const int bytes = (k0 + q < K) ? bytes_row : 0;
                cp_async16(HTN_BDST(q), bytes ? src + q * step : reinterpret_cast<const char*>(Bg), bytes);
              }
            }
#undef HTN_BDST
          }
          constexpr int ASEGS = BK / 2, ARPP = 32 / ASEGS;  // 16-byte segments per A row, rows per pass
          const int arow = lane / ASEGS, aseg = lane % ASEGS;
          if (NPROD_WARPS == 1 || warp == NCONS_WARPS + 1) {
          // ---- A operand straight from one array; rows >= mt feed only discarded outputs ----
          int bytes_k = (K - (k0 + aseg * 2)) * 8;
          bytes_k = bytes_k < 0 ? 0 : (bytes_k > 16 ? 16 : bytes_k);
          const char* src = reinterpret_cast<const char*>(Ag + (long long)arow * sg.lda + k0 + aseg * 2);
          const long long step = (long long)sg.lda * 8 * ARPP;
          // row = q * ARPP + arow; swizzle (row & 3) << 2 applied to the low 16 k of the segment position
          double* dstv[4 / ARPP > 0 ? 4 / ARPP : 1];
#pragma unroll
          for (int v = 0; v < 4 / ARPP; ++v)
            dstv[v] = as + arow * LDAS + (((aseg * 2) & ~15) | (((aseg * 2) & 15) ^ ((((v * ARPP) + arow) & 3) << 2)));
          const int nq = (mt - arow + ARPP - 1) / ARPP;  // rows q*ARPP+arow < mt
          if (nq == BM / ARPP) {
#pragma unroll
            for (int q = 0; q < BM / ARPP; ++q)
              cp_async16(dstv[q % (4 / ARPP)] + q * ARPP * LDAS, bytes_k ? src + q * step : reinterpret_cast<const char*>(Ag), bytes_k);
          } else {
#pragma unroll
            for (int q = 0; q < BM / ARPP; ++q)
              if (q < nq)
                cp_async16(dstv[q % (4 / ARPP)] + q * ARPP * LDAS, bytes_k ? src + q * step : reinterpret_cast<const char*>(Ag), bytes_k);
          }
          }
          }
          cp_async_mbar_arrive(&rg.full[rg.stage]);   // fires when this lane's copies have landed
          if (lane == 0 && warp == NCONS_WARPS) {
            int krem = K - k0;
            rg.meta[rg.stage] = krem >= BK ? BK / 4 : (krem + 3) >> 2;
            mbar_arrive(&rg.full[rg.stage]);          // releases the meta word
          }
          rg.advance();
        }
        sg = sg_next;
      }
      it = itn;
      item = next_item;
    }
  } else {
    // =========================== CONSUMERS (four warps) =============================
    int it = item_index(0);
    GemmItem item = items[it];
    for (int j = 0; j < n_mine; ++j) {
      const int itn = j + 1 < n_mine ? item_index(j + 1) : nitems;
      GemmItem next_item = item;
      if (itn < nitems) next_item = items[itn];
      // rotate the strip a warp owns from item to item so that partially filled strips do not
      // always land on the same SM sub-partition
      const int role = (warp + it) & 3;
      if (item.mt == 0) {
        // padding item of the balanced schedule
      } else if (item.layout != 0)
        consume_dispatch<true>((item.nt + 7) >> 3, item, rg, role, lane, bases, dbg);
      else
        consume_dispatch<false>((item.mt + 7) >> 3, item, rg, role, lane, bases, dbg);
      it = itn;
      item = next_item;
    }
  }
}

int gemm_max_ctas_per_sm() {
  static int cached = -1;
  if (cached >= 0) return cached;
  cudaFuncSetAttribute(grouped_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, grouped_gemm_kernel, NTHREADS, GEMM_SMEM_BYTES);
  cached = n > 0 ? n : 1;
  return cached;
}

void launch_gemm(const GemmItem* items, const GemmSeg* segs, int nitems, const Bases& bases, int grid,
                 cudaStream_t st) {
  if (nitems <= 0) return;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(grouped_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    attr = true;
  }
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("HTN_GEMM_DEBUG");  // timing experiments only (results are wrong when set)
    dbg = e ? atoi(e) : 0;
  }
  grouped_gemm_kernel<<<grid, NTHREADS, GEMM_SMEM_BYTES, st>>>(items, segs, nitems, bases, dbg);
}

// ------------------------------------------------------------------------------------
// stacked stage L with the mix riding along (htn_stackl.cuh)
// ------------------------------------------------------------------------------------
// consumer groups per CTA of the stacked kernel: 1 (two CTAs per SM) or 2 (one CTA per SM, shared slab); HTN_STACK_NG
int stack_gemm_groups() {
  static int ng = -1;
  if (ng < 0) {
    const char* e = getenv("HTN_STACK_NG");
    ng = e ? atoi(e) : 1;
    if (ng != 2) ng = 1;
  }
  return ng;
}

int stack_gemm_cons_warps() { return stack_gemm_groups() == 2 ? SlCfg<2>::NCONS : SlCfg<1>::NCONS; }

template <int NG>
static int stack_occupancy() {
  cudaFuncSetAttribute(stack_gemm_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, SlCfg<NG>::SMEM_BYTES);
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, stack_gemm_kernel<NG>, SlCfg<NG>::THREADS, SlCfg<NG>::SMEM_BYTES);
  return n > 0 ? n : 1;
}

int stack_gemm_ctas_per_sm() {
  static int cached = -1;
  if (cached >= 0) return cached;
  cached = stack_gemm_groups() == 2 ? stack_occupancy<2>() : stack_occupancy<1>();
  return cached;
}

void launch_stack_gemm(const StackArgs& args, const Bases& bases, int grid, cudaStream_t st) {
  stack_gemm_ctas_per_sm();  // sets the shared-memory attribute of the chosen instantiation once
  if (stack_gemm_groups() == 2)
    stack_gemm_kernel<2><<<grid, SlCfg<2>::THREADS, SlCfg<2>::SMEM_BYTES, st>>>(args, bases);
  else
    stack_gemm_kernel<1><<<grid, SlCfg<1>::THREADS, SlCfg<1>::SMEM_BYTES, st>>>(args, bases);
}

// ------------------------------------------------------------------------------------
// final mix: y = sum coef * (T block | x block | split-K partial).  One CTA per chunk of a
// target; chunk sizes shrink with the number of sources so that every CTA streams a similar
// number of bytes; source descriptors are staged in shared memory; loads are unrolled x4.
// ------------------------------------------------------------------------------------
// Occupancy is what this kernel lives on: at 42 registers (5 CTAs/SM) it streamed 4.1 TB/s, capped at 32 registers
// (8 CTAs/SM = 64 warps) 4.9 TB/s; a grid of 32 CTAs per SM walking the chunk list adds another 2 %.
constexpr int MIX_MAX_SRC_SMEM = 128;

__global__ void __launch_bounds__(256, 8) mix_kernel(const MixTarget* __restrict__ tg, const MixSrc* __restrict__ src,
                                                  const MixChunk* __restrict__ chunks, int nchunks,
                                                  const __grid_constant__ Bases bases) {
  __shared__ const double* sptr[MIX_MAX_SRC_SMEM];
  __shared__ double scoef[MIX_MAX_SRC_SMEM];
  for (int cidx = blockIdx.x; cidx < nchunks; cidx += gridDim.x) {
  const MixChunk ch = chunks[cidx];
  const MixTarget T = tg[ch.target];
  double* dst = const_cast<double*>(resolve(T.off, T.base, bases));
  const int nsrc = T.src_end - T.src_begin;
  const int end = ch.elem0 + ch.nelem;
  for (int s0 = 0; s0 < nsrc || s0 == 0; s0 += MIX_MAX_SRC_SMEM) {
    const int ns = min(MIX_MAX_SRC_SMEM, nsrc - s0);
    __syncthreads();
    for (int s = threadIdx.x; s < ns; s += blockDim.x) {
      const MixSrc S = src[T.src_begin + s0 + s];
      sptr[s] = resolve(S.off, S.base, bases);
      scoef[s] = S.coef;
    }
    __syncthreads();
    // all blocks have even ld and 16-byte aligned offsets: work in double2
    for (int e = ch.elem0 + 2 * threadIdx.x; e < end; e += 2 * blockDim.x) {
      double2 acc = (s0 == 0) ? make_double2(0.0, 0.0) : *reinterpret_cast<double2*>(dst + e);
      int s = 0;
      for (; s + 4 <= ns; s += 4) {
        const double2 v0 = __ldg(reinterpret_cast<const double2*>(sptr[s] + e));
        const double2 v1 = __ldg(reinterpret_cast<const double2*>(sptr[s + 1] + e));
        const double2 v2 = __ldg(reinterpret_cast<const double2*>(sptr[s + 2] + e));
        const double2 v3 = __ldg(reinterpret_cast<const double2*>(sptr[s + 3] + e));
        acc.x = fma(scoef[s], v0.x, acc.x);
        acc.y = fma(scoef[s], v0.y, acc.y);
        acc.x = fma(scoef[s + 1], v1.x, acc.x);
        acc.y = fma(scoef[s + 1], v1.y, acc.y);
        acc.x = fma(scoef[s + 2], v2.x, acc.x);
        acc.y = fma(scoef[s + 2], v2.y, acc.y);
        acc.x = fma(scoef[s + 3], v3.x, acc.x);
        acc.y = fma(scoef[s + 3], v3.y, acc.y);
      }
      for (; s < ns; ++s) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(sptr[s] + e));
        acc.x = fma(scoef[s], v.x, acc.x);
        acc.y = fma(scoef[s], v.y, acc.y);
      }
      *reinterpret_cast<double2*>(dst + e) = acc;
    }
    if (nsrc == 0) break;
  }
  }
}

void launch_mix(const MixTarget* tg, const MixSrc* src, const MixChunk* chunks, int nchunks, const Bases& bases,
                cudaStream_t st) {
  if (nchunks <= 0) return;
  static int grid_cap = 0;
  if (grid_cap == 0) {
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("HTN_MIX_CTAS_PER_SM");  // 0 = one CTA per chunk
    const int per = e ? atoi(e) : 32;
    grid_cap = per > 0 ? per * nsm : 1 << 30;
  }
  mix_kernel<<<std::min(nchunks, grid_cap), 256, 0, st>>>(tg, src, chunks, nchunks, bases);
}

// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// pack / unpack (packed host layout <-> padded arena), dot, axpby
// chunk table: int3-like triples (block, row0, nrows)
// ------------------------------------------------------------------------------------
__global__ void pack_kernel(const DevBlock* __restrict__ blocks, const int* __restrict__ chunks,
                            const double* __restrict__ packed, double* __restrict__ padded) {
  const int b = chunks[3 * blockIdx.x], r0 = chunks[3 * blockIdx.x + 1], nr = chunks[3 * blockIdx.x + 2];
  const DevBlock B = blocks[b];
  const int n = nr * B.cols;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int r = r0 + e / B.cols, c = e % B.cols;
    padded[B.off + (long long)r * B.ld + c] = packed[B.hoff + (long long)r * B.cols + c];
  }
}

__global__ void unpack_kernel(const DevBlock* __restrict__ blocks, const int* __restrict__ chunks,
                              const double* __restrict__ padded, double* __restrict__ packed) {
  const int b = chunks[3 * blockIdx.x], r0 = chunks[3 * blockIdx.x + 1], nr = chunks[3 * blockIdx.x + 2];
  const DevBlock B = blocks[b];
  const int n = nr * B.cols;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    int r = r0 + e / B.cols, c = e % B.cols;
    packed[B.hoff + (long long)r * B.cols + c] = padded[B.off + (long long)r * B.ld + c];
  }
}

void launch_pack(const DevBlock* blocks, const int* chunks, int nchunks, const double* packed, double* padded,
                 cudaStream_t st) {
  if (nchunks > 0) pack_kernel<<<nchunks, 256, 0, st>>>(blocks, chunks, packed, padded);
}
void launch_unpack(const DevBlock* blocks, const int* chunks, int nchunks, const double* padded, double* packed,
                   cudaStream_t st) {
  if (nchunks > 0) unpack_kernel<<<nchunks, 256, 0, st>>>(blocks, chunks, padded, packed);
}

// deterministic two-pass weighted dot: partial[chunk] then a single-CTA ordered sum
__global__ void dot_partial_kernel(const DevBlock* __restrict__ blocks, const int* __restrict__ chunks,
                                   const double* __restrict__ x, const double* __restrict__ y,
                                   double* __restrict__ partial) {
  const int b = chunks[3 * blockIdx.x], r0 = chunks[3 * blockIdx.x + 1], nr = chunks[3 * blockIdx.x + 2];
  const DevBlock B = blocks[b];
  // rows are padded with zeros, so the flat padded range can be reduced directly
  const long long base = B.off + (long long)r0 * B.ld;
  const int n = nr * B.ld;
  double acc = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) acc = fma(x[base + e], y[base + e], acc);
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += sh[w];
    partial[blockIdx.x] = s * B.weight;
  }
}

__global__ void dot_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) acc += partial[e];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

void launch_dot(const DevBlock* blocks, const int* chunks, int nchunks, const double* x, const double* y,
                double* partial, double* out, cudaStream_t st) {
  dot_partial_kernel<<<nchunks, 256, 0, st>>>(blocks, chunks, x, y, partial);
  dot_final_kernel<<<1, 256, 0, st>>>(partial, nchunks, out);
}

__global__ void axpby_kernel(double alpha, const double* __restrict__ x, double beta, double* __restrict__ y,
                             long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = (beta == 0.0) ? alpha * x[i] : fma(alpha, x[i], beta * y[i]);
}

void launch_axpby(double alpha, const double* x, double beta, double* y, long long n, cudaStream_t st) {
  if (n <= 0) return;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  axpby_kernel<<<grid, 256, 0, st>>>(alpha, x, beta, y, n);
}

// ------------------------------------------------------------------------------------
// Krylov vector kernels (KrylovKit orthogonalisation): all inner products of
// one vector against the whole basis in ONE pass over the vector (classical Gram-Schmidt step),
// the matching rank-k update, and scaling by a device-resident scalar -- no host round trip
// between them.  Deterministic: per-chunk partials, then an ordered tree sum.
// ------------------------------------------------------------------------------------
constexpr int MD_MAXVEC = 64;

__global__ void __launch_bounds__(256)
multidot_partial_kernel(const DevBlock* __restrict__ blocks, const int* __restrict__ chunks,
                        const double* __restrict__ V, long long stride, int nvec, const double* __restrict__ w,
                        double* __restrict__ partial) {
  const int b = chunks[3 * blockIdx.x], r0 = chunks[3 * blockIdx.x + 1], nr = chunks[3 * blockIdx.x + 2];
  const DevBlock B = blocks[b];
  const long long base = B.off + (long long)r0 * B.ld;
  const int n = nr * B.ld;  // pads are zero in every vector
  __shared__ double sh[MD_MAXVEC][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // eight basis vectors at a time: every element of w is read once per group and meets eight
  // independent loads (memory-level parallelism), instead of one dependent pass per vector
  for (int g0 = 0; g0 < nvec; g0 += 8) {
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    const int ng = min(8, nvec - g0);
    const double* vg = V + (long long)g0 * stride + base;
    if (ng == 8) {
      for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double wv = w[base + e];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fma(vg[(long long)q * stride + e], wv, acc[q]);
      }
    } else {
      for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const double wv = w[base + e];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q < ng) acc[q] = fma(vg[(long long)q * stride + e], wv, acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double a = acc[q];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
      if (lane == 0 && q < ng) sh[g0 + q][warp] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x < nvec) {
    double s = 0.0;
    for (int wv = 0; wv < 8; ++wv) s += sh[threadIdx.x][wv];
    partial[(long long)blockIdx.x * nvec + threadIdx.x] = s * B.weight;
  }
}

__global__ void __launch_bounds__(256)
multidot_final_kernel(const double* __restrict__ partial, int nchunks, int nvec, double* __restrict__ out) {
  __shared__ double sh[256];
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int c = threadIdx.x; c < nchunks; c += blockDim.x) acc += partial[(long long)c * nvec + j];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[j] = sh[0];
}

void launch_multidot(const DevBlock* blocks, const int* chunks, int nchunks, const double* V, long long stride,
                     int nvec, const double* w, double* partial, double* out, cudaStream_t st) {
  if (nvec <= 0) return;
  if (nchunks <= 0) {
    cudaMemsetAsync(out, 0, nvec * sizeof(double), st);
    return;
  }
  multidot_partial_kernel<<<nchunks, 256, 0, st>>>(blocks, chunks, V, stride, nvec, w, partial);
  multidot_final_kernel<<<nvec, 256, 0, st>>>(partial, nchunks, nvec, out);
}

__global__ void __launch_bounds__(256)
multiaxpy_kernel(const double* __restrict__ V, long long stride, int nvec, const double* __restrict__ h, double sign,
                 double* __restrict__ w, long long n) {
  __shared__ double hs[MD_MAXVEC];
  if (threadIdx.x < nvec) hs[threadIdx.x] = sign * h[threadIdx.x];
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gs = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += gs) {
    double acc = w[i];
    for (int j = 0; j < nvec; ++j) acc = fma(hs[j], V[(long long)j * stride + i], acc);
    w[i] = acc;
  }
}

void launch_multiaxpy(const double* V, long long stride, int nvec, const double* h, double sign, double* w,
                      long long n, cudaStream_t st) {
  if (n <= 0 || nvec <= 0) return;
  int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  multiaxpy_kernel<<<grid, 256, 0, st>>>(V, stride, nvec, h, sign, w, n);
}

// mode 0: y = x * s ; 1: y = x / s ; 2: y = x / sqrt(s) ; 3: y = x / sqrt(s), or 0 when s < 1e-28
__global__ void scale_dev_kernel(const double* __restrict__ x, const double* __restrict__ scal, int mode,
                                 double* __restrict__ y, long long n) {
  double s = *scal;
  if (mode == 1) s = 1.0 / s;
  if (mode == 2) s = 1.0 / sqrt(s);
  if (mode == 3) s = s < 1e-28 ? 0.0 : 1.0 / sqrt(s);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gs = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += gs) y[i] = x[i] * s;
}

void launch_scale_dev(const double* x, const double* scal, int mode, double* y, long long n, cudaStream_t st) {
  if (n <= 0) return;
  int grid = (int)std::min<long long>((n + 255) / 256, 148 * 8);
  scale_dev_kernel<<<grid, 256, 0, st>>>(x, scal, mode, y, n);
}

// ------------------------------------------------------------------------------------
// blockwise transpose with a per-block scale (tiles of <= 32 x 32 listed by the host)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const TrBlock* __restrict__ tiles, const double* __restrict__ src,
                                                        double* __restrict__ dst) {
  __shared__ double t[32][33];
  const TrBlock T = tiles[blockIdx.x];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < T.rows; r += 8)
    if (tx < T.cols) t[r][tx] = src[T.soff + (long long)r * T.lds + tx];
  __syncthreads();
  for (int c = ty; c < T.cols; c += 8)
    if (tx < T.rows) dst[T.doff + (long long)c * T.ldd + tx] = T.scale * t[tx][c];
}

void launch_transpose(const TrBlock* tiles, int ntiles, const double* src, double* dst, cudaStream_t st) {
  if (ntiles > 0) transpose_kernel<<<ntiles, 256, 0, st>>>(tiles, src, dst);
}

__global__ void __launch_bounds__(256) fill_kernel(const FillBlock* __restrict__ blocks, double* __restrict__ dst) {
  const FillBlock B = blocks[blockIdx.x];
  const int n = B.rows * B.ld;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int r = e / B.ld, c = e % B.ld;
    dst[B.off + e] = (B.mode == 1 && r == c && c < B.cols) ? 1.0 : 0.0;
  }
}

// dst = inverse of the diagonal of src, off-diagonal zero (IDMRG2 edge update: inv(C) with C diagonal)
__global__ void __launch_bounds__(256) diag_inv_kernel(const FillBlock* __restrict__ blocks, const double* __restrict__ src,
                                                       double* __restrict__ dst) {
  const FillBlock B = blocks[blockIdx.x];
  const int n = B.rows * B.ld;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int r = e / B.ld, c = e % B.ld;
    dst[B.off + e] = (r == c && c < B.cols) ? 1.0 / src[B.off + e] : 0.0;
  }
}

void launch_diag_inv(const FillBlock* blocks, int nblocks, const double* src, double* dst, cudaStream_t st) {
  if (nblocks > 0) diag_inv_kernel<<<nblocks, 256, 0, st>>>(blocks, src, dst);
}

void launch_fill(const FillBlock* blocks, int nblocks, double* dst, cudaStream_t st) {
  if (nblocks > 0) fill_kernel<<<nblocks, 256, 0, st>>>(blocks, dst);
}

// ------------------------------------------------------------------------------------
// Positive QR of tall row-major panels, one CTA per panel (replaces LAPACK geqrf/orgqr behind
// TensorKit `leftorth!(.., QRpos())`).  Column-by-column classical
// Gram-Schmidt with one re-orthogonalisation pass (CGS2): orthogonality at machine precision for
// numerically full-rank panels, and diag(R) > 0 by construction.  Q overwrites A.
// Thread layout: 32 x 16; tx runs over already-orthogonalised columns (coalesced along a row).
// ------------------------------------------------------------------------------------
constexpr int QR_TY = 16;

__global__ void __launch_bounds__(32 * QR_TY) qr_cgs2_kernel(const QrPanel* __restrict__ panels, double* __restrict__ Abase,
                                                             double* __restrict__ Rbase, int* __restrict__ status) {
  const QrPanel P = panels[blockIdx.x];
  double* A = Abase + P.off_a;
  double* R = Rbase + P.off_r;
  const int m = P.m, n = P.n, lda = P.lda, ldr = P.ldr;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  __shared__ double red[QR_TY][33];
  __shared__ double hs[512];  // projection coefficients of the current column (n <= 512)
  __shared__ double snorm;
  for (int e = threadIdx.x; e < n * ldr; e += blockDim.x) R[e] = 0.0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    for (int pass = 0; pass < 2; ++pass) {
      // h[c] = sum_i Q[i][c] * v[i],  c < j
      for (int c0 = 0; c0 < j; c0 += 32) {
        const int c = c0 + tx;
        double acc = 0.0;
        if (c < j)
          for (int i = ty; i < m; i += QR_TY) acc = fma(A[(long long)i * lda + c], A[(long long)i * lda + j], acc);
        red[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && c < j) {
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < QR_TY; ++q) s += red[q][tx];
          hs[c] = s;
        }
        __syncthreads();
      }
      // v[i] -= sum_c Q[i][c] h[c]   (one warp per row)
      for (int i = ty; i < m; i += QR_TY) {
        double acc = 0.0;
        for (int c = tx; c < j; c += 32) acc = fma(A[(long long)i * lda + c], hs[c], acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (tx == 0) A[(long long)i * lda + j] -= acc;
      }
      for (int c = threadIdx.x; c < j; c += blockDim.x) R[(long long)c * ldr + j] += hs[c];
      __syncthreads();
    }
    // norm of the remainder
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const double v = A[(long long)i * lda + j];
      acc = fma(v, v, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (tx == 0) red[ty][0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int q = 0; q < QR_TY; ++q) s += red[q][0];
      snorm = sqrt(s);
      R[(long long)j * ldr + j] = snorm;
      if (!(snorm > 0.0)) atomicOr(status, 1);
    }
    __syncthreads();
    const double inv = snorm > 0.0 ? 1.0 / snorm : 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) A[(long long)i * lda + j] *= inv;
    __syncthreads();
  }
}

// Blocked variant (BCGS2 with the current column block resident in shared memory): per block J of QB columns
//   P = A[:, J] -> smem; then TWICE { project P against the finished Q (32-column chunks: H = Q_c^T P, P -= Q_c H);
//   column-by-column CGS2 inside the block (all in smem) };  P -> A[:, J].
// The second round acts on an already orthonormal block, which restores orthogonality against the finished
// columns to machine precision even when the block itself is ill-conditioned (a single round loses
// eps * cond(block): seen as 1e-8 at cond ~ 1e8).  R = [H1 + H2 R1 ; R2 R1].  Same result as qr_cgs2_kernel (the
// thin QR with diag(R) > 0 is unique) with ~20x fewer block-wide barriers and 16 FMAs per global load.
constexpr int QB = 16;

__global__ void __launch_bounds__(512) qr_bcgs2_kernel(const QrPanel* __restrict__ panels, double* __restrict__ Abase,
                                                       double* __restrict__ Rbase, int* __restrict__ status) {
  extern __shared__ __align__(16) double qsm[];
  const QrPanel Pn = panels[blockIdx.x];
  double* A = Abase + Pn.off_a;
  double* R = Rbase + Pn.off_r;
  const int m = Pn.m, n = Pn.n, lda = Pn.lda, ldr = Pn.ldr;
  double* P = qsm;                       // [m][QB]
  double* red = P + (size_t)m * QB;      // [16 warps][QB][32 lanes]  (also reused as [32][QB] partials)
  double* Hc = red + 16 * QB * 32;       // [32][QB]
  double* hs = Hc + 32 * QB;             // [QB] + scalar (+ padding to 32)
  double* Rb1 = hs + 32;                 // [QB][QB] intra-block factor of the first round
  double* Rcur = Rb1 + QB * QB;          // [QB][QB] intra-block factor of the current round
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < n * ldr; e += blockDim.x) R[e] = 0.0;
  __syncthreads();
  for (int j0 = 0; j0 < n; j0 += QB) {
    const int nb = min(QB, n - j0);
    for (int idx = tid; idx < m * QB; idx += blockDim.x) {
      const int i = idx / QB, b = idx % QB;
      P[idx] = b < nb ? A[(long long)i * lda + j0 + b] : 0.0;
    }
    __syncthreads();
    for (int rep = 0; rep < 2; ++rep) {
      // ---- projection against the finished columns ----
      for (int c0 = 0; c0 < j0; c0 += 32) {
        const int c = c0 + lane;
        const bool valid = c < j0;
        double acc[QB];
#pragma unroll
        for (int b = 0; b < QB; ++b) acc[b] = 0.0;
        for (int i = warp; i < m; i += 16) {
          const double q = valid ? A[(long long)i * lda + c] : 0.0;
          const double* pr = P + (size_t)i * QB;
#pragma unroll
          for (int b = 0; b < QB; ++b) acc[b] = fma(q, pr[b], acc[b]);
        }
#pragma unroll
        for (int b = 0; b < QB; ++b) red[(warp * QB + b) * 32 + lane] = acc[b];
        __syncthreads();
        {  // sum over the 16 row groups: thread -> (b, l)
          const int b = tid >> 5, l = tid & 31;
          double sacc = 0.0;
#pragma unroll
          for (int w = 0; w < 16; ++w) sacc += red[(w * QB + b) * 32 + l];
          Hc[l * QB + b] = sacc;
        }
        __syncthreads();
        {  // R[prev, J] += H (first round) or H R1 (second round)
          const int b = tid >> 5, l = tid & 31;
          if (c0 + l < j0 && b < nb) {
            double v;
            if (rep == 0) {
              v = Hc[l * QB + b];
            } else {
              v = 0.0;
              for (int k = 0; k <= b; ++k) v = fma(Hc[l * QB + k], Rb1[k * QB + b], v);
            }
            R[(long long)(c0 + l) * ldr + j0 + b] += v;
          }
        }
        {  // P -= Q_chunk H
          const int b = tid % QB, ig = tid / QB;  // 32 row groups
          const int nc = min(32, j0 - c0);
          for (int i = ig; i < m; i += 32) {
            const double* qr = A + (long long)i * lda + c0;
            double sacc = 0.0;
            for (int l = 0; l < nc; ++l) sacc = fma(qr[l], Hc[l * QB + b], sacc);
            P[(size_t)i * QB + b] -= sacc;
          }
        }
        __syncthreads();
      }
      // ---- CGS2 inside the block (factor Rcur, upper triangular) ----
      for (int e = tid; e < QB * QB; e += blockDim.x) Rcur[e] = 0.0;
      __syncthreads();
      for (int jb = 0; jb < nb; ++jb) {
        for (int pass = 0; pass < 2; ++pass) {
          if (jb > 0) {
            const int cb = tid % QB, ig = tid / QB;
            double sacc = 0.0;
            if (cb < jb)
              for (int i = ig; i < m; i += 32) sacc = fma(P[(size_t)i * QB + cb], P[(size_t)i * QB + jb], sacc);
            red[ig * QB + cb] = sacc;
            __syncthreads();
            if (tid < jb) {
              double h = 0.0;
              for (int g = 0; g < 32; ++g) h += red[g * QB + tid];
              hs[tid] = h;
              Rcur[tid * QB + jb] += h;
            }
            __syncthreads();
            for (int i = tid; i < m; i += blockDim.x) {
              double sacc2 = 0.0;
              for (int cbb = 0; cbb < jb; ++cbb) sacc2 = fma(P[(size_t)i * QB + cbb], hs[cbb], sacc2);
              P[(size_t)i * QB + jb] -= sacc2;
            }
            __syncthreads();
          }
        }
        double sacc = 0.0;
        for (int i = tid; i < m; i += blockDim.x) {
          const double v = P[(size_t)i * QB + jb];
          sacc = fma(v, v, sacc);
        }
        for (int o = 16; o > 0; o >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, o);
        if (lane == 0) red[warp] = sacc;
        __syncthreads();
        if (tid == 0) {
          double tot = 0.0;
          for (int w = 0; w < 16; ++w) tot += red[w];
          const double nrm = sqrt(tot);
          hs[QB] = nrm;
          Rcur[jb * QB + jb] = nrm;
          if (!(nrm > 0.0)) atomicOr(status, 1);
        }
        __syncthreads();
        const double inv = hs[QB] > 0.0 ? 1.0 / hs[QB] : 0.0;
        for (int i = tid; i < m; i += blockDim.x) P[(size_t)i * QB + jb] *= inv;
        __syncthreads();
      }
      if (rep == 0) {
        for (int e = tid; e < QB * QB; e += blockDim.x) Rb1[e] = Rcur[e];
      } else if (tid < QB * QB) {  // R[J, J] = R2 R1
        const int r = tid / QB, cc = tid % QB;
        if (r < nb && cc < nb && r <= cc) {
          double v = 0.0;
          for (int k = r; k <= cc; ++k) v = fma(Rcur[r * QB + k], Rb1[k * QB + cc], v);
          R[(long long)(j0 + r) * ldr + j0 + cc] = v;
        }
      }
      __syncthreads();
    }
    for (int idx = tid; idx < m * QB; idx += blockDim.x) {
      const int i = idx / QB, b = idx % QB;
      if (b < nb) A[(long long)i * lda + j0 + b] = P[idx];
    }
    __syncthreads();
  }
}

// max_m: largest panel height (host knows it); shared-memory need of the blocked kernel below
void launch_qr(const QrPanel* panels, int npanels, int max_m, double* A, double* R, int* status, cudaStream_t st) {
  if (npanels <= 0) return;
  const size_t need = ((size_t)max_m * QB + 16 * QB * 32 + 32 * QB + 32 + 2 * QB * QB) * sizeof(double);
  static const bool force_old = getenv("HTN_QR_UNBLOCKED") != nullptr;
  if (need <= 220 * 1024 && !force_old) {
    static size_t configured = 0;
    if (need > configured) {
      cudaFuncSetAttribute(qr_bcgs2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      configured = 220 * 1024;
    }
    qr_bcgs2_kernel<<<npanels, 512, need, st>>>(panels, A, R, status);
  } else {
    qr_cgs2_kernel<<<npanels, 32 * QR_TY, 0, st>>>(panels, A, R, status);
  }
}

// straight 2-D block copy with scale (tiles listed by the host): dst[r][c] = scale * src[r][c]
__global__ void __launch_bounds__(256) copy2d_kernel(const TrBlock* __restrict__ tiles, const double* __restrict__ src,
                                                     double* __restrict__ dst) {
  const TrBlock T = tiles[blockIdx.x];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < T.rows; r += 8)
    for (int c = tx; c < T.cols; c += 32) dst[T.doff + (long long)r * T.ldd + c] = T.scale * src[T.soff + (long long)r * T.lds + c];
}

void launch_copy2d(const TrBlock* tiles, int ntiles, const double* src, double* dst, cudaStream_t st) {
  if (ntiles > 0) copy2d_kernel<<<ntiles, 256, 0, st>>>(tiles, src, dst);
}

// ------------------------------------------------------------------------------------
// One-sided Jacobi SVD kernels (see launch_svd below): tournament ordering gives k/2 independent
// rotations per round, one warp per rotation pair.
// ------------------------------------------------------------------------------------
constexpr int SVD_WARPS = 16;

// Q = identity for every panel
__global__ void __launch_bounds__(256) svd_init_kernel(const SvdPanel* __restrict__ panels, double* __restrict__ Qb) {
  const SvdPanel P = panels[blockIdx.x];
  double* Q = Qb + P.offQ;
  for (int e = threadIdx.x; e < P.k * P.ldq; e += blockDim.x) Q[e] = (e / P.ldq == e % P.ldq) ? 1.0 : 0.0;
}

// One tournament round for ALL panels: entry e = (panel, pair slot j); one warp per rotation pair, so a
// round exposes sum_p k_p / 2 independent rotations to the whole GPU (a single CTA per panel left 147 SMs
// idle on the largest coupled block, which dominates: profiles/r1c notes).  nrot[panel] counts rotations.
__global__ void __launch_bounds__(256)
svd_round_kernel(const SvdPanel* __restrict__ panels, const int2* __restrict__ entries, int nentries, int round,
                 double* __restrict__ Gb, double* __restrict__ Qb, int* __restrict__ nrot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 8 + warp;
  if (e >= nentries) return;
  const int2 en = entries[e];
  const SvdPanel P = panels[en.x];
  const int k = P.k, len = P.len, ldg = P.ldg, ldq = P.ldq;
  const int kk = (k + 1) & ~1, n1 = kk - 1, j = en.y;
  if (round >= n1) return;
  int p = j == 0 ? round : (round + j) % n1;
  int q = j == 0 ? n1 : (round - j + n1) % n1;
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
  if (q >= k) return;
  double* gp = Gb + P.offG + (long long)p * ldg;
  double* gq = Gb + P.offG + (long long)q * ldg;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = lane; i < len; i += 32) {
    const double x = gp[i], y = gq[i];
    a = fma(x, x, a);
    b = fma(y, y, b);
    c = fma(x, y, c);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const double tol = fmax(1e-15, 5e-16 * sqrt((double)len));
  if (a == 0.0 || b == 0.0 || fabs(c) <= tol * sqrt(a * b)) return;
  const double zeta = (b - a) / (2.0 * c);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
  for (int i = lane; i < len; i += 32) {
    const double x = gp[i], y = gq[i];
    gp[i] = cs * x - sn * y;
    gq[i] = sn * x + cs * y;
  }
  double* qp = Qb + P.offQ + (long long)p * ldq;
  double* qq = Qb + P.offQ + (long long)q * ldq;
  for (int i = lane; i < k; i += 32) {
    const double x = qp[i], y = qq[i];
    qp[i] = cs * x - sn * y;
    qq[i] = sn * x + cs * y;
  }
  if (lane == 0) atomicAdd(nrot + en.x, 1);
}

// singular values, descending sort, normalisation and sign fix per panel (one CTA per panel)
__global__ void __launch_bounds__(32 * SVD_WARPS)
svd_finish_kernel(const SvdPanel* __restrict__ panels, const double* __restrict__ Gb, const double* __restrict__ Qb,
                  double* __restrict__ G2b, double* __restrict__ Q2b, double* __restrict__ sigb) {
  const SvdPanel P = panels[blockIdx.x];
  const double* G = Gb + P.offG;
  const double* Q = Qb + P.offQ;
  double* G2 = G2b + P.offG;
  double* Q2 = Q2b + P.offQ;
  double* sig = sigb + P.offS;
  const int k = P.k, len = P.len, ldg = P.ldg, ldq = P.ldq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ double ssig[1024];
  for (int i = warp; i < k; i += SVD_WARPS) {
    const double* gi = G + (long long)i * ldg;
    double a = 0.0;
    for (int e = lane; e < len; e += 32) a = fma(gi[e], gi[e], a);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ssig[i] = sqrt(a);
  }
  __syncthreads();
  for (int i = warp; i < k; i += SVD_WARPS) {
    const double si = ssig[i];
    int rank = 0;
    for (int j = lane; j < k; j += 32) rank += (ssig[j] > si || (ssig[j] == si && j < i)) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    const double* gi = G + (long long)i * ldg;
    const double* qi = Q + (long long)i * ldq;
    // left singular vector lives in G (u_in_g) or in Q: find its largest-magnitude entry (first one)
    const double* u = P.u_in_g ? gi : qi;
    const int ulen = P.u_in_g ? len : k;
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int e = lane; e < ulen; e += 32) {
      const double v = fabs(u[e]);
      if (v > best) {
        best = v;
        bidx = e;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) {
        best = ob;
        bidx = oi;
      }
    }
    const double sgn = (bidx < ulen && u[bidx] < 0.0) ? -1.0 : 1.0;
    const double inv = si > 0.0 ? sgn / si : 0.0;
    for (int e = lane; e < len; e += 32) G2[(long long)rank * ldg + e] = gi[e] * inv;
    for (int e = lane; e < k; e += 32) Q2[(long long)rank * ldq + e] = qi[e] * sgn;
    if (lane == 0) sig[rank] = si;
  }
}

// One-sided Jacobi SVD (Hestenes) of row panels (replaces LAPACK gesdd/gesvd behind TensorKit `tsvd!`;
// SURVEY.md 8(a) a9).  The k rows of G (length len) are rotated pairwise until mutually orthogonal; the
// same rotations accumulate in Q (k x k, starts as identity):  G_final = Q G_0 = diag(sigma) W^T.
// Output (sorted by descending sigma, sign-fixed so that the largest entry of every LEFT singular vector
// is positive): sig[k], G2 = W^T (unit rows), Q2 = Q.  High relative accuracy.  Host-steered: one launch
// per tournament round over all panels, one readback of the rotation counters per sweep.
int launch_svd(const SvdPanel* panels_dev, const SvdPanel* panels_host, int npanels, double* G, double* Q, double* G2,
               double* Q2, double* sig, cudaStream_t st) {
  if (npanels <= 0) return 0;
  std::vector<int2> entries;
  int max_n1 = 0;
  for (int p = 0; p < npanels; ++p) {
    const int kk = (panels_host[p].k + 1) & ~1;
    max_n1 = std::max(max_n1, kk - 1);
    if (panels_host[p].k >= 2)
      for (int j = 0; j < kk / 2; ++j) entries.push_back(make_int2(p, j));
  }
  int2* d_entries = nullptr;
  int* d_nrot = nullptr;
  int rc = 0;
  svd_init_kernel<<<npanels, 256, 0, st>>>(panels_dev, Q);
  if (!entries.empty()) {
    if (cudaMalloc(&d_entries, entries.size() * sizeof(int2)) != cudaSuccess ||
        cudaMalloc(&d_nrot, npanels * sizeof(int)) != cudaSuccess) {
      cudaFree(d_entries);
      return -1;
    }
    h2d_on_stream(d_entries, entries.data(), entries.size() * sizeof(int2), st);
    const int ne = (int)entries.size(), grid = (ne + 7) / 8;
    std::vector<int> nrot(npanels);
    bool converged = false;
    // one sweep = max_n1 dependent round launches: captured once into a CUDA graph and replayed per sweep
    // (a round on the largest panel runs ~5 us, so per-launch overhead would otherwise dominate)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool use_graph = max_n1 >= 16;
    if (use_graph) {
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        for (int r = 0; r < max_n1; ++r) svd_round_kernel<<<grid, 256, 0, st>>>(panels_dev, d_entries, ne, r, G, Q, d_nrot);
        if (cudaStreamEndCapture(st, &graph) != cudaSuccess || graph == nullptr ||
            cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
          use_graph = false;
          cudaGetLastError();
        }
      } else {
        use_graph = false;
        cudaGetLastError();
      }
    }
    for (int sweep = 0; sweep < 60 && !converged; ++sweep) {
      cudaMemsetAsync(d_nrot, 0, npanels * sizeof(int), st);
      if (use_graph)
        cudaGraphLaunch(gexec, st);
      else
        for (int r = 0; r < max_n1; ++r) svd_round_kernel<<<grid, 256, 0, st>>>(panels_dev, d_entries, ne, r, G, Q, d_nrot);
      cudaMemcpyAsync(nrot.data(), d_nrot, npanels * sizeof(int), cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) {
        rc = -2;
        break;
      }
      converged = true;
      for (int v : nrot) converged = converged && v == 0;
    }
    if (rc == 0 && !converged) rc = 1;
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    cudaFree(d_entries);
    cudaFree(d_nrot);
  }
  if (rc >= 0) svd_finish_kernel<<<npanels, 32 * SVD_WARPS, 0, st>>>(panels_dev, G, Q, G2, Q2, sig);
  return rc;
}

// ------------------------------------------------------------------------------------
// FP64 peak probes (roofline denominators measured on the box)
// ------------------------------------------------------------------------------------
template <int ITER>
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int ITER>
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

// DMMA issue-rate probe with few resident warps: `nacc` independent accumulators per warp
template <int NACC>
__global__ void __launch_bounds__(1024) probe_dmma_warps_kernel(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

double probe_fp64(int which, int sm_count, cudaStream_t st) {
  constexpr int ITER = 4096;
  if (which >= 10) {
    // which = 10 + warps_per_sm (4, 8, 16, 32): one CTA per SM, 14 independent DMMAs per warp
    const int warps = which - 10;
    double* d = nullptr;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0, st);
      probe_dmma_warps_kernel<14><<<sm_count, warps * 32, 0, st>>>(d, ITER);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      double tf = (double)sm_count * warps * ITER * 14.0 * 512.0 / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return best;
  }
  double* d = nullptr;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    if (which == 0)
      probe_dmma_kernel<ITER><<<grid, threads, 0, st>>>(d);
    else
      probe_dfma_kernel<ITER><<<grid, threads, 0, st>>>(d);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = which == 0 ? (double)grid * (threads / 32) * ITER * 8.0 * (2.0 * 8 * 8 * 4)
                              : (double)grid * threads * ITER * 16.0 * 2.0;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}  // namespace htn
