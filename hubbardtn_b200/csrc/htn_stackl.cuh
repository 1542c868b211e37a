// Stage L of the effective Hamiltonians as STACKED, B-STATIONARY FP64 tensor-core GEMMs (sm_100a).
//
//   T[a,l',l,s,r] = GL[a,l',l] . x[l,s,r]          (MPSKit `∂AC`: the GL . AC leg)
//
// All GL blocks that share the contracted sector l are stored back to back as one tall row-major panel
// A_l = [ GL[a,l',l] ]_(l',a)  (M_l x n_l, csrc/htn_api.cpp:finalize_tensor), so for one block x[l,s,r] the whole
// family { T[a,l',l,s,r] }_(a,l') is ONE product  A_l (M_l x n_l) . x[l,s,r] (n_l x n_r)  whose 64-row tiles run
// across block boundaries and are always full.  A job = (row run of A_l) x (<= 64 columns of one x block):
//   * the x columns (n_l x NT, the "slab", <= 80 KB) are loaded ONCE per job and stay in shared memory,
//   * A streams through a 4-stage ring of 64 x 16 chunks.  When the panel is a bound environment tensor the chunk is
//     ONE 2-D TMA box copy (cp.async.bulk.tensor.2d, 8 KB, SWIZZLE_128B, out-of-range rows/columns zero-filled by the
//     hardware, completion by mbarrier complete_tx; SASS UTMALDG); otherwise 16-byte cp.async into the same layout.
//     (1-D cp.async.bulk copies of one 128-byte row each were measured first: the TMA unit serves ~1 such copy per
//     40 cycles per SM, 0.9 TB/s chip-wide -- 12.7 TF/s; profiles/r2_stackl_probe.txt.)
//   * four consumer warps own 16 rows x NT columns each (2 x CA DMMA.8x8x4 atoms), software-specialised on CA; the
//     fragment rows of an 8-row atom are permuted (0,2,4,6 | 1,3,5,7) so that the 128-byte TMA swizzle is
//     conflict-free for the 8 x 4 FP64 fragment loads,
//   * jobs are handed out dynamically (one atomic ticket per job, descending cost order inside a wave),
//   * STAGE W RIDES ALONG: three mixer warps per CTA form the recoupled blocks U = sum coef T (and the direct T -> y
//     terms) for the wave of l' sectors whose T blocks have just been finished, while the DMMA warps work on the next
//     wave -- T is read back out of L2 instead of HBM and the mix costs no time of its own (it was 20 % of the apply).
// Per 16-k chunk a CTA moves 8 KB for 2*64*NT*16 flops (twice the intensity of the 64x64 two-operand ring of
// grouped_gemm_kernel), and two CTAs per SM keep two warps per scheduler on the DMMA pipe, which
// tools/dmma_tile_probe.cu shows is enough for 99 % of the issue-loop peak.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "htn_internal.hpp"

namespace htn {

constexpr int SL_TM = 64;            // rows of A per tile
constexpr int SL_KC = 16;            // k extent of one ring stage (128 bytes: one swizzle span)
#ifndef SL_RING_STAGES
#define SL_RING_STAGES 4
#endif
constexpr int SL_STAGES = SL_RING_STAGES;
constexpr int SL_STAGE_ELEMS = SL_TM * SL_KC;  // 8 KB, unpadded, 128B-swizzled
constexpr int SL_SLAB = STACK_SLAB_ELEMS;      // slab capacity (doubles): round4(K) * (8 CA + 4) must fit
constexpr int SL_GW = SL_TM / 16;              // warps of one consumer group: 16 rows each
constexpr int SL_JOBWORDS = (int)(sizeof(StackJob) / 4);
constexpr int SL_MIXSRC = 32;                  // source descriptors staged per mixer warp
constexpr int SL_JOBQ = 2;
// Two shapes of the CTA (template parameter NG = consumer groups):
//   NG = 1: 4 consumer warps + 1 producer + 3 mixers = 256 threads, 2 CTAs per SM (each with its own 80 KB slab);
//   NG = 2: 2 x 4 consumer warps working on alternate 64-row tiles of the SAME job (ONE slab), one producer warp and one
//           A ring per group, 6 mixers = 512 threads, 1 CTA per SM.  The second slab's 80 KB become the staging area
//           into which the mixers pull their sources with cp.async.bulk (see sl_mix_chunk_tma).
constexpr int SLM_SLOT_BYTES = 2048;           // one staged piece of one mix source: 256 doubles
constexpr int SLM_SLOTS = 6;                   // pieces in flight per mixer warp
template <int NG>
struct SlCfg {
  static constexpr int NCONS = SL_GW * NG;
  static constexpr int NPROD = NG;
  static constexpr int NMIX = NG == 1 ? 3 : 6;
  static constexpr int WARPS = NCONS + NPROD + NMIX;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int RING_ELEMS = NG * SL_STAGES * SL_STAGE_ELEMS;
  static constexpr int MIXSTAGE_BYTES = NG == 1 ? 0 : NMIX * SLM_SLOTS * SLM_SLOT_BYTES;
  static constexpr int NBAR = NG * 2 * SL_STAGES + 2 * SL_JOBQ + (NG == 1 ? WARPS * SLM_SLOTS : (NMIX + NCONS) * SLM_SLOTS);
  static constexpr int SMEM_BYTES = (RING_ELEMS + SL_SLAB) * 8 + MIXSTAGE_BYTES + NBAR * 8 + SL_JOBQ * SL_JOBWORDS * 4 + 32;
  static constexpr int CTAS_PER_SM = NG == 1 ? 2 : 1;
};
// 2 CTAs per SM need 2 x (this + 1 KB) <= 228 KB: the slab leaves no room for per-warp staging of the mix sources
// (they travel through warp shuffles instead)
static_assert(2 * (SlCfg<1>::SMEM_BYTES + 1024) <= 233472, "two CTAs per SM must fit");
static_assert(SlCfg<2>::SMEM_BYTES + 1024 <= 233472, "the one-CTA shape must fit");

__device__ __forceinline__ unsigned sl_smem(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sl_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(sl_smem(bar)), "r"(count));
}
__device__ __forceinline__ void sl_mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(sl_smem(bar)) : "memory");
}
__device__ __forceinline__ void sl_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(sl_smem(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ unsigned sl_mbar_try(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(sl_smem(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void sl_mbar_wait(uint64_t* bar, unsigned parity) {
  while (!sl_mbar_try(bar, parity)) {
  }
}
// 2-D TMA box copy global -> shared through a tensor map, completion counted in bytes on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void sl_tma_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   sl_smem(dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(sl_smem(bar))
               : "memory");
}
// the same with an L2 eviction-priority hint (the environment panels are re-read by every column piece of every x block
// of their sector: evict_last keeps them in front of the streaming mix traffic)
__device__ __forceinline__ void sl_tma_2d_hint(void* dst, const void* tmap, int c0, int c1, uint64_t* bar, unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(
          sl_smem(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(sl_smem(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ unsigned long long sl_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long sl_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ void sl_cp16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sl_smem(dst)), "l"(src), "r"(src_bytes));
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-charged at init)
__device__ __forceinline__ void sl_cp_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(sl_smem(bar)) : "memory");
}
__device__ __forceinline__ void sl_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ unsigned long long sl_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

struct SlRing {
  double* ring;
  uint64_t* full;
  uint64_t* empty;
  int stage;
  unsigned phase;
  __device__ __forceinline__ void advance() {
    if (++stage == SL_STAGES) {
      stage = 0;
      phase ^= 1u;
    }
  }
};

// tile-level waves: one consumer warp's strip of a tile is stored -- report to every wave whose T blocks the tile covers
__device__ __forceinline__ void sl_report_tile(const StackArgs& a, int2 tw, int dbg) {
  if (tw.y < tw.x) return;
  __threadfence();
  for (int w = tw.x; w <= tw.y; ++w) {
    const unsigned long long old = atomicAdd(a.ctr + 4 + w, 1ull);
    if ((dbg & 32) && a.dbg_ts && old + 1 == a.epoch * (unsigned long long)a.wave_need[w]) a.dbg_ts[2 * a.nmix + w] = sl_globaltimer();
  }
}

// consumer side of one job, specialised on the number of 8-column atoms of the slab
template <int CA>
__device__ __forceinline__ void sl_consume_job(const StackJob& job, SlRing& rg, const double* __restrict__ slab, int role,
                                               int lane, const Bases& bases, int dbg, int m_begin, int m_step,
                                               const StackArgs& a) {
  const int g = lane >> 2, t = lane & 3;
  const int rho = 2 * (g & 3) + (g >> 2);  // tile row (within an 8-row atom) held by fragment row g
  const int SB = CA * 8 + 4;
  const int K = job.K;
  const int nchunks = (K + SL_KC - 1) / SL_KC;
  // A element (row, k = 4 kk + t) of a stage: row * 16 + ((2 kk + (t >> 1)) ^ (row & 7)) * 2 + (t & 1)
  const double* ap = rg.ring + (role * 16 + rho) * SL_KC + (t & 1);
  int koff[4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) koff[kk] = ((2 * kk + (t >> 1)) ^ rho) << 1;
  const double* bp = slab + t * SB + g;  // + (16 c + 4 kk) * SB + 8 j
  double* Cb = const_cast<double*>(resolve(job.c_off, job.c_base, bases));
  const long long ldc = job.ldc;
  const int nt = job.nt;
  unsigned ready = 0;
  int2 pend = make_int2(0, -1);  // waves of the previous tile: its stores have drained by the time the next K loop ends,
                                 // so the fence in front of the report costs nothing there
  for (int m0 = m_begin; m0 < job.M; m0 += m_step) {  // this group's tiles of the job
    int2 tw_cur = make_int2(0, -1);
    if (job.tile0 >= 0 && lane == 0) tw_cur = a.tile_waves[job.tile0 + m0 / SL_TM];
    double acc[CA][2][2];
#pragma unroll
    for (int j = 0; j < CA; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
    const bool last_tile = m0 + m_step >= job.M;
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
      if (!ready) sl_mbar_wait(&rg.full[rg.stage], rg.phase);
      const double* as = ap + rg.stage * SL_STAGE_ELEMS;
      const double* bs = bp + c * (SL_KC * SB);
      const int cur = rg.stage;
      rg.advance();
      const bool more = !(last_tile && c + 1 == nchunks);
      const int krem = K - c * SL_KC;
      if (krem >= SL_KC) {
#pragma unroll
        for (int kk = 0; kk < SL_KC / 4; ++kk) {
          const double a0 = as[koff[kk]], a1 = as[8 * SL_KC + koff[kk]];
          double b[CA];
#pragma unroll
          for (int j = 0; j < CA; ++j) b[j] = bs[kk * 4 * SB + j * 8];
          // poll the NEXT stage's barrier while the last k4-step of this one is still to be issued: its ~90-cycle
          // latency disappears behind the DMMAs
          if (kk == SL_KC / 4 - 1) ready = more ? sl_mbar_try(&rg.full[rg.stage], rg.phase) : 0u;
#pragma unroll
          for (int j = 0; j < CA; ++j) {
            sl_dmma(acc[j][0][0], acc[j][0][1], a0, b[j]);
            sl_dmma(acc[j][1][0], acc[j][1][1], a1, b[j]);
          }
        }
      } else {
        const int nk4 = (krem + 3) >> 2;
        ready = more ? sl_mbar_try(&rg.full[rg.stage], rg.phase) : 0u;
        for (int kk = 0; kk < nk4; ++kk) {
          const double a0 = as[koff[kk]], a1 = as[8 * SL_KC + koff[kk]];
          double b[CA];
#pragma unroll
          for (int j = 0; j < CA; ++j) b[j] = bs[kk * 4 * SB + j * 8];
#pragma unroll
          for (int j = 0; j < CA; ++j) {
            sl_dmma(acc[j][0][0], acc[j][0][1], a0, b[j]);
            sl_dmma(acc[j][1][0], acc[j][1][1], a1, b[j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) sl_mbar_arrive(&rg.empty[cur]);
    }
    if (job.tile0 >= 0 && lane == 0) sl_report_tile(a, pend, dbg);
    // ---- store the 16 x nt strip of this warp (rows beyond the run / columns beyond nt are dropped) ----
    const int mt = min(SL_TM, job.M - m0);
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int row = role * 16 + f * 8 + rho;
      if (row < mt && !(dbg & 2)) {  // timing experiment 2: no stores
        double* p = Cb + (long long)(m0 + row) * ldc + 2 * t;
#pragma unroll
        for (int j = 0; j < CA; ++j) {
          const int col = j * 8 + 2 * t;
          if (j < CA - 1 || col + 1 < nt)
            *reinterpret_cast<double2*>(p + j * 8) = make_double2(acc[j][f][0], acc[j][f][1]);
          else if (col < nt)
            p[j * 8] = acc[j][f][0];
        }
      }
    }
    pend = tw_cur;  // reported behind the next tile's K loop (or at the end of the job)
  }
  if (job.tile0 >= 0) {
    __syncwarp();
    if (lane == 0) sl_report_tile(a, pend, dbg);
  }
}

__device__ __forceinline__ unsigned long long sl_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// one mix chunk by one warp: dst[e] = sum_s coef_s src_s[e] over a flat element range (identical padded layouts)
// weak 16-byte global load that does not allocate in L1 (the data was written by other SMs earlier in this launch and is
// read exactly once here; weak loads keep all 2 U requests of a trip in flight -- ld.global.cg compiles to ordered
// LDG.STRONG.GPU and was 5x slower)
__device__ __forceinline__ double2 sl_ld_stream(const double* p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// streaming 16-byte store: U is written once and read next by stage R (another launch) -- keep it from pushing T out of L2
__device__ __forceinline__ void sl_st_stream(double* p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// one mix chunk by one warp: dst[e] = sum_s coef_s src_s[e] over a flat element range (identical padded layouts)
template <int U>
__device__ __forceinline__ void sl_mix_chunk(const StackArgs& a, const MixChunk& ch, const Bases& bases, int lane) {
  const MixTarget T = a.mt[ch.target];
  double* dst = const_cast<double*>(resolve(T.off, T.base, bases));
  const int nsrc = T.src_end - T.src_begin;
  const int end = ch.elem0 + ch.nelem;  // even, >= 2
  for (int s0 = 0; s0 < nsrc || s0 == 0; s0 += SL_MIXSRC) {
    const int ns = min(SL_MIXSRC, nsrc - s0);
    // lane s holds source s of this batch; the trip loop below broadcasts them by shuffle
    unsigned long long my_ptr = 0;
    double my_coef = 0.0;
    if (lane < ns) {
      const MixSrc S = a.ms[T.src_begin + s0 + lane];
      my_ptr = reinterpret_cast<unsigned long long>(resolve(S.off, S.base, bases));
      my_coef = S.coef;
    }
    // U positions (double2 each) per lane and trip: 64 U elements per warp trip; positions beyond the chunk are
    // clamped to its last pair (loaded, never stored) so that the loads of a trip are straight-line code
    for (int eb = ch.elem0; eb < end; eb += 64 * U) {  // warp-uniform trip count (the shuffles below need all lanes)
      const int e0 = eb + 2 * lane;
      int eu[U];
#pragma unroll
      for (int u = 0; u < U; ++u) eu[u] = min(e0 + 64 * u, end - 2);
      double2 acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = s0 == 0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(dst + eu[u]);
      int s = 0;
      for (; s + 2 <= ns; s += 2) {  // two sources at a time: 2 U independent loads in flight per lane
        const double* sp0 = reinterpret_cast<const double*>(__shfl_sync(0xffffffffu, my_ptr, s));
        const double* sp1 = reinterpret_cast<const double*>(__shfl_sync(0xffffffffu, my_ptr, s + 1));
        const double cf0 = __shfl_sync(0xffffffffu, my_coef, s), cf1 = __shfl_sync(0xffffffffu, my_coef, s + 1);
        double2 v0[U], v1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v0[u] = sl_ld_stream(sp0 + eu[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) v1[u] = sl_ld_stream(sp1 + eu[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[u].x = fma(cf0, v0[u].x, acc[u].x);
          acc[u].y = fma(cf0, v0[u].y, acc[u].y);
          acc[u].x = fma(cf1, v1[u].x, acc[u].x);
          acc[u].y = fma(cf1, v1[u].y, acc[u].y);
        }
      }
      if (s < ns) {
        const double* sp = reinterpret_cast<const double*>(__shfl_sync(0xffffffffu, my_ptr, s));
        const double cf = __shfl_sync(0xffffffffu, my_coef, s);
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = sl_ld_stream(sp + eu[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          acc[u].x = fma(cf, v[u].x, acc[u].x);
          acc[u].y = fma(cf, v[u].y, acc[u].y);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + 64 * u;
        if (e < end) {
          if (nsrc <= SL_MIXSRC)
            sl_st_stream(dst + e, acc[u]);
          else
            *reinterpret_cast<double2*>(dst + e) = acc[u];  // re-read by the next source batch
        }
      }
    }
    if (nsrc == 0) break;
  }
}

// mixer loop: mix chunks by ticket, each after the wave of stack jobs it reads from is complete.
// PROF (HTN_STACK_DEBUG & 32, a separate instantiation so that the product loop carries none of it): clocks spent per
// chunk on ticket + record, on the wave wait and on the data, summed into the probe slots behind the wave counters.
template <int U, bool PROF>
__device__ __forceinline__ bool sl_mixer_loop_t(const StackArgs& a, const Bases& bases, int lane, int nwarps_total,
                                                const volatile int* switch_flag = nullptr) {
  const unsigned long long base = (a.epoch - 1ull) * (unsigned long long)(a.nmix + nwarps_total);
  long long c_desc = 0, c_wait = 0, c_work = 0, n_chunks = 0;
  bool switched = false;
  while (true) {
    // the CTA's shared memory has become free (all jobs done): leave before drawing a ticket, the caller carries on with the
    // bulk-copy fed loop
    if (switch_flag && __shfl_sync(0xffffffffu, *switch_flag, 0)) {
      switched = true;
      break;
    }
    const long long t0 = PROF ? clock64() : 0;
    unsigned long long tk = 0;
    if (lane == 0) tk = atomicAdd(a.ctr + 1, 1ull) - base;
    tk = __shfl_sync(0xffffffffu, tk, 0);
    if (tk >= (unsigned long long)a.nmix) break;
    const MixChunk ch = a.mc[tk];
    const long long t1 = PROF ? clock64() : 0;
    if (ch.pad_ >= 0) {
      if (lane == 0) {
        // wave 0 holds the jobs of the light panels (every left sector reads from them), wave pad_ the heavy ones
        for (int w = ch.pad_;; w = 0) {
          const unsigned long long need = a.epoch * (unsigned long long)a.wave_need[w];
          unsigned spins = 0;
          while (sl_ld_acquire(a.ctr + 4 + w) < need) {
            __nanosleep(256);
            if (++spins > (1u << 22)) {  // ~1 s: something is wrong; flag it and go on (results are then invalid)
              atomicAdd(a.ctr + 2, 1ull);
              break;
            }
          }
          if (w == 0) break;
        }
      }
      __syncwarp();
    }
    const long long t2 = PROF ? clock64() : 0;
    if (PROF && a.dbg_ts && lane == 0) a.dbg_ts[2 * tk] = sl_globaltimer();
    if (!(a.dbg & 4)) sl_mix_chunk<U>(a, ch, bases, lane);
    if (a.mix_lag > 0 && ch.pad_ >= 0) {  // back-pressure experiments: count the finished chunks of the wave
      __syncwarp();
      if (lane == 0) atomicAdd(a.ctr + 4 + a.nwaves + ch.pad_, 1ull);
    }
    if (PROF) {
      const long long t3 = clock64();
      if (a.dbg_ts && lane == 0) a.dbg_ts[2 * tk + 1] = sl_globaltimer();
      c_desc += t1 - t0;
      c_wait += t2 - t1;
      c_work += t3 - t2;
      ++n_chunks;
    }
  }
  if (PROF && lane == 0) {
    unsigned long long* d = a.ctr + 4 + 2 * a.nwaves;
    atomicMax(d + 2, sl_globaltimer());
    atomicAdd(d + 4, (unsigned long long)c_desc);
    atomicAdd(d + 5, (unsigned long long)c_wait);
    atomicAdd(d + 6, (unsigned long long)c_work);
    atomicAdd(d + 7, (unsigned long long)n_chunks);
  }
  return switched;
}
template <int U>
__device__ __forceinline__ bool sl_mixer_loop_prof(const StackArgs& a, const Bases& bases, int lane, int nwarps_total,
                                                   const volatile int* switch_flag) {
  return sl_mixer_loop_t<U, true>(a, bases, lane, nwarps_total, switch_flag);
}
template <int U>
__device__ __forceinline__ bool sl_mixer_loop(const StackArgs& a, const Bases& bases, int lane, int nwarps_total,
                                              const volatile int* switch_flag = nullptr) {
  if (a.dbg & 32) return sl_mixer_loop_prof<U>(a, bases, lane, nwarps_total, switch_flag);
  return sl_mixer_loop_t<U, false>(a, bases, lane, nwarps_total, switch_flag);
}

// ---- mixers of the one-CTA shape: sources staged through shared memory by cp.async.bulk ----
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void sl_bulk_g2s_hint(void* dst, const void* src, unsigned bytes, uint64_t* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                   sl_smem(dst)),
               "l"(src), "r"(bytes), "r"(sl_smem(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void sl_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sl_smem(dst)),
               "l"(src), "r"(bytes), "r"(sl_smem(bar))
               : "memory");
}

// One mixer warp: chunks by ticket; the (piece, source) items of a chunk -- 2 KB of one source each -- are pulled into the
// warp's SLM_SLOTS staging slots SLM_SLOTS items ahead of their use, so that 12 KB per warp (72 KB per SM) are in flight
// all the time without holding registers; a lane reads 4 x 16 bytes of a landed slot, accumulates coef * v and, after the
// last source of a piece, streams the piece out.  The ticket of the NEXT chunk is drawn before the data loop of the
// current one, its 32-byte record is loaded right after (both hidden behind the data).
__device__ __forceinline__ void sl_mixer_loop_tma(const StackArgs& a, const Bases& bases, int lane, int nwarps_total,
                                                  unsigned char* stage, uint64_t* bar) {
  const unsigned long long base = (a.epoch - 1ull) * (unsigned long long)(a.nmix + nwarps_total);
  const bool prof = (a.dbg & 32) != 0;
  long long c_wait = 0, c_work = 0, n_chunks = 0;
  unsigned it = 0;  // items issued so far by this warp (slot = it % SLM_SLOTS, parity = (it / SLM_SLOTS) & 1)
  const unsigned long long pol_first = sl_policy_evict_first();
  unsigned long long tk = 0;
  if (lane == 0) tk = atomicAdd(a.ctr + 1, 1ull) - base;
  tk = __shfl_sync(0xffffffffu, tk, 0);
  while (tk < (unsigned long long)a.nmix) {
    const MixChunkX ch = a.mcx[tk];
    // next ticket: in flight during this chunk
    unsigned long long tk_next = 0;
    if (lane == 0) tk_next = atomicAdd(a.ctr + 1, 1ull) - base;
    const long long t1 = prof ? clock64() : 0;
    if (ch.wave >= 0) {
      if (lane == 0) {
        for (int w = ch.wave;; w = 0) {
          const unsigned long long need = a.epoch * (unsigned long long)a.wave_need[w];
          unsigned spins = 0;
          while (sl_ld_acquire(a.ctr + 4 + w) < need) {
            __nanosleep(256);
            if (++spins > (1u << 22)) {
              atomicAdd(a.ctr + 2, 1ull);
              break;
            }
          }
          if (w == 0) break;
        }
      }
      __syncwarp();
    }
    const long long t2 = prof ? clock64() : 0;
    if (prof && a.dbg_ts && lane == 0) a.dbg_ts[2 * tk] = sl_globaltimer();
    if (!(a.dbg & 4)) {
      double* dst = const_cast<double*>(resolve(ch.off, ch.base, bases)) + ch.elem0;
      const int npieces = (ch.nelem + 255) >> 8;
      for (int s0 = 0; s0 < ch.nsrc || s0 == 0; s0 += 32) {
        const int ns = min(32, ch.nsrc - s0);
        unsigned long long my_ptr = 0;
        double my_coef = 0.0;
        if (lane < ns) {
          const MixSrc S = a.ms[ch.src_begin + s0 + lane];
          my_ptr = reinterpret_cast<unsigned long long>(resolve(S.off, S.base, bases) + ch.elem0);
          my_coef = S.coef;
        }
        const int total = npieces * max(ns, 0);
        int issued = 0;
        auto issue = [&](int item) {  // all lanes call (shuffle); lane 0 issues the copy
          const int p = item / ns, s = item - p * ns;
          const unsigned long long sp = __shfl_sync(0xffffffffu, my_ptr, s);
          if (lane == 0) {
            const unsigned slot = it % SLM_SLOTS;
            const unsigned bytes = (unsigned)min(256, ch.nelem - (p << 8)) * 8u;
            sl_mbar_expect_tx(&bar[slot], bytes);
            if (a.dbg & 256)
              sl_bulk_g2s_hint(stage + slot * SLM_SLOT_BYTES, reinterpret_cast<const double*>(sp) + (p << 8), bytes, &bar[slot], pol_first);
            else
              sl_bulk_g2s(stage + slot * SLM_SLOT_BYTES, reinterpret_cast<const double*>(sp) + (p << 8), bytes, &bar[slot]);
          }
          ++it;
        };
        const unsigned it0 = it;
        for (; issued < total && issued < SLM_SLOTS; ++issued) issue(issued);
        double2 acc[4];
        for (int c = 0; c < total; ++c) {
          const int p = c / ns, s = c - p * ns;
          const unsigned gi = it0 + (unsigned)c;
          const unsigned slot = gi % SLM_SLOTS, par = (gi / SLM_SLOTS) & 1u;
          sl_mbar_wait(&bar[slot], par);
          const double cf = __shfl_sync(0xffffffffu, my_coef, s);
          const double2* sv = reinterpret_cast<const double2*>(stage + slot * SLM_SLOT_BYTES) + lane;
          const int e = (p << 8) + 2 * lane;  // element of the chunk held by this lane in quarter q: e + 64 q
          if (s == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              acc[q] = (s0 == 0 || e + 64 * q >= ch.nelem) ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(dst + e + 64 * q);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 v = sv[32 * q];
            acc[q].x = fma(cf, v.x, acc[q].x);
            acc[q].y = fma(cf, v.y, acc[q].y);
          }
          if (s == ns - 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (e + 64 * q < ch.nelem) {
                if (s0 + 32 >= ch.nsrc)
                  sl_st_stream(dst + e + 64 * q, acc[q]);
                else
                  *reinterpret_cast<double2*>(dst + e + 64 * q) = acc[q];
              }
          }
          __syncwarp();  // every lane has read the slot: it may be refilled
          if (issued < total) {
            issue(issued);
            ++issued;
          }
        }
        if (ch.nsrc == 0) {  // a target without sources is cleared
          for (int e = 2 * lane; e < ch.nelem; e += 64) *reinterpret_cast<double2*>(dst + e) = make_double2(0.0, 0.0);
          break;
        }
      }
    }
    if (a.mix_lag > 0 && ch.wave >= 0) {
      __syncwarp();
      if (lane == 0) atomicAdd(a.ctr + 4 + a.nwaves + ch.wave, 1ull);
    }
    if (prof) {
      const long long t3 = clock64();
      c_wait += t2 - t1;
      c_work += t3 - t2;
      ++n_chunks;
      if (a.dbg_ts && lane == 0) a.dbg_ts[2 * tk + 1] = sl_globaltimer();
    }
    tk = __shfl_sync(0xffffffffu, tk_next, 0);
  }
  if (prof && lane == 0) {
    unsigned long long* d = a.ctr + 4 + 2 * a.nwaves;
    atomicMax(d + 2, sl_globaltimer());
    atomicAdd(d + 5, (unsigned long long)c_wait);
    atomicAdd(d + 6, (unsigned long long)c_work);
    atomicAdd(d + 7, (unsigned long long)n_chunks);
  }
}

// Roles: warps 0 .. NCONS-1 = consumers in groups of four (DMMA; all of them load the slab of a job together while the A
// chunks the producers have already queued wait in the rings: a job switch costs one slab latency); warps NCONS ..
// NCONS+NPROD-1 = producers, one per consumer group (producer 0 also draws the job tickets and publishes the job records);
// the rest = mixers (stage W).
// Register budget: launched with 128 per thread; the consumer warp groups grow (with 124 the DMMA loops lose a third of
// their speed: fewer fragment loads in flight), the other warp groups shrink.
template <int NG>
__global__ void __launch_bounds__(SlCfg<NG>::THREADS, SlCfg<NG>::CTAS_PER_SM)
    stack_gemm_kernel(const __grid_constant__ StackArgs a, const __grid_constant__ Bases bases) {
  using Cfg = SlCfg<NG>;
  constexpr int NCONS = Cfg::NCONS, NPROD = Cfg::NPROD;
  extern __shared__ __align__(1024) double sl_sm[];
  double* rings = sl_sm;  // 1024-byte aligned stages (128-byte swizzle atom = 8 rows x 128 B)
  double* slab = sl_sm + Cfg::RING_ELEMS;
  unsigned char* mixstage = reinterpret_cast<unsigned char*>(slab + SL_SLAB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(mixstage + Cfg::MIXSTAGE_BYTES);
  uint64_t* full_all = bars;                                // [NG][SL_STAGES]
  uint64_t* empty_all = full_all + NG * SL_STAGES;          // [NG][SL_STAGES]
  uint64_t* jfull = empty_all + NG * SL_STAGES;
  uint64_t* jempty = jfull + SL_JOBQ;
  uint64_t* mixbar = jempty + SL_JOBQ;                      // [NMIX][SLM_SLOTS] (NG = 2)
  int* job_slot = reinterpret_cast<int*>(bars + Cfg::NBAR);  // SL_JOBQ records
  volatile int* smem_free = job_slot + SL_JOBQ * SL_JOBWORDS;  // NG = 1: set when rings and slab may be reused as mix staging
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dbg = a.dbg;
  // the rings must never hold NaN patterns: K tails multiply stale columns by the zero rows of the slab
  for (int i = tid; i < Cfg::RING_ELEMS; i += Cfg::THREADS) rings[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < NG * SL_STAGES; ++s) {
      sl_mbar_init(&full_all[s], 32);      // every producer lane arrives once (cp.async noinc, or plain / expect_tx)
      sl_mbar_init(&empty_all[s], SL_GW);  // one elected lane per consumer warp of the group
    }
    for (int s = 0; s < SL_JOBQ; ++s) {
      sl_mbar_init(&jfull[s], 1);
      sl_mbar_init(&jempty[s], NCONS + NPROD - 1);
    }
    for (int s = 0; s < (NG == 1 ? Cfg::WARPS : Cfg::NMIX + NCONS) * SLM_SLOTS; ++s) sl_mbar_init(&mixbar[s], 1);
    *smem_free = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic zero fill before async-proxy (TMA) writes
  __syncthreads();
  int jq = 0;
  unsigned jphase = 0;
  // warps that draw mix tickets (each overshoots the ticket counter by one): in the one-CTA shape the producers do not mix
  const int nwarps_total = (int)gridDim.x * (NG == 1 ? Cfg::WARPS : Cfg::WARPS - NPROD);

  if (warp < NCONS) {
    // =========================== CONSUMERS ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;\n");
    const int group = warp / SL_GW, role = warp % SL_GW;
    SlRing rg;
    rg.ring = rings + group * SL_STAGES * SL_STAGE_ELEMS;
    rg.full = full_all + group * SL_STAGES;
    rg.empty = empty_all + group * SL_STAGES;
    rg.stage = 0;
    rg.phase = 0;
    while (true) {
      sl_mbar_wait(&jfull[jq], jphase);
      StackJob job;
      const int* slot = job_slot + jq * SL_JOBWORDS;
#pragma unroll
      for (int w = 0; w < SL_JOBWORDS; ++w) reinterpret_cast<int*>(&job)[w] = slot[w];
      if (job.M < 0) break;
      // ---- slab: K rows of nb doubles, rows K .. round4(K) zero; loaded by all consumer threads ----
      asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");  // every warp has left the previous slab
      {
        const int CA = (job.nt + 7) >> 3, SB = CA * 8 + 4, K = job.K;
        const double* Bg = resolve(job.b_off, job.b_base, bases);
        const int np = job.nb >> 1, total = K * np;  // 16-byte pieces per row
        constexpr int NT_ = NCONS * 32;
        int k = tid / np, q = tid - k * np;
        const int dk = NT_ / np, dq = NT_ - dk * np;
        for (int i = tid; i < total; i += NT_) {
          sl_cp16(slab + k * SB + 2 * q, Bg + (long long)k * job.ldb + 2 * q, 16);
          k += dk;
          q += dq;
          if (q >= np) {
            q -= np;
            ++k;
          }
        }
        const int kz = ((K + 3) & ~3) - K;  // 0..3 rows
        for (int i = tid; i < kz * SB; i += NT_) slab[K * SB + i] = 0.0;
        asm volatile("cp.async.wait_all;\n" ::: "memory");
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");
      const int mb = group * SL_TM, ms = NG * SL_TM;
      switch ((job.nt + 7) >> 3) {
        case 1: sl_consume_job<1>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 2: sl_consume_job<2>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 3: sl_consume_job<3>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 4: sl_consume_job<4>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 5: sl_consume_job<5>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 6: sl_consume_job<6>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        case 7: sl_consume_job<7>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
        default: sl_consume_job<8>(job, rg, slab, role, lane, bases, dbg, mb, ms, a); break;
      }
      __syncwarp();
      if (lane == 0) {
        sl_mbar_arrive(&jempty[jq]);
        if (job.wave >= 0) {  // this warp's strips of the job's T tiles are written: publish them to the mixers
          __threadfence();
          const unsigned long long old = atomicAdd(a.ctr + 4 + job.wave, 1ull);
          if ((dbg & 32) && a.dbg_ts && old + 1 == a.epoch * (unsigned long long)a.wave_need[job.wave])
            a.dbg_ts[2 * a.nmix + job.wave] = sl_globaltimer();  // the arrival that completed the wave
        }
      }
      if (++jq == SL_JOBQ) {
        jq = 0;
        jphase ^= 1u;
      }
    }
    if ((dbg & 32) && lane == 0) {
      atomicMax(a.ctr + 4 + 2 * a.nwaves + 1, sl_globaltimer());
      // mix tickets drawn so far when this warp ran out of jobs (max over warps = at the end of the job phase)
      const unsigned long long drawn = sl_ld_acquire(a.ctr + 1) - (a.epoch - 1ull) * (unsigned long long)(a.nmix + nwarps_total);
      atomicMax(a.ctr + 4 + 2 * a.nwaves + 3, drawn);
    }
    // no stack jobs left: help with the mix (the last wave's targets are still to be formed)
    if (a.nmix > 0) {
      if (NG == 1 && !(dbg & 64)) {
        // every consumer warp has left its last tile: ring and slab (112 KB) are free.  They become staging slots of the
        // bulk-copy fed mix loop for all eight warps of the CTA (the mixers and the producer switch over at their next chunk)
        asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if (tid == 0) {
          *smem_free = 1;
          __threadfence_block();
        }
        sl_mixer_loop_tma(a, bases, lane, nwarps_total, reinterpret_cast<unsigned char*>(sl_sm) + warp * SLM_SLOTS * SLM_SLOT_BYTES,
                          mixbar + warp * SLM_SLOTS);
      } else if (NG == 1) {
        sl_mixer_loop<8>(a, bases, lane, nwarps_total);
      } else if (dbg & 64) {
        sl_mixer_loop<8>(a, bases, lane, nwarps_total);
      } else {
        // every consumer warp has left its last tile: rings and slab (144 KB) are free and become staging slots, so the
        // tail of the mix runs with 8 more bulk-copy fed warps
        asm volatile("bar.sync 1, %0;\n" ::"n"(NCONS * 32) : "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        sl_mixer_loop_tma(a, bases, lane, nwarps_total, reinterpret_cast<unsigned char*>(sl_sm) + warp * SLM_SLOTS * SLM_SLOT_BYTES,
                          mixbar + (Cfg::NMIX + warp) * SLM_SLOTS);
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;\n");
    if (warp < NCONS + NPROD) {
      // =========================== PRODUCERS ===========================
      const int pg = warp - NCONS;  // the consumer group this warp feeds
      SlRing rg;
      rg.ring = rings + pg * SL_STAGES * SL_STAGE_ELEMS;
      rg.full = full_all + pg * SL_STAGES;
      rg.empty = empty_all + pg * SL_STAGES;
      rg.stage = 0;
      rg.phase = 0;
      const unsigned long long base = (a.epoch - 1ull) * (unsigned long long)(a.njobs + (int)gridDim.x);
      if ((dbg & 32) && lane == 0 && pg == 0) atomicMax(a.ctr + 4 + 2 * a.nwaves, ~sl_globaltimer());  // = min start time
      int mix_passed = 0;  // waves whose mix is known to be complete: 0 .. mix_passed - 1
      const unsigned long long pol_last = sl_policy_evict_last();
      while (true) {
        StackJob job{};
        bool done;
        if (pg == 0) {
          unsigned long long tk = 0;
          if (lane == 0) tk = atomicAdd(a.ctr, 1ull) - base;
          tk = __shfl_sync(0xffffffffu, tk, 0);
          done = tk >= (unsigned long long)a.njobs;
          if (!done) job = a.jobs[tk];
          if (!done && a.mix_lag > 0 && a.nmix > 0 && job.wave - a.mix_lag >= mix_passed && job.wave - a.mix_lag >= 1) {
            // back-pressure (experiments): the T blocks of wave w - lag must have been mixed before wave w is started
            const int need_wave = job.wave - a.mix_lag;
            if (lane == 0) {
              for (int w = mix_passed > 1 ? mix_passed : 1; w <= need_wave; ++w) {
                const unsigned long long need = a.epoch * (unsigned long long)a.wave_need[a.nwaves + w];
                unsigned spins = 0;
                while (sl_ld_acquire(a.ctr + 4 + a.nwaves + w) < need) {
                  __nanosleep(128);
                  if (++spins > (1u << 22)) {
                    atomicAdd(a.ctr + 2, 1ull);
                    break;
                  }
                }
              }
            }
            __syncwarp();
            mix_passed = need_wave + 1;
          }
          sl_mbar_wait(&jempty[jq], jphase ^ 1u);
          int* slot = job_slot + jq * SL_JOBWORDS;
          if (done) job.M = -1;  // stop record
          if (lane < SL_JOBWORDS) slot[lane] = reinterpret_cast<const int*>(&job)[lane];
          __syncwarp();
          if (lane == 0) sl_mbar_arrive(&jfull[jq]);
        } else {
          sl_mbar_wait(&jfull[jq], jphase);
          const int* slot = job_slot + jq * SL_JOBWORDS;
#pragma unroll
          for (int w = 0; w < SL_JOBWORDS; ++w) reinterpret_cast<int*>(&job)[w] = slot[w];
          __syncwarp();
          if (lane == 0) sl_mbar_arrive(&jempty[jq]);
          done = job.M < 0;
        }
        if (++jq == SL_JOBQ) {
          jq = 0;
          jphase ^= 1u;
        }
        if (done) break;
        const int K = job.K;
        // ---- A tiles of this group ----
        if (job.tmap >= 0) {
          const unsigned char* tm = a.tmaps + (size_t)job.tmap * 128;
          for (int m0 = pg * SL_TM; m0 < job.M; m0 += NG * SL_TM) {
            for (int k0 = 0; k0 < K; k0 += SL_KC) {
              sl_mbar_wait(&rg.empty[rg.stage], rg.phase ^ 1u);
              if (lane == 0) {
                sl_mbar_expect_tx(&rg.full[rg.stage], SL_STAGE_ELEMS * 8);
                if (dbg & 128)
                  sl_tma_2d_hint(rg.ring + rg.stage * SL_STAGE_ELEMS, tm, k0, job.arow + m0, &rg.full[rg.stage], pol_last);
                else if (!(dbg & 1))
                  sl_tma_2d(rg.ring + rg.stage * SL_STAGE_ELEMS, tm, k0, job.arow + m0, &rg.full[rg.stage]);
                else
                  asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(sl_smem(&rg.full[rg.stage])),
                               "r"(SL_STAGE_ELEMS * 8)
                               : "memory");
              } else {
                sl_mbar_arrive(&rg.full[rg.stage]);
              }
              rg.advance();
            }
          }
        } else {
          // lane = (row % 4, 16-byte piece of the 128-byte chunk row); piece q of row r lands at chunk q ^ (r & 7)
          const double* Ag = resolve(job.a_off, job.a_base, bases);
          const int arow = lane >> 3, aq = lane & 7;
          for (int m0 = pg * SL_TM; m0 < job.M; m0 += NG * SL_TM) {
            const int mt = min(SL_TM, job.M - m0);
            const double* At = Ag + (long long)(m0 + arow) * job.lda + 2 * aq;
            const long long step = 4ll * job.lda;
            const int nq = (mt - arow + 3) >> 2;  // rows arow + 4 i < mt
            for (int k0 = 0; k0 < K; k0 += SL_KC) {
              sl_mbar_wait(&rg.empty[rg.stage], rg.phase ^ 1u);
              double* st = rg.ring + rg.stage * SL_STAGE_ELEMS + arow * SL_KC;
              double* d0 = st + ((aq ^ arow) << 1);        // rows arow + 8 i     : r & 7 = arow
              double* d1 = st + ((aq ^ (arow + 4)) << 1);  // rows arow + 4 + 8 i : r & 7 = arow + 4
              int bytes = (K - (k0 + 2 * aq)) * 8;
              bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
              const double* src = bytes ? At + k0 : Ag;
              if (dbg & 1) {  // timing experiment 1: no operand loads
              } else {
#pragma unroll
                for (int i = 0; i < SL_TM / 4; ++i)
                  if (i < nq) sl_cp16(((i & 1) ? d1 : d0) + i * 4 * SL_KC, src + (bytes ? i * step : 0), bytes);
              }
              sl_cp_arrive(&rg.full[rg.stage]);
              rg.advance();
            }
          }
        }
      }
      if (a.nmix > 0 && NG == 1) {
        if (sl_mixer_loop<8>(a, bases, lane, nwarps_total, (dbg & 64) ? nullptr : smem_free))
          sl_mixer_loop_tma(a, bases, lane, nwarps_total, reinterpret_cast<unsigned char*>(sl_sm) + warp * SLM_SLOTS * SLM_SLOT_BYTES,
                            mixbar + warp * SLM_SLOTS);
      }
    } else if (a.nmix > 0) {
      // =========================== MIXERS (stage W) ===========================
      if (NG == 1) {
        if (sl_mixer_loop<8>(a, bases, lane, nwarps_total, (dbg & 64) ? nullptr : smem_free))
          sl_mixer_loop_tma(a, bases, lane, nwarps_total, reinterpret_cast<unsigned char*>(sl_sm) + warp * SLM_SLOTS * SLM_SLOT_BYTES,
                            mixbar + warp * SLM_SLOTS);
      } else if (dbg & 64) {
        sl_mixer_loop<8>(a, bases, lane, nwarps_total);
      } else {
        const int mw = warp - NCONS - NPROD;
        sl_mixer_loop_tma(a, bases, lane, nwarps_total, mixstage + mw * SLM_SLOTS * SLM_SLOT_BYTES, mixbar + mw * SLM_SLOTS);
      }
    }
  }
}

}  // namespace htn
