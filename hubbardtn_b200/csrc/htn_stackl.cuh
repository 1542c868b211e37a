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
constexpr int SL_NCONS = SL_TM / 16;           // consumer warps: 16 rows each
constexpr int SL_NMIX = 3;                     // mixer warps (stage W), warps SL_NCONS+1 ..
constexpr int SL_THREADS = (SL_NCONS + 1 + SL_NMIX) * 32;
constexpr int SL_JOBWORDS = (int)(sizeof(StackJob) / 4);
constexpr int SL_MIXSRC = 32;                  // source descriptors staged per mixer warp
// 2 CTAs per SM need 2 x (this + 1 KB) <= 228 KB: the slab leaves no room for per-warp staging of the mix sources
// (they travel through warp shuffles instead)
constexpr int SL_SMEM_BYTES = (SL_STAGES * SL_STAGE_ELEMS + SL_SLAB) * 8 + (2 * SL_STAGES + 4) * 8 + 2 * SL_JOBWORDS * 4 + 16;
static_assert(2 * (SL_SMEM_BYTES + 1024) <= 233472, "two CTAs per SM must fit");

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

// consumer side of one job, specialised on the number of 8-column atoms of the slab
template <int CA>
__device__ __forceinline__ void sl_consume_job(const StackJob& job, SlRing& rg, const double* __restrict__ slab, int role,
                                               int lane, const Bases& bases, int dbg) {
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
  for (int m0 = 0; m0 < job.M; m0 += SL_TM) {
    double acc[CA][2][2];
#pragma unroll
    for (int j = 0; j < CA; ++j) acc[j][0][0] = acc[j][0][1] = acc[j][1][0] = acc[j][1][1] = 0.0;
    const bool last_tile = m0 + SL_TM >= job.M;
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
__device__ __forceinline__ unsigned long long sl_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
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
__device__ __forceinline__ void sl_mixer_loop_t(const StackArgs& a, const Bases& bases, int lane, int nwarps_total) {
  const unsigned long long base = (a.epoch - 1ull) * (unsigned long long)(a.nmix + nwarps_total);
  long long c_desc = 0, c_wait = 0, c_work = 0, n_chunks = 0;
  while (true) {
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
    if (!(a.dbg & 4)) sl_mix_chunk<U>(a, ch, bases, lane);
    if (a.mix_lag > 0 && ch.pad_ >= 0) {  // back-pressure experiments: count the finished chunks of the wave
      __syncwarp();
      if (lane == 0) atomicAdd(a.ctr + 4 + a.nwaves + ch.pad_, 1ull);
    }
    if (PROF) {
      const long long t3 = clock64();
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
}
template <int U>
__device__ __forceinline__ void sl_mixer_loop_prof(const StackArgs& a, const Bases& bases, int lane, int nwarps_total) {
  sl_mixer_loop_t<U, true>(a, bases, lane, nwarps_total);
}
template <int U>
__device__ __forceinline__ void sl_mixer_loop(const StackArgs& a, const Bases& bases, int lane, int nwarps_total) {
  if (a.dbg & 32)
    sl_mixer_loop_prof<U>(a, bases, lane, nwarps_total);
  else
    sl_mixer_loop_t<U, false>(a, bases, lane, nwarps_total);
}

// Roles: warps 0 .. SL_NCONS-1 = consumers (DMMA; they also load the slab of a job themselves, all 128 threads,
// while the A chunks the producer has already queued wait in the ring: a job switch costs one slab latency);
// warp SL_NCONS = producer (job tickets, job records, the A ring); warps SL_NCONS+1 .. = mixers (stage W).
// Register budget: launched with 128 per thread (2 CTAs x 256 threads); the consumer warp group grows to 168 (with 124
// the DMMA loops lose a third of their speed: fewer fragment loads in flight), the other warp group shrinks to 88,
// enough for 12 independent 16-byte loads per mixer lane.
constexpr int SL_JOBQ = 2;
__global__ void __launch_bounds__(SL_THREADS, 2) stack_gemm_kernel(const __grid_constant__ StackArgs a,
                                                                    const __grid_constant__ Bases bases) {
  extern __shared__ __align__(1024) double sl_sm[];
  SlRing rg;
  rg.ring = sl_sm;  // 1024-byte aligned stages (128-byte swizzle atom = 8 rows x 128 B)
  double* slab = sl_sm + SL_STAGES * SL_STAGE_ELEMS;
  rg.full = reinterpret_cast<uint64_t*>(slab + SL_SLAB);
  rg.empty = rg.full + SL_STAGES;
  uint64_t* jfull = rg.empty + SL_STAGES;
  uint64_t* jempty = jfull + SL_JOBQ;
  int* job_slot = reinterpret_cast<int*>(jempty + SL_JOBQ);  // SL_JOBQ records
  rg.stage = 0;
  rg.phase = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dbg = a.dbg;
  // the ring must never hold NaN patterns: K tails multiply stale columns by the zero rows of the slab
  for (int i = tid; i < SL_STAGES * SL_STAGE_ELEMS; i += SL_THREADS) rg.ring[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < SL_STAGES; ++s) {
      sl_mbar_init(&rg.full[s], 32);          // every producer lane arrives once (cp.async noinc, or plain / expect_tx)
      sl_mbar_init(&rg.empty[s], SL_NCONS);   // one elected lane per consumer warp
    }
    for (int s = 0; s < SL_JOBQ; ++s) {
      sl_mbar_init(&jfull[s], 1);
      sl_mbar_init(&jempty[s], SL_NCONS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic zero fill before async-proxy (TMA) writes
  __syncthreads();
  int jq = 0;
  unsigned jphase = 0;

  if (warp < SL_NCONS) {
    // =========================== CONSUMERS ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;\n");
    while (true) {
      sl_mbar_wait(&jfull[jq], jphase);
      StackJob job;
      const int* slot = job_slot + jq * SL_JOBWORDS;
#pragma unroll
      for (int w = 0; w < SL_JOBWORDS; ++w) reinterpret_cast<int*>(&job)[w] = slot[w];
      if (job.M < 0) break;
      // ---- slab: K rows of nb doubles, rows K .. round4(K) zero; loaded by all consumer threads ----
      asm volatile("bar.sync 1, %0;\n" ::"n"(SL_NCONS * 32) : "memory");  // every warp has left the previous slab
      {
        const int CA = (job.nt + 7) >> 3, SB = CA * 8 + 4, K = job.K;
        const double* Bg = resolve(job.b_off, job.b_base, bases);
        const int np = job.nb >> 1, total = K * np;  // 16-byte pieces per row
        constexpr int NT_ = SL_NCONS * 32;
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
      asm volatile("bar.sync 1, %0;\n" ::"n"(SL_NCONS * 32) : "memory");
      const int role = warp;
      switch ((job.nt + 7) >> 3) {
        case 1: sl_consume_job<1>(job, rg, slab, role, lane, bases, dbg); break;
        case 2: sl_consume_job<2>(job, rg, slab, role, lane, bases, dbg); break;
        case 3: sl_consume_job<3>(job, rg, slab, role, lane, bases, dbg); break;
        case 4: sl_consume_job<4>(job, rg, slab, role, lane, bases, dbg); break;
        case 5: sl_consume_job<5>(job, rg, slab, role, lane, bases, dbg); break;
        case 6: sl_consume_job<6>(job, rg, slab, role, lane, bases, dbg); break;
        case 7: sl_consume_job<7>(job, rg, slab, role, lane, bases, dbg); break;
        default: sl_consume_job<8>(job, rg, slab, role, lane, bases, dbg); break;
      }
      __syncwarp();
      if (lane == 0) {
        sl_mbar_arrive(&jempty[jq]);
        if (job.wave >= 0) {  // this warp's strips of the job's T tiles are written: publish them to the mixers
          __threadfence();
          atomicAdd(a.ctr + 4 + job.wave, 1ull);
        }
      }
      if (++jq == SL_JOBQ) {
        jq = 0;
        jphase ^= 1u;
      }
    }
    if ((dbg & 32) && lane == 0) atomicMax(a.ctr + 4 + 2 * a.nwaves + 1, sl_globaltimer());
    // no stack jobs left: help with the mix (the last wave's targets are still to be formed)
    if (a.nmix > 0) sl_mixer_loop<8>(a, bases, lane, (int)gridDim.x * (SL_NMIX + SL_NCONS + 1));
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 112;\n");
    if (warp == SL_NCONS) {
      // =========================== PRODUCER ===========================
      const unsigned long long base = (a.epoch - 1ull) * (unsigned long long)(a.njobs + (int)gridDim.x);
      if ((dbg & 32) && lane == 0) atomicMax(a.ctr + 4 + 2 * a.nwaves, ~sl_globaltimer());  // = min start time
      int mix_passed = 0;  // waves whose mix is known to be complete: 0 .. mix_passed - 1
      while (true) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(a.ctr, 1ull) - base;
        tk = __shfl_sync(0xffffffffu, tk, 0);
        const bool done = tk >= (unsigned long long)a.njobs;
        StackJob job{};
        if (!done) job = a.jobs[tk];
        if (!done && a.mix_lag > 0 && a.nmix > 0 && job.wave - a.mix_lag >= mix_passed && job.wave - a.mix_lag >= 1) {
          // back-pressure: the T blocks of wave w - lag must have been mixed before wave w is started, so that what the
          // mixers read is still in L2 (without it the DMMA warps run ahead and T makes a round trip through HBM)
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
        if (++jq == SL_JOBQ) {
          jq = 0;
          jphase ^= 1u;
        }
        if (done) break;
        const int K = job.K;
        // ---- A tiles ----
        if (job.tmap >= 0) {
          const unsigned char* tm = a.tmaps + (size_t)job.tmap * 128;
          for (int m0 = 0; m0 < job.M; m0 += SL_TM) {
            for (int k0 = 0; k0 < K; k0 += SL_KC) {
              sl_mbar_wait(&rg.empty[rg.stage], rg.phase ^ 1u);
              if (lane == 0) {
                sl_mbar_expect_tx(&rg.full[rg.stage], SL_STAGE_ELEMS * 8);
                if (!(dbg & 1))
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
          for (int m0 = 0; m0 < job.M; m0 += SL_TM) {
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
      if (a.nmix > 0) sl_mixer_loop<8>(a, bases, lane, (int)gridDim.x * (SL_NMIX + SL_NCONS + 1));
    } else if (a.nmix > 0) {
      // =========================== MIXERS (stage W) ===========================
      sl_mixer_loop<8>(a, bases, lane, (int)gridDim.x * (SL_NMIX + SL_NCONS + 1));
    }
  }
}

}  // namespace htn
