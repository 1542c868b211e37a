// What register tile / fragment-load scheme lets a smem-fed DMMA.8x8x4 loop reach the issue-loop peak?
// Consumer-only model of a GEMM main loop (operands already in shared memory, no producer, no global traffic):
// every warp owns an RA x CA grid of 8x8 atoms; per k4-step it loads RA A-fragments and CA B-fragments and
// issues RA*CA DMMAs.  Variants:
//   MODE 0  LDS.64 per fragment, loads of step k issued right before its DMMAs
//   MODE 1  LDS.64, fragments of step k+1 prefetched into a second register set (software pipelining)
//   MODE 2  LDS.128 k-pair fragments (A[row][2t,2t+1], B^T[col][2t,2t+1]) : one load feeds two k4-steps
//   MODE 3  MODE 2 + an mbarrier try_wait (already complete) every 16 k, as a ring consumer would do
//   MODE 4  MODE 1 + the same try_wait
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_tile_probe tools/dmma_tile_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
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

template <int RA, int CA, int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) k(double* out, int iters) {
  // A: per warp RA*8 rows x 16 k.  stride 20 (LDS.64) / 24 (LDS.128) doubles: conflict free.
  // B: 16 k x CA*8 cols (+4 pad) for LDS.64 ; B^T: CA*8 cols x 16 k, stride 24, for LDS.128
  extern __shared__ double sm[];
  constexpr bool PAIR = (MODE == 2 || MODE == 3);
  constexpr int SA = PAIR ? 24 : 20;
  constexpr int SB = PAIR ? 24 : CA * 8 + 4;
  double* As = sm;
  double* Bs = sm + NWARPS * RA * 8 * SA;
  __shared__ uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int na = NWARPS * RA * 8 * SA, nb = PAIR ? CA * 8 * SB : 16 * SB;
  for (int i = threadIdx.x; i < na; i += blockDim.x) As[i] = 1.0 + 1e-9 * i;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) Bs[i] = 1.0 - 1e-9 * i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // complete phase 0 once: waits on parity 0 succeed from now on
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  double acc[RA][CA][2];
#pragma unroll
  for (int i = 0; i < RA; ++i)
#pragma unroll
    for (int j = 0; j < CA; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  if (!PAIR) {
    const double* as = As + warp * RA * 8 * SA + g * SA + t;
    const double* bs = Bs + t * SB + g;
    double fa[2][RA], fb[2][CA];
    auto load = [&](int buf, int kk) {
#pragma unroll
      for (int i = 0; i < RA; ++i) fa[buf][i] = as[i * 8 * SA + kk * 4];
#pragma unroll
      for (int j = 0; j < CA; ++j) fb[buf][j] = bs[kk * 4 * SB + j * 8];
    };
    auto mma = [&](int buf) {
#pragma unroll
      for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < CA; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[buf][i], fb[buf][j]);
    };
    if (MODE == 0) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          load(0, kk);
          mma(0);
        }
      }
    } else {
      load(0, 0);
      for (int it = 0; it < iters; ++it) {
        if (MODE == 4) mbar_wait(&bar, 0);
        load(1, 1);
        mma(0);
        load(0, 2);
        mma(1);
        load(1, 3);
        mma(0);
        load(0, 0);  // first step of the next chunk
        mma(1);
      }
    }
  } else {
    const double2* as = reinterpret_cast<const double2*>(As + warp * RA * 8 * SA + g * SA + 2 * t);
    const double2* bs = reinterpret_cast<const double2*>(Bs + g * SB + 2 * t);
    double2 fa[2][RA], fb[2][CA];
    auto load = [&](int buf, int kp) {  // kp: pair index 0/1 (k 0..7 / 8..15)
#pragma unroll
      for (int i = 0; i < RA; ++i) fa[buf][i] = as[(i * 8 * SA + kp * 8) / 2];
#pragma unroll
      for (int j = 0; j < CA; ++j) fb[buf][j] = bs[(j * 8 * SB + kp * 8) / 2];
    };
    auto mma = [&](int buf) {
#pragma unroll
      for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < CA; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[buf][i].x, fb[buf][j].x);
#pragma unroll
      for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < CA; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[buf][i].y, fb[buf][j].y);
    };
    load(0, 0);
    for (int it = 0; it < iters; ++it) {
      if (MODE == 3) mbar_wait(&bar, 0);
      load(1, 1);
      mma(0);
      load(0, 0);
      mma(1);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < RA; ++i)
#pragma unroll
    for (int j = 0; j < CA; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int RA, int CA, int MODE, int NWARPS>
void run(double* out, int sms) {
  constexpr bool PAIR = (MODE == 2 || MODE == 3);
  const int SA = PAIR ? 24 : 20, SB = PAIR ? 24 : CA * 8 + 4;
  const int smem = (NWARPS * RA * 8 * SA + (PAIR ? CA * 8 * SB : 16 * SB)) * 8;
  auto kern = k<RA, CA, MODE, NWARPS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2048;
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    kern<<<sms, NWARPS * 32, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  printf("tile %dx%d atoms  mode %d  warps/SM %2d  regs %3d : %6.2f TF/s  (%s)\n", RA, CA, MODE, NWARPS, fa.numRegs,
         (double)sms * NWARPS * iters * 4.0 * RA * CA * 512.0 / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
}

template <int RA, int CA, int NWARPS>
void run_modes(double* out, int sms) {
  run<RA, CA, 0, NWARPS>(out, sms);
  run<RA, CA, 1, NWARPS>(out, sms);
  run<RA, CA, 2, NWARPS>(out, sms);
  run<RA, CA, 3, NWARPS>(out, sms);
  run<RA, CA, 4, NWARPS>(out, sms);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, 8);
  run_modes<2, 7, 4>(out, sms);
  run_modes<2, 7, 8>(out, sms);
  run_modes<2, 7, 12>(out, sms);
  run_modes<2, 8, 4>(out, sms);
  run_modes<2, 8, 8>(out, sms);
  run_modes<4, 4, 4>(out, sms);
  run_modes<4, 4, 8>(out, sms);
  run_modes<4, 4, 12>(out, sms);
  run_modes<4, 6, 4>(out, sms);
  run_modes<4, 6, 8>(out, sms);
  run_modes<4, 7, 4>(out, sms);
  run_modes<4, 7, 8>(out, sms);
  run_modes<4, 8, 4>(out, sms);
  run_modes<4, 8, 8>(out, sms);
  run_modes<8, 4, 4>(out, sms);
  run_modes<8, 4, 8>(out, sms);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
