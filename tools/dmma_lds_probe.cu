// Does interleaving shared-memory fragment loads with DMMA cost tensor throughput?  Mimics the consumer
// inner loop of the grouped GEMM (per k4-step: 8 A-fragment LDS.64 + 2 B-fragment LDS.64 + 16 DMMA) with
// w warps per SM sub-partition, no barriers, no global traffic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/dmma_lds_probe tools/dmma_lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <bool WITH_LDS>
__global__ void __launch_bounds__(128) k(double* out, int iters) {
  __shared__ double As[64 * 20];
  __shared__ double Bs[16 * 68];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) As[i] = 1.0 + 1e-9 * i;
  for (int i = threadIdx.x; i < 16 * 68; i += blockDim.x) Bs[i] = 1.0 - 1e-9 * i;
  __syncthreads();
  double acc[8][2][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
  const double* as = As + g * 20 + t;
  const double* bs = Bs + warp * 16 + t * 68 + g;
  double fx[8], ff[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) fx[i] = as[i * 160];
  ff[0] = bs[0];
  ff[1] = bs[8];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (WITH_LDS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) fx[i] = as[i * 160 + kk * 4];
        ff[0] = bs[kk * 4 * 68];
        ff[1] = bs[kk * 4 * 68 + 8];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dmma884(acc[i][0][0], acc[i][0][1], fx[i], ff[0]);
        dmma884(acc[i][1][0], acc[i][1][1], fx[i], ff[1]);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0][0] + acc[i][0][1] + acc[i][1][0] + acc[i][1][1];
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2048;
  for (int lds = 0; lds < 2; ++lds)
    for (int ctas = 1; ctas <= 4; ++ctas) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (lds)
          k<true><<<sms * ctas, 128>>>(out, iters);
        else
          k<false><<<sms * ctas, 128>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("%s, %d warps per sub-partition: %.2f TF/s\n", lds ? "DMMA + fragment LDS" : "DMMA only          ", ctas,
             (double)sms * ctas * 4 * iters * 64.0 * 512.0 / (best * 1e-3) / 1e12);
    }
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
