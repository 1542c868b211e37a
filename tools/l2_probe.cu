// Probes used to bound the grouped GEMM: (1) L2 -> SM read bandwidth with the access pattern of the
// GEMM producer (16-byte cp.async.cg into shared memory), as a function of the footprint;
// (2) DMMA.8x8x4 throughput versus the number of issuing warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2_probe tools/l2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}

__global__ void __launch_bounds__(256) read_kernel(const double* __restrict__ buf, long long n16, int iters, double* out) {
  extern __shared__ double sm[];
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    for (long long j = i; j < n16; j += stride * 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        long long k = j + u * stride;
        if (k < n16) cp_async16(sm + ((threadIdx.x + u * 256) & 2047) * 2, buf + k * 2);
      }
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 2;\n" ::);
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  if (sm[threadIdx.x] == 123.456) out[0] = 1.0;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(128) dmma_kernel(double* out, int iters) {
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

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const long long sizes_mb[] = {64};
  for (long long mb : sizes_mb) {
    const long long bytes = mb << 20;
    double* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    const long long n16 = bytes / 16;
    const int iters = (int)((8LL << 30) / bytes) + 1;
    for (int ctas = 2; ctas <= 6; ctas += 2) {
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        read_kernel<<<sms * ctas, 256, 32768>>>(buf, n16, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("read footprint %5lld MB  %d CTAs/SM: %.2f TB/s\n", mb, ctas, (double)bytes * iters / (best * 1e-3) / 1e12);
    }
    cudaFree(buf);
  }
  // one warp per sub-partition and CTA (128 threads); `ctas` CTAs per SM => `ctas` warps per sub-partition
  for (int nacc = 4; nacc <= 16; nacc *= 2)
    for (int ctas = 1; ctas <= 16; ctas *= 2) {
      float best = 1e30f;
      const int iters = 4096;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (nacc == 4) dmma_kernel<4><<<sms * ctas, 128>>>(out, iters);
        if (nacc == 8) dmma_kernel<8><<<sms * ctas, 128>>>(out, iters);
        if (nacc == 16) dmma_kernel<16><<<sms * ctas, 128>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("DMMA %2d warps per sub-partition, %2d independent accumulators: %.2f TF/s\n", ctas, nacc,
             (double)sms * ctas * 4 * iters * nacc * 512.0 / (best * 1e-3) / 1e12);
    }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
