// FP64 roofline denominators of this repo (MEASURED_PEAKS.json has no FP64 entry): DMMA.8x8x4 issue loop,
// DFMA issue loop and cuBLAS DGEMM at 4096^3 and 8192^3, each timed with CUDA events (best and mean of the
// repetitions).  Prints one JSON object; run through tools/run_fp64_peak.sh, which adds the clock record and writes
// profiles/fp64_peak.json.  cuBLAS is linked by THIS TOOL only (libhtn.so links no library).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak_probe tools/fp64_peak_probe.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

struct Stat {
  double best, mean;
};

template <class F>
Stat timeit(F f, double flops, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();  // warm-up
  cudaDeviceSynchronize();
  double best = 0, sum = 0;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = flops / (ms * 1e-3) / 1e12;
    best = tf > best ? tf : best;
    sum += tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return Stat{best, sum / reps};
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 8192, grid = sms * 8;
  Stat dm = timeit([&] { dmma_loop<<<grid, 256>>>(d, iters); }, (double)grid * 8 * iters * 8.0 * 512.0, 10);
  Stat df = timeit([&] { dfma_loop<<<grid, 256>>>(d, iters); }, (double)grid * 256 * iters * 16.0 * 2.0, 10);
  cublasHandle_t h;
  cublasCreate(&h);
  printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"dmma_loop_tflops\": {\"best\": %.3f, \"mean\": %.3f}, \"dfma_loop_tflops\": {\"best\": %.3f, \"mean\": %.3f}",
         p.name, sms, dm.best, dm.mean, df.best, df.mean);
  for (int n : {4096, 8192}) {
    double *A, *B, *C;
    const size_t bytes = (size_t)n * n * 8;
    cudaMalloc(&A, bytes);
    cudaMalloc(&B, bytes);
    cudaMalloc(&C, bytes);
    std::vector<double> hst((size_t)n * n);
    for (size_t i = 0; i < hst.size(); ++i) hst[i] = (double)((i * 2654435761u) & 1023) / 1024.0 - 0.5;
    cudaMemcpy(A, hst.data(), bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(B, hst.data(), bytes, cudaMemcpyHostToDevice);
    const double one = 1.0, zero = 0.0;
    Stat s = timeit([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); },
                    2.0 * n * n * n, n == 4096 ? 10 : 5);
    printf(", \"cublas_dgemm_%d_tflops\": {\"best\": %.3f, \"mean\": %.3f}", n, s.best, s.mean);
    cudaFree(A);
    cudaFree(B);
    cudaFree(C);
  }
  printf(", \"status\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
