// DMMA.8x8x4 issue-rate microbenchmarks with realistic operand patterns (sm_100a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: all DMMAs share a,b. MODE 1: c[i][f] += a[i]*b[f], i-major (kernel order). MODE 2: f-major.
// MODE 3: 4x4 pattern i-major.  MODE 4: like 1 but operands reloaded from shared memory each k-step.
template <int MODE, int NI, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters, const double* in) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
  __syncthreads();
  double c[NI][NF][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int f = 0; f < NF; ++f) c[i][f][0] = c[i][f][1] = 0.0;
  double a[NI], b[NF];
#pragma unroll
  for (int i = 0; i < NI; ++i) a[i] = in[threadIdx.x + 32 * i];
#pragma unroll
  for (int f = 0; f < NF; ++f) b[f] = in[threadIdx.x + 512 + 32 * f];
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < NI; ++i) a[i] = sm[((it & 3) * 8 + i) * 32 + lane];
#pragma unroll
      for (int f = 0; f < NF; ++f) b[f] = sm[512 + ((it & 3) * 4 + f) * 32 + lane];
    }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma(c[i][f][0], c[i][f][1], a[0], b[0]);
    } else if (MODE == 2) {
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int i = 0; i < NI; ++i) dmma(c[i][f][0], c[i][f][1], a[i], b[f]);
    } else {
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma(c[i][f][0], c[i][f][1], a[i], b[f]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int f = 0; f < NF; ++f) s += c[i][f][0] + c[i][f][1];
  if (s == 123.456) out[0] = s;
}
template <int MODE, int NI, int NF>
void run(const char* name, int warps_per_sm, double* d, const double* in) {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sm = p.multiProcessorCount, iters = 4096;
  int ctas = warps_per_sm >= 8 ? warps_per_sm / 8 : 1, thr = warps_per_sm >= 8 ? 256 : warps_per_sm * 32;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<MODE, NI, NF><<<sm * ctas, thr>>>(d, iters, in);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = (double)sm * ctas * (thr / 32) * iters * NI * NF * 512.0 / (ms * 1e-3) / 1e12;
    if (r && tf > best) best = tf;
  }
  printf("%-34s warps/SM %2d : %6.2f TFLOP/s\n", name, warps_per_sm, best);
}
int main() {
  double *d, *in;
  cudaMalloc(&d, 8);
  cudaMalloc(&in, 1024 * 8);
  cudaMemset(in, 0, 1024 * 8);
  for (int w : {4, 8, 16, 64}) {
    run<0, 7, 2>("same operands (7x2 acc)", w, d, in);
    run<1, 7, 2>("a[i]*b[f] i-major 7x2", w, d, in);
    run<2, 7, 2>("a[i]*b[f] f-major 7x2", w, d, in);
    run<1, 4, 4>("a[i]*b[f] 4x4", w, d, in);
    run<1, 8, 2>("a[i]*b[f] i-major 8x2", w, d, in);
    run<4, 7, 2>("7x2 operands from smem each step", w, d, in);
  }
  return 0;
}
