// Throughput probe: cp.reduce.async.bulk.global.shared::cta.add.f64 (SASS UBLKRED.G.S.ADD.F64) issued from every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/blkred_probe tools/blkred_probe.cu
// Each warp owns a staging strip in shared memory and issues `rows` bulk reductions of `bytes` each per "tile" into
// pseudo-random row-pitched destinations of a large buffer (the U workspace of the H_AC apply: 250 MB).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) red_kernel(double* buf, long long nelem, int bytes, int rows, int pitch_elems, int tiles,
                                                  int mode, int depth) {
  extern __shared__ __align__(128) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* strip = sm + (size_t)warp * 2 * rows * (bytes / 8);  // two strips per warp (double buffering)
  for (int i = lane; i < 2 * rows * (bytes / 8); i += 32) strip[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncwarp();
  unsigned long long h = (blockIdx.x * 4 + warp) * 0x9E3779B97F4A7C15ull + 12345;
  const long long span = (long long)rows * pitch_elems;
  for (int t = 0; t < tiles; ++t) {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    long long off = (long long)((h >> 20) % (unsigned long long)(nelem - span - 64));
    off &= ~1ll;  // 16-byte aligned
    double* s = strip + (size_t)(t & 1) * rows * (bytes / 8);
    if (mode == 2) {  // what the epilogue would do: rewrite the strip (generic proxy) before handing it to the TMA
      for (int i = lane; i < rows * (bytes / 8); i += 32) s[i] = 1.0 + t;
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
    }
    if (lane == 0) {
      for (int r = 0; r < rows; ++r) {
        if (mode == 0 || mode == 2)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;\n" ::"l"(buf + off + (long long)r * pitch_elems),
                       "r"(smem_u32(s + (size_t)r * (bytes / 8))), "r"(bytes)
                       : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(buf + off + (long long)r * pitch_elems),
                       "r"(smem_u32(s + (size_t)r * (bytes / 8))), "r"(bytes)
                       : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
}

int main(int argc, char** argv) {
  const long long nelem = 250ll * 1000 * 1000 / 8;
  double* buf;
  cudaMalloc(&buf, nelem * 8);
  cudaMemset(buf, 0, nelem * 8);
  cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  struct Cfg { int bytes, rows, mode, ctas; };
  std::vector<Cfg> cfgs;
  for (int mode : {0, 1, 2})
    for (int ctas : {1, 2})
      for (auto br : {std::pair<int,int>{128, 16}, {448, 16}, {1344, 8}, {4096, 2}, {7168, 1}})
        cfgs.push_back(Cfg{br.first, br.second, mode, ctas});
  for (const Cfg& c : cfgs) {
    const int tiles = 400;
    const size_t smem = (size_t)4 * 2 * c.rows * c.bytes;
    const int grid = 148 * c.ctas;
    red_kernel<<<grid, 128, smem>>>(buf, nelem, c.bytes, c.rows, 168, 20, c.mode, 1);
    cudaEventRecord(e0);
    red_kernel<<<grid, 128, smem>>>(buf, nelem, c.bytes, c.rows, 168, tiles, c.mode, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 4 * tiles * c.rows, by = ops * c.bytes;
    printf("mode %d (%s) ctas/SM %d  op %5d B x %2d rows : %.3f ms  %.2f Mops/s/SM (%.1f clk/op/SM @1.9GHz)  %.2f TB/s   %s\n", c.mode,
           c.mode == 1 ? "bulk store" : (c.mode == 0 ? "bulk red.add.f64" : "red.add.f64 + strip rewrite"), c.ctas, c.bytes, c.rows, ms,
           ops / 148 / ms / 1e3, 1.9e9 * ms * 1e-3 * 148 / ops, by / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
