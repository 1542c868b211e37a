// Standalone check + timing of the stacked stage-L kernel (csrc/htn_stackl.cuh) on dense stand-ins of the C4 sector shapes:
//   for each (M, K, column blocks): T_j = A (M x K) . X_j (K x n_j), jobs of `tiles` 64-row tiles x <= 56 columns.
// Verifies against a naive kernel and prints TF/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I include -I hubbardtn_b200/csrc -o build/stackl_probe tools/stackl_probe.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "htn_stackl.cuh"

using namespace htn;

__global__ void naive(const double* A, int lda, const double* B, int ldb, double* C, int ldc, int M, int N, int K) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= N) return;
  double s = 0;
  for (int k = 0; k < K; ++k) s += A[(long long)i * lda + k] * B[(long long)k * ldb + j];
  C[(long long)i * ldc + j] = s;
}
__global__ void fillr(double* p, long long n, unsigned seed) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)i * 2654435761u + seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = ((double)(h & 0xFFFFF) / 1048576.0 - 0.5);
}
__global__ void maxdiff(const double* a, const double* b, long long n, double* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = fabs(a[i] - b[i]);
  if (d > 1e-9) atomicAdd((unsigned long long*)(out + 1), 1ULL);
  unsigned long long* o = (unsigned long long*)out;
  unsigned long long v = __double_as_longlong(d);
  atomicMax(o, v);
}

struct Shape { int M, K; std::vector<int> cols; };

static int even_up_(int x) { return (x + 1) & ~1; }

int main(int argc, char** argv) {
  const int tiles_per_job = argc > 1 ? atoi(argv[1]) : 8;
  const int ntmax = argc > 2 ? atoi(argv[2]) : 56;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  // the six heavy sectors of the C4 stage L (tools output of /tmp stats: M_l, K=n_l, n_r of the x blocks)
  std::vector<Shape> shapes = {{10963, 154, {109, 109, 210}}, {16549, 167, {120, 120, 210, 94}}, {15252, 149, {210, 53, 109, 120}},
                               {15696, 149, {53, 210, 109, 120}}, {10606, 67, {94, 23, 120, 27}}, {10969, 67, {23, 94, 120, 27}},
                               {6486, 38, {27, 27, 94, 12}}, {10251, 42, {120, 7, 53, 23}}};
  const int use_tma = argc > 5 ? atoi(argv[5]) : 1;
  PFN_cuTensorMapEncodeTiled encode = nullptr;
  {
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    encode = (PFN_cuTensorMapEncodeTiled)fn;
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
  }
  std::vector<CUtensorMap> maps;
  std::vector<StackJob> jobs;
  std::vector<double*> dA, dX, dT, dR;
  std::vector<long long> tsize;
  double flops = 0;
  struct Ref { int s, j; };
  for (size_t s = 0; s < shapes.size(); ++s) {
    const Shape& sh = shapes[s];
    const int lda = even_up_(sh.K);
    double* A;
    cudaMalloc(&A, (size_t)sh.M * lda * 8);
    fillr<<<(unsigned)(((long long)sh.M * lda + 255) / 256), 256>>>(A, (long long)sh.M * lda, 17 + s);
    dA.push_back(A);
    {
      CUtensorMap m;
      cuuint64_t dims[2] = {(cuuint64_t)sh.K, (cuuint64_t)sh.M};
      cuuint64_t strides[1] = {(cuuint64_t)lda * 8};
      cuuint32_t box[2] = {16, 64};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("tensor map encode failed: %d\n", (int)r); return 4; }
      maps.push_back(m);
    }
    for (size_t j = 0; j < sh.cols.size(); ++j) {
      const int n = sh.cols[j], ld = even_up_(n);
      double *X, *T, *R;
      cudaMalloc(&X, (size_t)sh.K * ld * 8);
      cudaMalloc(&T, (size_t)sh.M * ld * 8);
      cudaMalloc(&R, (size_t)sh.M * ld * 8);
      cudaMemset(T, 0, (size_t)sh.M * ld * 8);
      cudaMemset(R, 0, (size_t)sh.M * ld * 8);
      fillr<<<(unsigned)(((long long)sh.K * ld + 255) / 256), 256>>>(X, (long long)sh.K * ld, 1000 + 10 * s + j);
      dX.push_back(X); dT.push_back(T); dR.push_back(R); tsize.push_back((long long)sh.M * ld);
      naive<<<dim3((n + 127) / 128, sh.M), 128>>>(A, lda, X, ld, R, ld, sh.M, n, sh.K);
      flops += 2.0 * sh.M * n * sh.K;
      // N tiles: split the 8-column atoms evenly into pieces of <= ntmax columns
      const int atoms = (n + 7) / 8, maxat = ntmax / 8, npieces = (atoms + maxat - 1) / maxat;
      int c0 = 0;
      for (int p = 0; p < npieces; ++p) {
        const int at = atoms / npieces + (p < atoms % npieces ? 1 : 0);
        const int nt = std::min(at * 8, n - c0);
        if (!stack_job_fits(sh.K, nt)) { printf("job does not fit K=%d nt=%d\n", sh.K, nt); return 1; }
        for (int m0 = 0, step = tiles_per_job * SL_TM; m0 < sh.M; m0 += step) {
          // the last quarter of every row range is cut into smaller jobs (dynamic scheduling evens out the tail)
          step = (sh.M - m0 > 4 * tiles_per_job * SL_TM || tiles_per_job <= 2) ? tiles_per_job * SL_TM
                                                                                : std::max(2, tiles_per_job / 4) * SL_TM;
          StackJob jb{};
          jb.a_off = (long long)(A + (long long)m0 * lda); jb.a_base = 0;
          jb.b_off = (long long)(X + c0); jb.b_base = 0;
          jb.c_off = (long long)(T + (long long)m0 * ld + c0); jb.c_base = 0;
          jb.lda = lda; jb.ldb = ld; jb.ldc = ld; jb.K = sh.K; jb.nt = nt; jb.nb = std::min(even_up_(nt), ld - c0);
          jb.M = std::min(step, sh.M - m0);
          jb.tmap = use_tma ? (int)s : -1;
          jb.arow = m0;
          jb.wave = -1;
          jobs.push_back(jb);
        }
        c0 += nt;
      }
    }
  }
  // dynamic schedule: descending cost; the tail of every row range is cut finer (see tiles_per_job loop) by the caller
  const int per_sm = 2, G = sms * per_sm;
  auto cost = [&](const StackJob& j) { return (double)((j.M + 63) / 64) * (((j.K + 15) / 16) * (((j.nt + 7) / 8) + 0.75) + 6.0) + 20.0; };
  std::stable_sort(jobs.begin(), jobs.end(), [&](const StackJob& a, const StackJob& b) { return cost(a) > cost(b); });
  std::vector<StackJob>& table = jobs;
  printf("jobs %zu, flops %.3f GF\n", jobs.size(), flops / 1e9);
  StackJob* dj;
  cudaMalloc(&dj, table.size() * sizeof(StackJob));
  cudaMemcpy(dj, table.data(), table.size() * sizeof(StackJob), cudaMemcpyHostToDevice);
  Bases bases{};
  unsigned char* dmaps;
  cudaMalloc(&dmaps, maps.size() * sizeof(CUtensorMap));
  cudaMemcpy(dmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  unsigned long long* ctr;
  cudaMalloc(&ctr, 64 * 8);
  cudaMemset(ctr, 0, 64 * 8);
  StackArgs sa{};
  sa.jobs = dj;
  sa.njobs = (int)table.size();
  sa.tmaps = dmaps;
  sa.nmix = 0;
  sa.ctr = ctr;
  sa.epoch = 0;
  cudaFuncSetAttribute(stack_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SL_SMEM_BYTES);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stack_gemm_kernel, SL_THREADS, SL_SMEM_BYTES);
  printf("occupancy %d CTAs/SM, smem %d B\n", occ, SL_SMEM_BYTES);
  sa.epoch++;
  sa.dbg = 0;
  stack_gemm_kernel<<<G, SL_THREADS, SL_SMEM_BYTES>>>(sa, bases);
  cudaError_t e = cudaDeviceSynchronize();
  printf("first run: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  double* dm;
  cudaMalloc(&dm, 16);
  double worst = 0;
  for (size_t i = 0; i < dT.size(); ++i) {
    cudaMemset(dm, 0, 16);
    maxdiff<<<(unsigned)((tsize[i] + 255) / 256), 256>>>(dT[i], dR[i], tsize[i], dm);
    double h[2];
    cudaMemcpy(h, dm, 16, cudaMemcpyDeviceToHost);
    unsigned long long bad = *(unsigned long long*)&h[1];
    worst = std::max(worst, h[0]);
    if (bad) printf("  array %zu: %llu elements differ by > 1e-9 (max %.3e)\n", i, bad, h[0]);
  }
  printf("max abs difference vs naive: %.3e\n", worst);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int dbg = argc > 4 ? atoi(argv[4]) : 0;
  float best = 1e30f, sum = 0;
  const int reps = 20;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    sa.epoch++;
    sa.dbg = dbg;
    stack_gemm_kernel<<<G, SL_THREADS, SL_SMEM_BYTES>>>(sa, bases);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
    sum += ms;
  }
  printf("tma %d dbg %d tiles/job %d ntmax %d : best %.4f ms (%.2f TF/s), mean %.4f ms (%.2f TF/s)   status %s\n", use_tma, dbg, tiles_per_job, ntmax, best,
         flops / (best * 1e-3) / 1e12, sum / reps, flops / (sum / reps * 1e-3) / 1e12, cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
