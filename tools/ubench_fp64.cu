// Microbenchmark: FP64 FMA and shared-memory gather throughput of one SM at the occupancy of the on-chip PCG kernels
// (1 CTA of 256 or 512 threads per SM).   nvcc -O3 -arch=sm_100a -o tools/bin/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  const double m = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  __syncthreads();
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// gather: every lane reads idx[...] from shared memory; VEC = 1 (8 B) or 2 (16 B)
template <int VEC>
__global__ void k_gather(const int* idx, int n_idx, double* out, int iters, long long* cyc) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096 * VEC; i += blockDim.x) sm[i] = i;
  int my[16];
  for (int k = 0; k < 16; ++k) my[k] = idx[(threadIdx.x * 16 + k) % n_idx];
  __syncthreads();
  double a0 = 0, a1 = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (VEC == 1) {
        a0 += sm[(my[k] + it * 8) & 4095];
      } else {
        const double2 v = reinterpret_cast<const double2*>(sm)[(my[k] + it * 8) & 4095];
        a0 += v.x;
        a1 += v.y;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  const int iters = 2000;
  for (int threads : {256, 512, 1024}) {
#define RUN(ILP)                                                                                             \
  k_dfma<ILP><<<148, threads>>>(out, iters, cyc);                                                           \
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);                                                      \
  printf("DFMA threads=%d ILP=%d: %.2f lanes/clk/SM (%.1f cycles)\n", threads, ILP, (double)threads * ILP * iters / h[0], (double)h[0]);
    RUN(1) RUN(2) RUN(4) RUN(8) RUN(16)
  }
  int* idx;
  int hidx[65536];
  cudaMalloc(&idx, sizeof hidx);
  for (int mode = 0; mode < 3; ++mode) {
    unsigned s = 12345;
    for (int i = 0; i < 65536; ++i) {
      s = s * 1664525u + 1013904223u;
      const int t = i / 16, k = i % 16;
      hidx[i] = mode == 0 ? (int)((s >> 8) % 4096) : mode == 1 ? (t + k * 37) % 4096 : (t * 5 + (int)((s >> 8) % 64)) % 4096;
    }
    cudaMemcpy(idx, hidx, sizeof hidx, cudaMemcpyHostToDevice);
    for (int threads : {256, 512}) {
      cudaFuncSetAttribute(k_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      cudaFuncSetAttribute(k_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      k_gather<1><<<148, threads, 65536>>>(idx, 65536, out, iters, cyc);
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      printf("gather 8B  mode=%d threads=%d: %.2f bytes/clk/SM\n", mode, threads, (double)threads * 16 * iters * 8 / h[0]);
      k_gather<2><<<148, threads, 65536>>>(idx, 65536, out, iters, cyc);
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      printf("gather 16B mode=%d threads=%d: %.2f bytes/clk/SM\n", mode, threads, (double)threads * 16 * iters * 16 / h[0]);
    }
  }
  printf("modes: 0 random, 1 consecutive lanes -> consecutive elements, 2 locally random (window of 64)\n");
  return 0;
}
