// Micro-benchmark: latency of one grid-wide reduction of 3 doubles per PCG iteration of the on-chip kernel
// (k_pcg_patch), for the schemes considered in round 2:
//   fx     the production scheme: 96-bit fixed-point partials added with red.add.u64 to HF_NREP replica lines,
//          arrival counts in the low byte, warp 0 polls (hf_persist.cuh)
//   a2a    all-to-all packets: every CTA stores its 3 partials as two self-validating 16-byte packets in its own
//          slot, every CTA polls all G slots and adds them in slot order (no atomics, one store->load trip)
//   cl<k>  clusters of k CTAs: partials go to the cluster leader through DSMEM (st.shared::cluster + remote
//          mbarrier arrive), the G/k leaders run the a2a exchange, totals return through DSMEM
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/ubench_gridred tools/ubench_gridred.cu
// Run  : tools/bin/ubench_gridred [iterations]      (prints us per iteration for each scheme)
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../heatflow_b200/csrc/hf_persist.cuh"

namespace cg = cooperative_groups;

thread_local std::string hf_err_msg;
int hf_fail(int code, const std::string& msg) {
  fprintf(stderr, "error %d: %s\n", code, msg.c_str());
  return code;
}

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__);  \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ double fake_work(double x, int n) {
  for (int i = 0; i < n; ++i) x = fma(x, 0.999999, 1e-9);
  return x;
}

// ---- production scheme ------------------------------------------------------------------------------
__global__ void __launch_bounds__(HF_PT, 1) k_fx(int nit, int work, unsigned long long* acc, unsigned long long* acc_prev,
                                                 unsigned* genp, int* fail, double* out) {
  extern __shared__ double red[];
  const int G = gridDim.x;
  unsigned gen = *genp;
  FxState fx;
  hf_fx_load_state(fx, acc_prev);
  double x = 1.0 + threadIdx.x * 1e-3 + blockIdx.x, s = 0.0;
  __syncthreads();
  for (int it = 0; it < nit; ++it) {
    ++gen;
    x = fake_work(x, work);
    double d[3] = {x * 1e-3, x * 2e-3, x * 3e-3};
    const int eb[3] = {40, 40, 40};
    hf_fx_arrive<3>(d, eb, acc, gen, red, fail);
    double tot[3];
    hf_fx_wait<3>(tot, eb, acc, G, gen, red, fx, fail);
    s += tot[0] + tot[1] + tot[2];
    x = 1.0 + 1e-12 * tot[0];
    __syncthreads();
  }
  hf_fx_store_state(fx, acc_prev);
  if (threadIdx.x == 0) {
    out[blockIdx.x] = s;
    if (blockIdx.x == 0) *genp = gen;
  }
}

// ---- all-to-all packets ------------------------------------------------------------------------------
// slot = 2 x uint4: {v0.lo, v0.hi, v1.lo, gen} {v1.hi, v2.lo, v2.hi, gen}
#define SLOT_U4 8   // uint4 per slot: one 128-byte line
__device__ __forceinline__ void st_pkt(uint4* p, unsigned a, unsigned b, unsigned c, unsigned gen) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(gen) : "memory");
}

// every thread t < n_slots polls slot t; values land in sv[3][n_slots]; then warps 0..2 add them in slot order
__device__ __forceinline__ void a2a_exchange(const double (&mine)[3], int my_slot, int n_slots, uint4* slots, unsigned gen,
                                             double* sv, double* s_tot, bool publisher) {
  uint4* set = slots + (size_t)(gen & 1u) * (HF_MAX_GRID + 1) * SLOT_U4;
  if (publisher && threadIdx.x == 0) {
    const unsigned a0 = __double2loint(mine[0]), a1 = __double2hiint(mine[0]), b0 = __double2loint(mine[1]),
                   b1 = __double2hiint(mine[1]), c0 = __double2loint(mine[2]), c1 = __double2hiint(mine[2]);
    st_pkt(set + (size_t)my_slot * SLOT_U4, a0, a1, b0, gen);
    st_pkt(set + (size_t)my_slot * SLOT_U4 + 1, b1, c0, c1, gen);
  }
  const int t = threadIdx.x;
  if (t < n_slots) {
    uint4 p0, p1;
    const uint4* q = set + (size_t)t * SLOT_U4;
    do {
      p0 = hf_pkt_load(q);
      p1 = hf_pkt_load(q + 1);
    } while (p0.w != gen || p1.w != gen);
    sv[t] = __hiloint2double((int)p0.y, (int)p0.x);
    sv[HF_MAX_GRID + t] = __hiloint2double((int)p1.x, (int)p0.z);
    sv[2 * HF_MAX_GRID + t] = __hiloint2double((int)p1.z, (int)p1.y);
  }
  __syncthreads();
  const int lane = t & 31, warp = t >> 5;
  if (warp < 3) {
    double s = 0.0;
    for (int i = lane; i < n_slots; i += 32) s += sv[warp * HF_MAX_GRID + i];
    s = hf_warp_sum(s);
    if (lane == 0) s_tot[warp] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void block_partials(const double (&d)[3], double* red, double (&t3)[3]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double t = hf_warp_sum(d[i]);
    if (lane == 0) red[warp * 3 + i] = t;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < HF_PW; ++w) s += red[w * 3 + i];
    t3[i] = s;
  }
}

__global__ void __launch_bounds__(HF_PT, 1) k_a2a(int nit, int work, uint4* slots, unsigned* genp, double* out) {
  extern __shared__ double sm[];
  double* red = sm;                       // [HF_PW*3]
  double* sv = sm + 32;                   // [3][HF_MAX_GRID]
  double* s_tot = sv + 3 * HF_MAX_GRID;   // [3]
  const int G = gridDim.x;
  unsigned gen = *genp;
  double x = 1.0 + threadIdx.x * 1e-3 + blockIdx.x, s = 0.0;
  for (int it = 0; it < nit; ++it) {
    ++gen;
    x = fake_work(x, work);
    double d[3] = {x * 1e-3, x * 2e-3, x * 3e-3}, t3[3];
    block_partials(d, red, t3);
    a2a_exchange(t3, blockIdx.x, G, slots, gen, sv, s_tot, true);
    s += s_tot[0] + s_tot[1] + s_tot[2];
    x = 1.0 + 1e-12 * s_tot[0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[blockIdx.x] = s;
    if (blockIdx.x == 0) *genp = gen;
  }
}

// ---- clusters + a2a among the leaders -----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote_f64(unsigned raddr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}
__device__ __forceinline__ void arrive_remote(unsigned rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
__device__ __forceinline__ void wait_cluster(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W1:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D1;\n"
      "bra W1;\n"
      "D1:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int CS>
__global__ void __launch_bounds__(HF_PT, 1) k_cluster(int nit, int work, uint4* slots, unsigned* genp, double* out) {
  extern __shared__ double sm[];
  double* red = sm;                        // [32]
  double* sv = sm + 32;                    // [3][HF_MAX_GRID]
  double* s_tot = sv + 3 * HF_MAX_GRID;    // [4]
  double* s_part = s_tot + 4;              // [2][CS][4]   leader: members' partials
  double* s_bc = s_part + 2 * CS * 4;      // [2][4]       member: totals from the leader
  __shared__ __align__(8) unsigned long long barA, barB;
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  const int cid = blockIdx.x / CS, NC = gridDim.x / CS;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&barA)), "r"(CS > 1 ? CS - 1 : 1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&barB)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl.sync();
  unsigned gen = *genp;
  double x = 1.0 + threadIdx.x * 1e-3 + blockIdx.x, s = 0.0;
  for (int it = 0; it < nit; ++it) {
    ++gen;
    const unsigned par = (unsigned)it & 1u;
    x = fake_work(x, work);
    double d[3] = {x * 1e-3, x * 2e-3, x * 3e-3}, t3[3];
    block_partials(d, red, t3);
    if (rank != 0) {
      if (threadIdx.x == 0) {
        const unsigned base = mapa(smem_u32(s_part + (par * CS + rank) * 4), 0);
#pragma unroll
        for (int i = 0; i < 3; ++i) st_remote_f64(base + 8 * i, t3[i]);
        arrive_remote(mapa(smem_u32(&barA), 0));
        wait_cluster(&barB, par);
        s_tot[0] = s_bc[par * 4 + 0];
        s_tot[1] = s_bc[par * 4 + 1];
        s_tot[2] = s_bc[par * 4 + 2];
      }
      __syncthreads();
    } else {
      if (CS > 1) {
        if (threadIdx.x == 0) {
          wait_cluster(&barA, par);
#pragma unroll
          for (int r = 1; r < CS; ++r)
#pragma unroll
            for (int i = 0; i < 3; ++i) t3[i] += s_part[(par * CS + r) * 4 + i];
          red[24] = t3[0];
          red[25] = t3[1];
          red[26] = t3[2];
        }
        __syncthreads();
        t3[0] = red[24];
        t3[1] = red[25];
        t3[2] = red[26];
      }
      a2a_exchange(t3, cid, NC, slots, gen, sv, s_tot, true);
      if (threadIdx.x > 0 && threadIdx.x < CS) {
        const unsigned r = threadIdx.x;
        const unsigned base = mapa(smem_u32(s_bc + par * 4), r);
#pragma unroll
        for (int i = 0; i < 3; ++i) st_remote_f64(base + 8 * i, s_tot[i]);
        arrive_remote(mapa(smem_u32(&barB), r));
      }
    }
    s += s_tot[0] + s_tot[1] + s_tot[2];
    x = 1.0 + 1e-12 * s_tot[0];
    __syncthreads();
  }
  cl.sync();
  if (threadIdx.x == 0) {
    out[blockIdx.x] = s;
    if (blockIdx.x == 0) *genp = gen;
  }
}

template <typename F>
static float timed(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int CS>
static void run_cluster(int nit, int work, int smem, uint4* slots, unsigned* gen, double* out, int sm_count) {
  auto fn = k_cluster<CS>;
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (CS > 8) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(HF_PT);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  cfg.gridDim = dim3(CS);
  int ncl = 0;
  CK(cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg));
  int G = std::min(ncl * CS, (HF_MAX_GRID / CS) * CS);
  G = std::min(G, (sm_count / CS) * CS);
  cfg.gridDim = dim3(G);
  for (int w : {work, 0}) {
    const float ms = timed([&] { CK(cudaLaunchKernelEx(&cfg, fn, nit, w, slots, gen, out)); }, 3);
    printf("cl%-2d   G=%3d (max clusters %d) work=%3d : %.3f us / iteration\n", CS, G, ncl, w, ms * 1e3 / nit);
  }
}

static double first_out(double* out) {
  double h = 0.0;
  CK(cudaMemcpy(&h, out, sizeof(double), cudaMemcpyDeviceToHost));
  return h;
}

int main(int argc, char** argv) {
  const int nit = argc > 1 ? atoi(argv[1]) : 4000;
  const int work = 150;   // dependent fp64 FMAs per iteration (~600 cycles): stands for the SpMV + updates
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %d iterations per launch\n", prop.name, sms, nit);
  const int smem = 120 * 1024;   // > half an SM: one CTA per SM, as in k_pcg_patch<.,.,1>
  unsigned long long *acc, *acc_prev;
  unsigned* gen;
  int* fail;
  double* out;
  uint4* slots;
  CK(cudaMalloc(&acc, sizeof(unsigned long long) * 2 * HF_NREP * HF_ACC_LINE));
  CK(cudaMalloc(&acc_prev, sizeof(unsigned long long) * 2 * HF_NREP * HF_ACC_LINE));
  CK(cudaMemset(acc, 0, sizeof(unsigned long long) * 2 * HF_NREP * HF_ACC_LINE));
  CK(cudaMemset(acc_prev, 0, sizeof(unsigned long long) * 2 * HF_NREP * HF_ACC_LINE));
  CK(cudaMalloc(&gen, 4));
  CK(cudaMemset(gen, 0, 4));
  CK(cudaMalloc(&fail, 4));
  CK(cudaMemset(fail, 0, 4));
  CK(cudaMalloc(&out, sizeof(double) * 1024));
  CK(cudaMalloc(&slots, sizeof(uint4) * 2 * (HF_MAX_GRID + 1) * SLOT_U4));
  CK(cudaMemset(slots, 0, sizeof(uint4) * 2 * (HF_MAX_GRID + 1) * SLOT_U4));
  CK(cudaFuncSetAttribute(k_fx, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_a2a, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int G : {138, 37}) {
    for (int w : {work, 0}) {
      int nit_ = nit, w_ = w;
      void* a1[] = {&nit_, &w_, &acc, &acc_prev, &gen, &fail, &out};
      float ms = timed([&] { CK(cudaLaunchCooperativeKernel((const void*)k_fx, dim3(G), dim3(HF_PT), a1, smem, 0)); }, 3);
      printf("fx     G=%3d work=%3d : %.3f us / iteration  (HF_NREP %d, poll delay %d)  check %.15e\n", G, w, ms * 1e3 / nit, HF_NREP,
             HF_POLL_DELAY, first_out(out));
      void* a2[] = {&nit_, &w_, &slots, &gen, &out};
      ms = timed([&] { CK(cudaLaunchCooperativeKernel((const void*)k_a2a, dim3(G), dim3(HF_PT), a2, smem, 0)); }, 3);
      printf("a2a    G=%3d work=%3d : %.3f us / iteration\n", G, w, ms * 1e3 / nit);
    }
  }
  run_cluster<1>(nit, work, smem, slots, gen, out, sms);
  run_cluster<2>(nit, work, smem, slots, gen, out, sms);
  run_cluster<4>(nit, work, smem, slots, gen, out, sms);
  run_cluster<8>(nit, work, smem, slots, gen, out, sms);
  run_cluster<16>(nit, work, smem, slots, gen, out, sms);
  int hfail = 0;
  CK(cudaMemcpy(&hfail, fail, 4, cudaMemcpyDeviceToHost));
  printf("fx failures: %d\n", hfail);
  return 0;
}
