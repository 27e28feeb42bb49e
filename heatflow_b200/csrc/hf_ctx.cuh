// heatflow_b200 - shared context, device buffers and reduction helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/heatflow_b200.h"

#define HF_BLOCK 256          // threads per CTA of the streaming kernels (8 warps)
#define HF_SLICE 32           // rows per sliced-ELL slice = one warp, one row per lane
#define HF_MAX_PART 2368      // max CTAs of a reduction-producing kernel (148 SMs x 16)
#ifndef HF_IT
#define HF_IT 256             // threads per CTA of the streaming PCG kernel (hf_pcg.cu)
#endif
#ifndef HF_IT_MINB
#define HF_IT_MINB 3           // its CTAs per SM
#endif

extern thread_local std::string hf_err_msg;
int hf_fail(int code, const std::string& msg);

#define HF_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return hf_fail(HF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
  } while (0)

#define HF_TRY(call)                 \
  do {                               \
    int rc__ = (call);               \
    if (rc__ != HF_OK) return rc__;  \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  // zero-filled allocation; the fill is ordered on stream `s` (the context stream is
  // non-blocking, so a legacy-stream cudaMemset could land after later kernels on `s`)
  int alloc(size_t count, cudaStream_t s) {
    release();
    n = count;
    if (count == 0) return HF_OK;
    HF_CUDA(cudaMalloc(&p, count * sizeof(T)));
    HF_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
    return HF_OK;
  }
  int upload(const T* h, size_t count, cudaStream_t s) {
    if (count != n || !p) HF_TRY(alloc(count, s));
    if (count) HF_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return HF_OK;
  }
  int download(T* h, size_t count, cudaStream_t s) const {
    if (count > n) return hf_fail(HF_ERR_ARG, "download larger than buffer");
    if (count) HF_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    HF_CUDA(cudaStreamSynchronize(s));
    return HF_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Device-side control block of one PCG solve.  Every CTA reduces the partial sums itself (fixed
// order => bit-reproducible), so alpha/beta/convergence never touch the host.
struct HfCtrl {
  double thr;        // rtol^2 * ||b_free||^2 in the Jacobi-scaled norm
  double bn2;        // ||b_free||^2
  double rr;         // ||r||^2 when the solve stopped
  int done;          // set by the SpMV kernel once rr <= thr
  int itA;           // iteration counter read by the SpMV kernel (written by the update kernel)
  int itB;           // copy read by the update kernel (written by the SpMV kernel)
  int nparts;        // number of valid partial sums (= grid of the producing kernels)
  double alpha;      // streaming kernel: step length of the iteration just finished
  double beta;       //                   direction update factor of the next iteration
  unsigned counter;  //                   CTAs that have published their partial sums
  unsigned pad_;
  double part_rr[2][HF_MAX_PART];
  double part_pq[HF_MAX_PART];
  double part_bn[HF_MAX_PART];
};

// Sliced-ELL view of a Jacobi-scaled symmetric operator  D^-1/2 A D^-1/2.
struct SellView {
  int nslices;
  const int* __restrict__ slice_ptr;   // [nslices+1] element offsets
  const int* __restrict__ col;         // [slice_ptr[nslices]]
  const double* __restrict__ val;
};

// Patch view of the same operator for the streaming kernel: rows are cut into chunks of R
// consecutive rows (one CTA each).  A chunk's columns are its own rows plus a short halo list, so
// column indices are 16-bit positions in the CTA's shared-memory copy of the direction vector:
// [0, R) = own rows, R + h = halo_idx[halo_ptr[chunk] + h].
struct PatchView {
  int nslices, R, nchunks;
  const int* __restrict__ slice_ptr;
  const unsigned short* __restrict__ lcol;
  const double* __restrict__ val;
  const int* __restrict__ halo_ptr;    // [nchunks+1]
  const int* __restrict__ halo_idx;
};

struct SellOp {
  int nslices = 0;
  size_t padded_nnz = 0;
  DevBuf<int> slice_ptr, col;
  DevBuf<double> val;
  DevBuf<double> scale;                // s_i = sqrt(diag) (1 on Dirichlet rows)
  mutable cudaGraphExec_t chunk_exec[3] = {nullptr, nullptr, nullptr};   // captured PCG iteration chunks
  int st_grid = 0;                     // co-resident grid of the persistent streaming kernel (hf_stream.cu); 0 => not eligible
  // plan of the patch-based persistent kernel (hf_patch.cu); pp_rpt == 0 => not eligible
  int pp_rpt = 0, pp_grid = 0, pp_mat_cap = 0, pp_halo_cap = 0, pp_share = 1;
  size_t pp_smem = 0;
  DevBuf<unsigned short> pp_lcol;      // local columns for chunks of 256 * pp_rpt rows, sliced-ELL order
  DevBuf<int> pp_halo_ptr, pp_halo_idx;
  DevBuf<unsigned char> pp_pub;        // rows whose q packet some other CTA reads
  const SellOp* plan_from = nullptr;   // patch plan borrowed from an operator with the same sparsity pattern
  // patch decomposition (streaming kernel)
  int R = 0, nchunks = 0, halo_max = 0, mat_cap = 0, halo_cap = 0, nstages = 0;
  size_t stage_bytes = 0;
  DevBuf<unsigned short> lcol;
  DevBuf<unsigned short> lcol_csr;     // the same local columns in CSR slot order (input of the value fill)
  bool struct_valid = false;           // slices / patches / plans match the current mesh (values may be refilled)
  DevBuf<int> halo_ptr, halo_idx;
  size_t iter_smem = 0;
  SellView view() const { return SellView{nslices, slice_ptr.p, col.p, val.p}; }
  PatchView patch() const { return PatchView{nslices, R, nchunks, slice_ptr.p, lcol.p, val.p, halo_ptr.p, halo_idx.p}; }
  void drop_graphs() const {
    for (auto& g : chunk_exec) {
      if (g) cudaGraphExecDestroy(g);
      g = nullptr;
    }
  }
  ~SellOp() { drop_graphs(); }
};

struct PcgWork {
  DevBuf<double> x, r, p0, p1, q;
  DevBuf<double> r1, q1;               // second halves of the ping-pong pairs (streaming kernel)
  DevBuf<double> parts;                // [4][max chunks] per-CTA partial sums (streaming kernel)
  DevBuf<HfCtrl> ctrl;
  DevBuf<unsigned long long> acc, acc_prev;   // fixed-point reduction accumulators (persistent kernel)
  DevBuf<uint4> qpk;                   // [2][Npad] q = A p exchange packets (persistent kernel)
  DevBuf<double> sparts;               // [2][4][grid] per-CTA partial sums of the persistent streaming kernel
  DevBuf<unsigned long long> bar;      // its grid barrier: arrival counter, value at the end of the last launch
  DevBuf<unsigned> gen;                // barrier generation, monotonic across launches
  DevBuf<int> step_iters;              // per-step iteration counts written by the persistent kernel
  DevBuf<int> fail;                    // number of solves that hit max_iters
  HfCtrl* h_ctrl = nullptr;            // pinned mirror of the control-block header
  int grid = 0;
};

// Initial guess from the previous solves (hf_recycle.cu): Ahat-orthogonal corrections of up to `cap`
// solves, kept as W and inv[k] = 1 / (w_k . Ahat w_k); frozen once full.  The correction of the last solve waits
// in (d, ad = Ahat d) until the next projection orthogonalises and finalises it; inc = W c of the last projection.
struct Recycle {
  int cap = 0, count = 0, nseg = 0, nn_parts = 0;
  bool pending = false;
  size_t ld = 0;                       // row stride of W / AW (Npad rounded up to the dot-kernel segment)
  DevBuf<double> W, inv, coef, hn, parts, part_nn, d, ad, inc, x0;
};

struct EnsState;

struct hf_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  // mesh
  int N = 0, E = 0, nv = 3, Npad = 0;
  // internal node numbering (locality order for the streaming kernels): internal = rank[user],
  // user = order[internal].  Every C-ABI array is in the caller's numbering; the library permutes.
  int ordering_req = 0;                // 0 auto, 1 as given, 2 Hilbert curve
  bool permuted = false;
  std::vector<int> h_rank, h_order;
  DevBuf<int> rank_d, order_d;
  DevBuf<double> stage;                // [2N] staging buffer of permuted host transfers
  DevBuf<double> xy;
  DevBuf<int> cells, cell_tag;
  DevBuf<int> n2c_ptr, n2c_idx;
  std::vector<int> h_rowptr, h_col;
  DevBuf<int> rowptr, col;
  int64_t nnz = 0;
  // materials
  std::vector<int> mat_tags;
  std::vector<double> mat_kappa, mat_rhoc;
  DevBuf<double> cm, ck;               // per-cell coefficients fed to the assembly kernel
  DevBuf<int> sc_tags, sc_flag;        // set-up scratch (material table, error flag), reused across calls
  DevBuf<double> sc_kap, sc_rc;
  // boundary conditions
  int n_bc = 0, n_gauss = 0;
  DevBuf<int> bc_dofs, gauss_dof;
  DevBuf<double> gauss_r;
  DevBuf<unsigned char> bcflag;
  DevBuf<double> gfull;                // g on Dirichlet dofs, 0 elsewhere
  // operator
  bool op_built = false;
  bool op_transient = true;            // every material has rho c > 0 (mass term present): the pipelined CG kernel may be used
  double dt = 0.0;
  int axisym = 1;
  DevBuf<double> valM, valA0, valA, valM1;
  SellOp opA;
  // projection operator (built lazily)
  bool proj_built = false;
  SellOp opMr;
  DevBuf<double> proj_b, proj_g;
  // state
  DevBuf<double> u, uprev, b, source;
  bool have_prev = false, have_source = false;
  // solver
  double rtol = 1e-14, warm = 0.0;
  int share = 1;                       // hf_set_sharing: solves expected to run concurrently on this device (1 or 2)
  int max_iters = 20000, mode = 0, last_iters = 0;
  int debug_fx_shift = 0;              // hf_debug_fx_shift (tests): shrinks the fixed-point range of the on-chip reduction
  DevBuf<long long> debug_phase;       // hf_debug_phase_times: per-CTA phase clocks of the last on-chip solve
  int force_mode = -1;                 // >= 0: overrides `mode` (the retry of a failed single-launch run)
  unsigned long long stat_retries = 0; // runs repeated with the host-polled kernel after a failed single-launch solve
  DevBuf<double> u0_keep;              // [2N] u, uprev at the start of hf_run
  // counters (hf_get_stats)
  double stat_run_ms = 0.0, stat_relres = 0.0;
  unsigned long long stat_launches = 0, stat_iters = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // hf_set_profile: CUDA events around every PCG solve of hf_run (kernel time for the roofline numbers)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;
  double stat_solve_ms = 0.0;
  unsigned long long stat_solve_launches = 0;
  PcgWork ws;
  Recycle rc;
  DevBuf<double> hist;
  DevBuf<int> watch;
  DevBuf<unsigned char> flush;         // L2 flush buffer for hf_bench_kernels
  EnsState* ens = nullptr;
};

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double hf_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA, result returned to every thread; fixed order => deterministic.
// `sh` must hold HF_BLOCK/32 doubles; may be reused after the call returns.
__device__ __forceinline__ double hf_block_sum(double v, double* sh) {
  v = hf_warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < HF_BLOCK / 32; ++i) t += sh[i];
  return t;
}

// Sum of n partial sums stored in global memory, same value in every thread of every CTA.
__device__ __forceinline__ double hf_sum_parts(const double* part, int n, double* sh) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += HF_BLOCK) v += __ldcg(part + i);
  return hf_block_sum(v, sh);
}

__device__ __forceinline__ double hf_ld_stream(const double* p) {
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int hf_ld_stream(const unsigned short* p) {
  unsigned short v;
  asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return (int)v;
}
__device__ __forceinline__ int hf_ld_stream(const int* p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// shared between translation units
int hf_pcg_alloc(hf_ctx* c);
int hf_pcg_solve(hf_ctx* c, const SellOp& op, int* iters_out, double* relres_out);
int hf_patch_plan(hf_ctx* c, SellOp& op);
int hf_stream_plan(hf_ctx* c, SellOp& op);
int hf_stream_solve_async(hf_ctx* c, const SellOp& op, int step_slot, bool sum_parts = false);
int hf_patch_solve_async(hf_ctx* c, const SellOp& op, int step_slot, bool sum_parts = false);
bool hf_patch_pipelined(const hf_ctx* c, const SellOp& op);
int hf_build_sell(hf_ctx* c, const DevBuf<double>& csr_val, bool apply_bc, SellOp& op, DevBuf<double>* val_bc_out);
int hf_assemble_values(hf_ctx* c, const double* cm, const double* ck, int axisym, double* out);
void hf_ens_free(hf_ctx* c);
void hf_rc_reset(hf_ctx* c);
int hf_rc_project(hf_ctx* c);
int hf_rc_store(hf_ctx* c, const SellOp& op);
int hf_build_patches(const hf_ctx* c, int R, std::vector<int>& halo_ptr, std::vector<int>& halo_idx,
                     std::vector<unsigned short>& lcol, int* halo_max);
int hf_upload_nodal(hf_ctx* c, const double* h_user, double* d_internal);
int hf_download_nodal(hf_ctx* c, const double* d_internal, double* h_user, int ncomp);
int hf_internal_nodes(hf_ctx* c, int n, const int32_t* user_nodes, std::vector<int>& out);
