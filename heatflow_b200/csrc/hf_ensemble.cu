// heatflow_b200 - batched multi-RHS ensemble for parameter_sweep (sm_100a).
//
// Replaces the reference's one-process-per-parameter-set loop (parameter_sweep.py:123-192,
// :436-438): B variants of one simulation - different sample conductivity k_s and Gaussian
// heating width fwhm_s, same mesh, same materials otherwise - advance together.
//
//   operator of variant s :  A0_s = base0 + k_s * S0      (two value arrays on the shared pattern)
//       base0 = M(rho_c) + dt K(kappa with kappa_sample = 0),  S0 = dt K(1 on sample cells, else 0)
//   Dirichlet rows/cols are masked on the fly (bcflag), so the same two arrays serve the RHS
//   (un-BC'd operator, apply_lifting) and the solve (BC'd operator).
//   vectors               :  [N, B] with the variant index fastest, so the gather of column j
//                            reads B contiguous doubles and the matrix entry is shared by B lanes.
//   thread mapping        :  one thread per (row, variant); a warp covers 32/B consecutive rows.
//
// Jacobi-PCG per variant with its own alpha/beta/convergence (converged variants are frozen with
// alpha = beta = 0), in the Jacobi-scaled space of each variant (shat_b = 1/sqrt(diag A_b)), with the
// same one-launch-per-iteration structure as the single-simulation streaming kernel (hf_pcg.cu):
// k_ens_iter, kernel n of a solve, on chunks of R rows x B variants (one CTA each):
//   phase 1  x += alpha p_old shat ; r_n = r - alpha q ; p_n = r_n + beta p_old   (own rows; halo rows
//            recomputed redundantly) ; shat * p_n -> shared memory [(R + halo), B]
//   phase 2  q_n = shat_i sum_j (base0_ij + k_b S0_ij) (shat_j p_n,j) from shared memory with 16-bit
//            local columns ; per-variant partials of r.r, p.q, r.q, q.q
//   tail     last CTA: rr_n direct -> convergence / alpha_n ; rr_{n+1} by the residual identity -> beta
// The "last CTA" adds the per-CTA partials in CTA order, so results are bit-reproducible.
// Algorithmic traffic per dof, iteration and variant (nnz ~ 7 / row, halo fraction h):
//   matrix (18*nnz + 4 + 1)/B (two value arrays + 16-bit columns, shared by the B variants)
//   vectors: read x r p q shat (40 + 32 h), write x r p q (32)          total ~ 72 + 32 h + 131/B bytes.
#include <algorithm>
#include <cmath>

#include "hf_ctx.cuh"

#define HF_EB 32            // max variants per tile
#define HF_ET 256           // threads per CTA

struct EnsCtrl {
  int done, it, n_active, pad;
  unsigned counter[4];
  int active[HF_EB];
  double thr[HF_EB], rz[HF_EB], alpha[HF_EB], beta[HF_EB], bn[HF_EB];
};

struct EnsState {
  int B = 0, LB = 0;        // tile width (power of two) and its log2
  int nb = 0;               // real variants in the tile (<= B)
  int grid = 0;
  int last_iters = 0;
  DevBuf<double> base0, s0;           // [nnz]
  DevBuf<double> ks, coeff;           // [B]
  DevBuf<double> dinv, g, u, x, r, r1, p0, p1, q, q1;   // [Nalloc*B]; dinv holds shat = 1/sqrt(diag)
  DevBuf<double> part;                // [4][max(grid, nchunks)][B]
  // patch decomposition for the iteration kernel
  int R = 0, nchunks = 0, halo_max = 0;
  size_t iter_smem = 0;
  DevBuf<int> halo_ptr, halo_idx;
  DevBuf<unsigned short> lcol;        // CSR slot order
  DevBuf<double2> bs;                 // {base0, S0} per CSR slot
  int mcap = 0;                       // max non-zeros of a chunk
  DevBuf<EnsCtrl> ctrl;
  EnsCtrl* h_ctrl = nullptr;          // pinned mirror
  DevBuf<double> hist, stage;
  DevBuf<int> watch;
  cudaGraphExec_t chunk_exec[3] = {nullptr, nullptr, nullptr};
  ~EnsState() {
    for (auto& g : chunk_exec)
      if (g) cudaGraphExecDestroy(g);
    if (h_ctrl) cudaFreeHost(h_ctrl);
  }
};

void hf_ens_free(hf_ctx* c) {
  delete c->ens;
  c->ens = nullptr;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
// Sum over all threads of the CTA that share the same variant (tid & (B-1)); the result for
// variant b lands in sh_out[b] (valid after the trailing __syncthreads()).
template <int LB, int NT = HF_ET>
__device__ __forceinline__ void ens_block_sum(double v, double* sh /*[NT/32][32]*/, double* sh_out /*[32]*/) {
  constexpr int B = 1 << LB;
#pragma unroll
  for (int o = 16; o >= B; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane < B) sh[w * 32 + lane] = v;
  __syncthreads();
  if (threadIdx.x < B) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += sh[i * 32 + threadIdx.x];
    sh_out[threadIdx.x] = t;
  }
  __syncthreads();
}

// Publishes this CTA's per-variant partials and elects the last CTA to finish.  Returns true in
// every thread of that CTA, after which part[0 .. gridDim.x) are all visible to it.
template <int LB>
__device__ __forceinline__ bool ens_publish(const double* sh_vals, double* part, unsigned* counter) {
  constexpr int B = 1 << LB;
  __shared__ int s_last;
  if (threadIdx.x < B) {
    __stcg(part + (size_t)blockIdx.x * B + threadIdx.x, sh_vals[threadIdx.x]);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;
    __threadfence();
  }
  __syncthreads();
  return s_last != 0;
}

// Fixed-order sum of the per-CTA partials, one result per variant in sh_out[b].
template <int LB, int NT = HF_ET>
__device__ __forceinline__ void ens_sum_parts(const double* part, int nparts, double* sh /*[NT]*/, double* sh_out) {
  constexpr int B = 1 << LB;
  constexpr int G = NT / B;             // groups of CTAs summed by different threads
  const int b = threadIdx.x & (B - 1), grp = threadIdx.x >> LB;
  double v = 0.0;
  for (int cta = grp; cta < nparts; cta += G) v += __ldcg(part + (size_t)cta * B + b);
  __syncthreads();
  sh[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x < B) {
    double t = 0.0;
    for (int k = 0; k < G; ++k) t += sh[k * B + threadIdx.x];
    sh_out[threadIdx.x] = t;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// per-variant Jacobi scaling: shat = 1 / sqrt(base0_ii + k_s S0_ii), 1 on Dirichlet rows
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_diag(int N, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ base0,
           const double* __restrict__ s0, const unsigned char* __restrict__ bcflag, const double* __restrict__ ks,
           double* __restrict__ dinv, int* __restrict__ bad) {
  constexpr int B = 1 << LB;
  const size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x;
  const int i = (int)(idx >> LB), b = (int)(idx & (B - 1));
  if (i >= N) return;
  double d = 1.0;
  if (!bcflag[i]) {
    d = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) d = fma(ks[b], s0[k], base0[k]);
    if (!(d > 0.0)) {
      atomicExch(bad, i + 1);
      d = 1.0;
    }
  }
  dinv[idx] = 1.0 / sqrt(d);
}

template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_bcast(int N, const double* __restrict__ src, double* __restrict__ dst) {
  const size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x;
  const int i = (int)(idx >> LB);
  if (i < N) dst[idx] = src[i];
}

// g_jb = (amp - t_ic) exp(coeff_b r_j^2) + t_ic on the Gaussian-profile dofs (bc.py:128-137,
// run_with_diamond.py:354-359 with the variant's own fwhm)
template <int LB>
__global__ void k_ens_gauss(int n, const int* __restrict__ dof, const double* __restrict__ r, double amp, double t_ic,
                            const double* __restrict__ coeff, double* __restrict__ g) {
  constexpr int B = 1 << LB;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = idx >> LB, b = idx & (B - 1);
  if (i < n) g[((size_t)dof[i] << LB) + b] = (amp - t_ic) * exp(coeff[b] * (r[i] * r[i])) + t_ic;
}

// assemble_vector + apply_lifting + set_bc + initial residual for all variants:
//   free row i : rhat = shat_i ((M u)_i - (A0_s x0)_i) with x0 = u on free columns, g on Dirichlet columns
//   bc row i   : x = g, rhat = 0
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_init(int N, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ valM,
           const double* __restrict__ base0, const double* __restrict__ s0, const unsigned char* __restrict__ bcflag,
           const double* __restrict__ ks, const double* __restrict__ g, const double* __restrict__ u,
           const double* __restrict__ shat, double* __restrict__ x, double* __restrict__ r,
           double* __restrict__ part, EnsCtrl* __restrict__ c, double rtol) {
  constexpr int B = 1 << LB;
  __shared__ double sh[HF_ET];
  __shared__ double s_bn[HF_EB];
  const int b = threadIdx.x & (B - 1);
  const double kb = ks[b];
  const size_t total = (size_t)N << LB;
  double l_bn = 0.0;
  for (size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x; idx < total; idx += (size_t)gridDim.x * HF_ET) {
    const int i = (int)(idx >> LB);
    double xv, rv = 0.0;
    if (bcflag[i]) {
      xv = g[idx];
    } else {
      double t1 = 0.0, t2 = 0.0, t3 = 0.0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const int j = col[k];
        const double a = fma(kb, s0[k], base0[k]);
        const double uj = u[((size_t)j << LB) + b];
        t1 = fma(valM[k], uj, t1);
        if (bcflag[j]) {
          const double gj = g[((size_t)j << LB) + b];
          t2 = fma(a, gj, t2);
          t3 = fma(a, gj, t3);
        } else {
          t2 = fma(a, uj, t2);
        }
      }
      const double si = shat[idx];
      xv = u[idx];
      rv = (t1 - t2) * si;
      const double bh = (t1 - t3) * si;
      l_bn = fma(bh, bh, l_bn);
    }
    x[idx] = xv;
    r[idx] = rv;
  }
  ens_block_sum<LB>(l_bn, sh, s_bn);
  if (ens_publish<LB>(s_bn, part, &c->counter[0])) {
    ens_sum_parts<LB>(part, gridDim.x, sh, s_bn);
    if (threadIdx.x < B) {
      c->thr[threadIdx.x] = rtol * rtol * s_bn[threadIdx.x];
      c->bn[threadIdx.x] = s_bn[threadIdx.x];
      c->rz[threadIdx.x] = 0.0;
      c->alpha[threadIdx.x] = 0.0;
      c->beta[threadIdx.x] = 0.0;
      c->active[threadIdx.x] = 1;
    }
    if (threadIdx.x == 0) {
      c->n_active = B;
      c->done = 0;
      c->it = 0;
    }
  }
}

struct EnsPatch {
  int R, nchunks, mcap;                      // rows per chunk, chunks, max non-zeros of a chunk
  const int* __restrict__ halo_ptr;
  const int* __restrict__ halo_idx;
  const unsigned short* __restrict__ lcol;   // CSR slot order
  const double2* __restrict__ bs;            // {base0, S0} per CSR slot
};

#define HF_EPAIRS 2048      // (row, variant) pairs per CTA: R = HF_EPAIRS / B rows
#define HF_ENT 512          // threads per CTA of the iteration kernel
#define HF_ERPT (HF_EPAIRS / HF_ENT)

__device__ __forceinline__ unsigned ens_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int LB>
__global__ void __launch_bounds__(HF_ENT, 2)
k_ens_iter(int N, EnsPatch P, int par, const int* __restrict__ rowptr, const unsigned char* __restrict__ bcflag,
           const double* __restrict__ ks, const double* __restrict__ shat, double* __restrict__ x,
           double* __restrict__ rb0, double* __restrict__ rb1, double* __restrict__ pb0, double* __restrict__ pb1,
           double* __restrict__ qb0, double* __restrict__ qb1, double* __restrict__ part, EnsCtrl* __restrict__ c) {
  constexpr int B = 1 << LB;
  constexpr int R = HF_EPAIRS / B;
  extern __shared__ __align__(128) unsigned char smraw[];
  double2* sbs = reinterpret_cast<double2*>(smraw);                 // [mcap] {base0, S0} of the chunk's rows
  double* sp = reinterpret_cast<double*>(sbs + P.mcap);              // shat * p_n, [(R + nh), B]
  __shared__ double sh[HF_ENT];
  __shared__ double s_tot[4][HF_EB];
  __shared__ double s_al[HF_EB], s_be[HF_EB];
  __shared__ int s_ctl[2];
  __shared__ int srow[HF_EPAIRS / 4 + 1];                           // chunk-local row pointers (R + 1 <= 513 used)
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const int b = tid & (B - 1);
  const int lo = blockIdx.x * R;
  const int hi = min(lo + R, N);
  const int hp0 = P.halo_ptr[blockIdx.x], nh = P.halo_ptr[blockIdx.x + 1] - hp0;
  unsigned short* slc = reinterpret_cast<unsigned short*>(sp + ((size_t)(R + nh) << LB));   // [mcap] local columns
  const double* __restrict__ ro = par ? rb1 : rb0;
  const double* __restrict__ po = par ? pb1 : pb0;
  const double* __restrict__ qo = par ? qb1 : qb0;
  double* __restrict__ rn = par ? rb0 : rb1;
  double* __restrict__ pn = par ? pb0 : pb1;
  double* __restrict__ qn = par ? qb0 : qb1;
  const int k0 = rowptr[lo < N ? lo : N], k1 = rowptr[hi];
  // ---- thread 0: the chunk's operator entries -> shared memory by TMA bulk copy
  if (tid < B) {                        // ONE reader per CTA and variant: the whole grid polls these L2 lines
    s_al[tid] = *(volatile double*)&c->alpha[tid];
    s_be[tid] = *(volatile double*)&c->beta[tid];
  }
  if (tid == 0) {
    const int d = *(volatile int*)&c->done;
    s_ctl[0] = d;
    s_ctl[1] = *(volatile int*)&c->it;
    if (d == 0) {
      const unsigned bytes = (unsigned)(k1 - k0) * 16u;
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ens_smem_u32(&mbar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ens_smem_u32(&mbar)), "r"(bytes) : "memory");
      for (unsigned off = 0; off < bytes; off += 16384u)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ens_smem_u32(reinterpret_cast<unsigned char*>(sbs) + off)),
                     "l"(reinterpret_cast<const unsigned char*>(P.bs + k0) + off), "r"(min(16384u, bytes - off)),
                     "r"(ens_smem_u32(&mbar))
                     : "memory");
    }
  }
  const size_t g0 = ((size_t)lo << LB) + tid;          // this thread's first pair; pairs are HF_ENT apart
  // ---- phase 1 (own rows): loads first, control block afterwards
  double xv[HF_ERPT], rv[HF_ERPT], pv[HF_ERPT], qv[HF_ERPT], sv[HF_ERPT];
#pragma unroll
  for (int t = 0; t < HF_ERPT; ++t) {
    const size_t g = g0 + (size_t)t * HF_ENT;
    xv[t] = x[g];
    rv[t] = ro[g];
    pv[t] = po[g];
    qv[t] = qo[g];
    sv[t] = shat[g];
  }
  for (int k = k0 + tid; k < k1; k += HF_ENT) slc[k - k0] = P.lcol[k];
  for (int i = tid; i <= R; i += HF_ENT) srow[i] = rowptr[min(lo + i, N)] - k0;
  const double kb = ks[b];
  __syncthreads();                      // control block (and the mbarrier init) visible to the CTA
  const int done = s_ctl[0], it = s_ctl[1];
  const double alpha = s_al[b], beta = s_be[b];
  if (done) return;
  double l_rr = 0.0;
#pragma unroll
  for (int t = 0; t < HF_ERPT; ++t) {
    const size_t g = g0 + (size_t)t * HF_ENT;
    if (it > 0) {
      if (alpha != 0.0) x[g] = fma(alpha * sv[t], pv[t], xv[t]);
      rv[t] = fma(-alpha, qv[t], rv[t]);
      pv[t] = fma(beta, pv[t], rv[t]);
    } else {
      pv[t] = rv[t];
    }
    rn[g] = rv[t];
    pn[g] = pv[t];
    sp[t * HF_ENT + tid] = sv[t] * pv[t];
    l_rr = fma(rv[t], rv[t], l_rr);
  }
  // halo rows: same update, result only in shared memory
  for (int hidx = tid; hidx < (nh << LB); hidx += HF_ENT) {
    const size_t g = ((size_t)P.halo_idx[hp0 + (hidx >> LB)] << LB) + b;
    const double r_old = ro[g];
    double p_new = r_old;
    if (it > 0) p_new = fma(beta, po[g], fma(-alpha, qo[g], r_old));
    sp[(R << LB) + hidx] = shat[g] * p_new;
  }
  __syncthreads();                      // sp, slc, srow complete
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ENS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra ENS_DONE;\n"
      "bra ENS_WAIT;\n"
      "ENS_DONE:\n"
      "}\n" ::"r"(ens_smem_u32(&mbar))
      : "memory");
  // ---- phase 2: q = shat_i * sum_j (base0 + k_b S0)_ij * sp_j, all operands in shared memory
  double l_pq = 0.0, l_rq = 0.0, l_qq = 0.0;
#pragma unroll
  for (int t = 0; t < HF_ERPT; ++t) {
    const int il = (t * HF_ENT + tid) >> LB;
    const int i = lo + il;
    double acc = 0.0;
    if (i < N && !bcflag[i]) {
      const int ka = srow[il], kz = srow[il + 1];
      double a0 = 0.0, a1 = 0.0;
      int k = ka;
      for (; k + 2 <= kz; k += 2) {
        const double2 m0 = sbs[k], m1 = sbs[k + 1];
        a0 = fma(fma(kb, m0.y, m0.x), sp[((int)slc[k] << LB) + b], a0);
        a1 = fma(fma(kb, m1.y, m1.x), sp[((int)slc[k + 1] << LB) + b], a1);
      }
      if (k < kz) {
        const double2 m0 = sbs[k];
        a0 = fma(fma(kb, m0.y, m0.x), sp[((int)slc[k] << LB) + b], a0);
      }
      acc = (a0 + a1) * sv[t];
    }
    qn[g0 + (size_t)t * HF_ENT] = acc;
    l_pq = fma(pv[t], acc, l_pq);
    l_rq = fma(rv[t], acc, l_rq);
    l_qq = fma(acc, acc, l_qq);
  }
  // ---- per-CTA, per-variant partials; the last CTA finalises the iteration
  const size_t GB = (size_t)gridDim.x * B;
  ens_block_sum<LB, HF_ENT>(l_rr, sh, s_tot[0]);
  ens_block_sum<LB, HF_ENT>(l_pq, sh, s_tot[1]);
  ens_block_sum<LB, HF_ENT>(l_rq, sh, s_tot[2]);
  ens_block_sum<LB, HF_ENT>(l_qq, sh, s_tot[3]);
  if (tid < B) {
#pragma unroll
    for (int a = 1; a < 4; ++a) __stcg(part + a * GB + (size_t)blockIdx.x * B + tid, s_tot[a][tid]);
  }
  if (!ens_publish<LB>(s_tot[0], part, &c->counter[1])) return;
#pragma unroll
  for (int a = 0; a < 4; ++a) ens_sum_parts<LB, HF_ENT>(part + a * GB, gridDim.x, sh, s_tot[a]);
  if (tid < B) {
    const double rr = s_tot[0][tid];
    c->rz[tid] = rr;
    double al = 0.0, be = 0.0;
    int act = c->active[tid];
    if (act) {
      if (!(rr > c->thr[tid])) {
        act = 0;                          // frozen from now on: x, r stay, p = r
      } else {
        al = rr / s_tot[1][tid];
        const double rr_next = fma(al * al, s_tot[3][tid], fma(-2.0 * al, s_tot[2][tid], rr));
        be = fmax(rr_next, 0.0) / rr;
      }
    }
    c->alpha[tid] = al;
    c->beta[tid] = be;
    c->active[tid] = act;
  }
  __syncthreads();
  if (tid == 0) {
    int na = 0;
    for (int k = 0; k < B; ++k) na += c->active[k];
    c->n_active = na;
    if (na == 0) c->done = 1;
    else c->it = it + 1;
  }
}

// {base0, S0} interleaved per CSR slot (16 bytes per entry: any row range is TMA-aligned)
__global__ void k_ens_pack(long long nnz, const double* __restrict__ base0, const double* __restrict__ s0, double2* __restrict__ bs) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) bs[k] = make_double2(base0[k], s0[k]);
}

template <int LB>
__global__ void k_ens_sample(int n_watch, int n_steps, int step, const int* __restrict__ nodes,
                             const double* __restrict__ x, double* __restrict__ hist) {
  constexpr int B = 1 << LB;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = idx >> LB, b = idx & (B - 1);
  if (w < n_watch) hist[((size_t)b * n_steps + step) * n_watch + w] = x[((size_t)nodes[w] << LB) + b];
}

// [N,B] internal numbering -> [B,N] caller numbering (rank == nullptr: identity)
template <int LB>
__global__ void k_ens_transpose(int N, const int* __restrict__ rank, const double* __restrict__ src, double* __restrict__ dst) {
  constexpr int B = 1 << LB;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = (int)(idx >> LB), b = (int)(idx & (B - 1));
  if (i < N) dst[(size_t)b * N + i] = src[((size_t)(rank ? rank[i] : i) << LB) + b];
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
#define ENS_DISPATCH(LBV, ...)                                 \
  switch (LBV) {                                               \
    case 2: { constexpr int LB = 2; __VA_ARGS__; } break;      \
    case 3: { constexpr int LB = 3; __VA_ARGS__; } break;      \
    case 4: { constexpr int LB = 4; __VA_ARGS__; } break;      \
    default: { constexpr int LB = 5; __VA_ARGS__; } break;     \
  }

// kappa table with the sample material replaced by `k_sample_value`, others scaled by `others`
static int ens_assemble(hf_ctx* c, int sample_tag, double fm, double f_others, double f_sample, double* out) {
  const int nt = (int)c->mat_tags.size();
  std::vector<double> cm_t(nt), ck_t(nt);
  for (int k = 0; k < nt; ++k) {
    cm_t[k] = fm * c->mat_rhoc[k];
    ck_t[k] = (c->mat_tags[k] == sample_tag) ? f_sample : f_others * c->mat_kappa[k];
  }
  std::vector<int> tag(c->E);
  HF_TRY(c->cell_tag.download(tag.data(), c->E, c->stream));
  std::vector<double> cm(c->E), ck(c->E);
  for (int e = 0; e < c->E; ++e) {
    int f = -1;
    for (int k = 0; k < nt; ++k)
      if (c->mat_tags[k] == tag[e]) f = k;
    if (f < 0) return hf_fail(HF_ERR_ARG, "cell tag " + std::to_string(tag[e]) + " has no material");
    cm[e] = cm_t[f];
    ck[e] = ck_t[f];
  }
  HF_CUDA(cudaMemcpyAsync(c->cm.p, cm.data(), sizeof(double) * c->E, cudaMemcpyHostToDevice, c->stream));
  HF_CUDA(cudaMemcpyAsync(c->ck.p, ck.data(), sizeof(double) * c->E, cudaMemcpyHostToDevice, c->stream));
  HF_TRY(hf_assemble_values(c, c->cm.p, c->ck.p, c->axisym, out));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

extern "C" int hf_ens_create(hf_ctx* c, int32_t batch, const double* k_sample, const double* coeff, int32_t sample_tag) {
  if (!c || !k_sample || !coeff) return hf_fail(HF_ERR_ARG, "hf_ens_create: null argument");
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_ens_create: call hf_build_operator first");
  if (batch < 1 || batch > HF_EB) return hf_fail(HF_ERR_ARG, "hf_ens_create: batch must be in [1, 32] (tile larger sweeps on the host)");
  if (std::find(c->mat_tags.begin(), c->mat_tags.end(), sample_tag) == c->mat_tags.end())
    return hf_fail(HF_ERR_ARG, "hf_ens_create: sample_tag is not a material tag");
  for (int s = 0; s < batch; ++s)
    if (!(k_sample[s] > 0.0) || !std::isfinite(coeff[s])) return hf_fail(HF_ERR_ARG, "hf_ens_create: k_sample must be positive, coeff finite");
  cudaSetDevice(c->device);
  hf_ens_free(c);
  EnsState* e = new EnsState();
  c->ens = e;
  e->nb = batch;
  e->LB = 2;                            // tiles of fewer than 4 variants are padded: chunks stay <= 512 rows
  while ((1 << e->LB) < batch) ++e->LB;
  e->B = 1 << e->LB;
  const int B = e->B, N = c->N;
  e->R = HF_EPAIRS / B;
  e->nchunks = (N + e->R - 1) / e->R;
  const size_t nb = (size_t)e->nchunks * e->R * B;      // rows padded to whole chunks (padding stays zero)
  // pad the tile by repeating the last variant
  std::vector<double> ks(B), cf(B);
  for (int s = 0; s < B; ++s) {
    ks[s] = k_sample[std::min(s, batch - 1)];
    cf[s] = coeff[std::min(s, batch - 1)];
  }
  HF_TRY(e->ks.upload(ks.data(), B, c->stream));
  HF_TRY(e->coeff.upload(cf.data(), B, c->stream));
  HF_TRY(e->base0.alloc(c->nnz, c->stream));
  HF_TRY(e->s0.alloc(c->nnz, c->stream));
  HF_TRY(ens_assemble(c, sample_tag, 1.0, c->dt, 0.0, e->base0.p));
  HF_TRY(ens_assemble(c, sample_tag, 0.0, 0.0, c->dt, e->s0.p));
  for (DevBuf<double>* v : {&e->dinv, &e->g, &e->u, &e->x, &e->r, &e->r1, &e->p0, &e->p1, &e->q, &e->q1})
    HF_TRY(v->alloc(nb, c->stream));
  const size_t blocks = ((size_t)N * B + HF_ET - 1) / HF_ET;
  e->grid = (int)std::max<size_t>(1, std::min<size_t>(blocks, (size_t)c->sm_count * 8));
  HF_TRY(e->part.alloc((size_t)4 * std::max(e->grid, e->nchunks) * B, c->stream));
  {
    std::vector<int> hptr, hidx;
    std::vector<unsigned short> lcol;
    HF_TRY(hf_build_patches(c, e->R, hptr, hidx, lcol, &e->halo_max));
    if (hidx.empty()) hidx.push_back(0);
    HF_TRY(e->halo_ptr.upload(hptr.data(), hptr.size(), c->stream));
    HF_TRY(e->halo_idx.upload(hidx.data(), hidx.size(), c->stream));
    HF_TRY(e->lcol.upload(lcol.data(), lcol.size(), c->stream));
    for (int ch = 0; ch < e->nchunks; ++ch) {
      const int lo = ch * e->R, hi = std::min(lo + e->R, N);
      e->mcap = std::max(e->mcap, c->h_rowptr[hi] - c->h_rowptr[lo]);
    }
    e->mcap = (e->mcap + 7) & ~7;
    HF_TRY(e->bs.alloc(c->nnz, c->stream));
    k_ens_pack<<<(unsigned)((c->nnz + 255) / 256), 256, 0, c->stream>>>(c->nnz, e->base0.p, e->s0.p, e->bs.p);
    HF_CUDA(cudaGetLastError());
    // {base0,S0} block + shat*p (own + halo rows, B variants) + 16-bit local columns
    e->iter_smem = (size_t)e->mcap * 16 + sizeof(double) * (size_t)(e->R + e->halo_max) * B + (size_t)e->mcap * 2 + 16;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    if (e->iter_smem + 24 * 1024 > (size_t)max_smem)
      return hf_fail(HF_ERR_STATE, "ensemble tile does not fit in shared memory with this node ordering (halo of " +
                                       std::to_string(e->halo_max) + " rows); use hf_set_ordering(ctx, 2) or a smaller batch");
    HF_CUDA(cudaStreamSynchronize(c->stream));
  }
  HF_TRY(e->ctrl.alloc(1, c->stream));
  HF_CUDA(cudaMallocHost(&e->h_ctrl, sizeof(EnsCtrl)));
  DevBuf<int> bad;
  HF_TRY(bad.alloc(1, c->stream));
  ENS_DISPATCH(e->LB, k_ens_diag<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->rowptr.p, c->col.p, e->base0.p, e->s0.p,
                                                                                c->bcflag.p, e->ks.p, e->dinv.p, bad.p));
  ENS_DISPATCH(e->LB, k_ens_bcast<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->u.p, e->u.p));
  ENS_DISPATCH(e->LB, k_ens_bcast<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->gfull.p, e->g.p));
  HF_CUDA(cudaGetLastError());
  int hbad = 0;
  HF_TRY(bad.download(&hbad, 1, c->stream));
  if (hbad) return hf_fail(HF_ERR_STATE, "ensemble operator has a non-positive diagonal at row " + std::to_string(hbad - 1));
  return HF_OK;
}

static const int kEnsChunk[3] = {8, 32, 128};

static int ens_set_smem(EnsState* e) {
  const int sm = (int)e->iter_smem;
  if (sm > 48 * 1024) ENS_DISPATCH(e->LB, HF_CUDA(cudaFuncSetAttribute(k_ens_iter<LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)));
  return HF_OK;
}

static void ens_launch_iteration(hf_ctx* c, EnsState* e, int par) {
  const EnsPatch P{e->R, e->nchunks, e->mcap, e->halo_ptr.p, e->halo_idx.p, e->lcol.p, e->bs.p};
  ENS_DISPATCH(e->LB, k_ens_iter<LB><<<e->nchunks, HF_ENT, e->iter_smem, c->stream>>>(
                          c->N, P, par, c->rowptr.p, c->bcflag.p, e->ks.p, e->dinv.p, e->x.p, e->r.p, e->r1.p, e->p0.p, e->p1.p,
                          e->q.p, e->q1.p, e->part.p, e->ctrl.p));
}

static int ens_build_chunks(hf_ctx* c, EnsState* e) {
  HF_TRY(ens_set_smem(e));
  for (int k = 0; k < 3; ++k) {
    cudaGraph_t g;
    HF_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < kEnsChunk[k]; ++i) ens_launch_iteration(c, e, i & 1);   // even chunk lengths: parity = iteration parity
    HF_CUDA(cudaStreamEndCapture(c->stream, &g));
    HF_CUDA(cudaGraphInstantiate(&e->chunk_exec[k], g, 0));
    HF_CUDA(cudaGraphDestroy(g));
  }
  return HF_OK;
}

static int ens_solve(hf_ctx* c, EnsState* e, int* iters_out) {
  if (!e->chunk_exec[0]) HF_TRY(ens_build_chunks(c, e));
  HF_TRY(ens_set_smem(e));
  int launched = 0;
  int want = std::max(8, std::min(e->last_iters + 4, c->max_iters));
  for (;;) {
    while (want > 0) {
      int k = 2;
      while (k > 0 && kEnsChunk[k] > want) --k;
      HF_CUDA(cudaGraphLaunch(e->chunk_exec[k], c->stream));
      want -= kEnsChunk[k];
      launched += kEnsChunk[k];
      c->stat_launches += (unsigned long long)kEnsChunk[k];
    }
    HF_CUDA(cudaMemcpyAsync(e->h_ctrl, e->ctrl.p, sizeof(EnsCtrl), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    if (e->h_ctrl->done) break;
    for (int s = 0; s < e->B; ++s)
      if (!std::isfinite(e->h_ctrl->thr[s]) || !std::isfinite(e->h_ctrl->rz[s]))
        return hf_fail(HF_ERR_NOCONV, "ensemble PCG: non-finite residual in variant " + std::to_string(s));
    if (launched >= c->max_iters) {
      if (iters_out) *iters_out = launched;
      return hf_fail(HF_ERR_NOCONV, "ensemble PCG did not converge in " + std::to_string(launched) + " iterations (" +
                                        std::to_string(e->h_ctrl->n_active) + " variants still active)");
    }
    want = std::max(32, launched / 4);
  }
  for (int s = 0; s < e->B; ++s)
    if (!std::isfinite(e->h_ctrl->rz[s])) return hf_fail(HF_ERR_NOCONV, "ensemble PCG: non-finite residual in variant " + std::to_string(s));
  const int its = e->h_ctrl->it;
  e->last_iters = its;
  c->stat_iters += (unsigned long long)its;
  double worst = 0.0;
  for (int s = 0; s < e->B; ++s)
    if (e->h_ctrl->bn[s] > 0.0) worst = std::max(worst, std::sqrt(e->h_ctrl->rz[s] / e->h_ctrl->bn[s]));
  c->stat_relres = worst;
  if (iters_out) *iters_out = its;
  return HF_OK;
}

extern "C" int hf_ens_run(hf_ctx* c, int32_t n_steps, const double* amp, double t_ic, int32_t n_watch,
                          const int32_t* watch_nodes, double* hist, int32_t* iters) {
  if (!c || !c->ens) return hf_fail(HF_ERR_STATE, "hf_ens_run: call hf_ens_create first");
  if (n_steps < 0 || (n_steps && !amp) || n_watch < 0 || (n_watch && (!watch_nodes || !hist)))
    return hf_fail(HF_ERR_ARG, "hf_ens_run: bad arguments");
  std::vector<int> wn;
  if (hf_internal_nodes(c, n_watch, watch_nodes, wn) != HF_OK) return hf_fail(HF_ERR_ARG, "hf_ens_run: watch node out of range");
  cudaSetDevice(c->device);
  EnsState* e = c->ens;
  const int B = e->B, N = c->N;
  const size_t nb = (size_t)N * B;
  if (n_watch && n_steps) {
    HF_TRY(e->watch.upload(wn.data(), n_watch, c->stream));
    if (e->hist.n < (size_t)B * n_steps * n_watch) HF_TRY(e->hist.alloc((size_t)B * n_steps * n_watch, c->stream));
  }
  HF_CUDA(cudaEventRecord(c->ev0, c->stream));
  for (int s = 0; s < n_steps; ++s) {
    if (c->n_gauss) {
      const int n = c->n_gauss * B;
      ENS_DISPATCH(e->LB, k_ens_gauss<LB><<<(n + 255) / 256, 256, 0, c->stream>>>(c->n_gauss, c->gauss_dof.p, c->gauss_r.p, amp[s], t_ic,
                                                                                 e->coeff.p, e->g.p));
      c->stat_launches += 1;
    }
    ENS_DISPATCH(e->LB, k_ens_init<LB><<<e->grid, HF_ET, 0, c->stream>>>(N, c->rowptr.p, c->col.p, c->valM.p, e->base0.p, e->s0.p,
                                                                        c->bcflag.p, e->ks.p, e->g.p, e->u.p, e->dinv.p, e->x.p,
                                                                        e->r.p, e->part.p, e->ctrl.p, c->rtol));
    c->stat_launches += 1;
    HF_CUDA(cudaGetLastError());
    int it = 0;
    HF_TRY(ens_solve(c, e, &it));
    if (iters) iters[s] = it;
    if (n_watch) {
      const int n = n_watch * B;
      ENS_DISPATCH(e->LB, k_ens_sample<LB><<<(n + 255) / 256, 256, 0, c->stream>>>(n_watch, n_steps, s, e->watch.p, e->x.p, e->hist.p));
      c->stat_launches += 1;
    }
    HF_CUDA(cudaMemcpyAsync(e->u.p, e->x.p, sizeof(double) * nb, cudaMemcpyDeviceToDevice, c->stream));
  }
  HF_CUDA(cudaEventRecord(c->ev1, c->stream));
  if (n_watch && n_steps)   // device layout [B, S, W]; only the first nb variants are real
    HF_CUDA(cudaMemcpyAsync(hist, e->hist.p, sizeof(double) * (size_t)e->nb * n_steps * n_watch, cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  HF_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stat_run_ms = ms;
  return HF_OK;
}

extern "C" int hf_ens_get_state(hf_ctx* c, double* u) {
  if (!c || !c->ens || !u) return hf_fail(HF_ERR_STATE, "hf_ens_get_state: no ensemble / null argument");
  cudaSetDevice(c->device);
  EnsState* e = c->ens;
  const size_t nb = (size_t)c->N * e->B;
  if (e->stage.n < nb) HF_TRY(e->stage.alloc(nb, c->stream));
  ENS_DISPATCH(e->LB, k_ens_transpose<LB><<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(c->N, c->permuted ? c->rank_d.p : nullptr, e->u.p, e->stage.p));
  HF_CUDA(cudaGetLastError());
  return e->stage.download(u, (size_t)c->N * e->nb, c->stream);
}

extern "C" int hf_ens_destroy(hf_ctx* c) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hf_ens_free(c);
  return HF_OK;
}
