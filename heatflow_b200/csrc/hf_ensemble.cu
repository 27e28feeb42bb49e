// heatflow_b200 - batched multi-RHS ensemble for parameter_sweep (sm_100a).
//
// Replaces the reference's one-process-per-parameter-set loop (parameter_sweep.py:123-192,
// :436-438): B variants of one simulation - different sample conductivity k_b and Gaussian
// heating width fwhm_b, same mesh, same materials otherwise - advance together.
//
//   operator of variant b :  A0_b = base0 + k_b * S0      (two value arrays on the shared pattern)
//       base0 = M(rho_c) + dt K(kappa with kappa_sample = 0),  S0 = dt K(1 on sample cells, else 0)
//   Dirichlet rows are masked through the diagonal array (d = 0 there), Dirichlet columns carry a zero
//   search direction, so the same two arrays serve the RHS (un-BC'd operator, apply_lifting) and the solve.
//   vectors               :  [N, B] with the variant index fastest, so the gather of column j
//                            reads B contiguous doubles and the matrix entry is shared by B lanes.
//   thread mapping        :  one thread per (row, variant) pair.
//
// Jacobi-PCG per variant with its own alpha/beta/convergence (converged variants are frozen with
// alpha = beta = 0), written on the preconditioned residual z = D^-1 r, the search direction p and
// w = D^-1 A p, so that the halo update needs no per-variant scaling:
//     z_n = z - alpha w ;  p_n = z_n + beta p ;  q = A_b p_n ;  w_n = q / d
//     r.z = sum z^2 d ;  p.Ap = sum p q ;  r.(D^-1 A p) = sum z q ;  |D^-1/2 A p|^2 = sum q^2 / d
// One launch per iteration (k_ens_iter), same structure as the single-simulation streaming kernel
// (hf_pcg.cu): persistent CTAs walk chunks of HF_EPAIRS/B rows x B variants; the chunk's operator
// entries, local row pointers, 16-bit local columns and own-row vectors (x z w p d) stream into
// shared-memory stages by TMA bulk copies; halo values go through a register pipeline; the last CTA
// adds the per-CTA partial sums in CTA order (bit-reproducible) and sets alpha_n, beta_{n+1}, flags.
// Algorithmic traffic per dof, iteration and variant (nnz ~ 7 / row, halo fraction h):
//   matrix (18*nnz + 4)/B (two value arrays + 16-bit columns, shared by the B variants)
//   vectors: read x z w p d (40 + 24 h), write x z w p (32)             total ~ 72 + 24 h + 130/B bytes.
#include <algorithm>
#include <cmath>

#include "hf_ens.cuh"

void hf_ens_free(hf_ctx* c) {
  delete c->ens;
  c->ens = nullptr;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
// Sum over all threads of the CTA that share the same variant (tid & (B-1)); the result for
// variant b lands in sh_out[b] (valid after the trailing __syncthreads()).
template <int LB, int NT>
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

__device__ __forceinline__ unsigned ens_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ens_bulk(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar,
                                         unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          ens_smem_u32(dst_smem)),
      "l"(src), "r"(bytes), "r"(ens_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void ens_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ENS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ENS_DONE;\n"
      "bra ENS_WAIT;\n"
      "ENS_DONE:\n"
      "}\n" ::"r"(ens_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------
// per-variant diagonal d = base0_ii + k_b S0_ii; 0 marks a Dirichlet row
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_diag(int N, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ base0,
           const double* __restrict__ s0, const unsigned char* __restrict__ bcflag, const double* __restrict__ ks,
           double* __restrict__ dg, double* __restrict__ drow /*[2][rows]: base0_ii, S0_ii (0 on Dirichlet rows)*/, size_t rows,
           int* __restrict__ bad) {
  constexpr int B = 1 << LB;
  const size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x;
  const int i = (int)(idx >> LB), b = (int)(idx & (B - 1));
  if (i >= N) return;
  double d = 0.0, db = 0.0, ds = 0.0;
  if (!bcflag[i]) {
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) {
        db = base0[k];
        ds = s0[k];
      }
    d = fma(ks[b], ds, db);                    // the iteration kernel recomputes exactly this from drow
    if (!(d > 0.0)) {
      atomicExch(bad, i + 1);
      d = 1.0;
      db = 1.0;
      ds = 0.0;
    }
  }
  if (b == 0) {
    drow[i] = db;
    drow[rows + i] = ds;
  }
  dg[idx] = d;
}

template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_bcast(int N, const double* __restrict__ src, double* __restrict__ dst) {
  const size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x;
  const int i = (int)(idx >> LB);
  if (i < N) dst[idx] = src[i];
}

// {base0, S0} interleaved per CSR slot
__global__ void k_ens_pack(long long nnz, const double* __restrict__ base0, const double* __restrict__ s0, double2* __restrict__ bs) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) bs[k] = make_double2(base0[k], s0[k]);
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
//   free row i : z = ((M u)_i - (A0_b x0)_i) / d_i with x0 = u on free columns, g on Dirichlet columns
//   bc row i   : x = g, z = 0
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_init(int N, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ valM,
           const double* __restrict__ base0, const double* __restrict__ s0, const unsigned char* __restrict__ bcflag,
           const double* __restrict__ ks, const double* __restrict__ g, const double* __restrict__ u,
           const double* __restrict__ uprev, double warm, const double* __restrict__ dg, double* __restrict__ x,
           double* __restrict__ z, double* __restrict__ part, EnsCtrl* __restrict__ c, double rtol) {
  constexpr int B = 1 << LB;
  __shared__ double sh[HF_ET];
  __shared__ double s_bn[HF_EB];
  __shared__ int s_last;
  const int b = threadIdx.x & (B - 1);
  const double kb = ks[b];
  const size_t total = (size_t)N << LB;
  double l_bn = 0.0;
  for (size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x; idx < total; idx += (size_t)gridDim.x * HF_ET) {
    const int i = (int)(idx >> LB);
    double xv, zv = 0.0;
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
          t2 = fma(a, (warm != 0.0) ? fma(warm, uj - uprev[((size_t)j << LB) + b], uj) : uj, t2);
        }
      }
      const double di = 1.0 / dg[idx];
      xv = (warm != 0.0) ? fma(warm, u[idx] - uprev[idx], u[idx]) : u[idx];
      zv = (t1 - t2) * di;
      const double bi = t1 - t3;
      l_bn = fma(bi * di, bi, l_bn);
    }
    x[idx] = xv;
    z[idx] = zv;
  }
  ens_block_sum<LB, HF_ET>(l_bn, sh, s_bn);
  if (threadIdx.x < B) {
    __stcg(part + (size_t)blockIdx.x * B + threadIdx.x, s_bn[threadIdx.x]);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&c->counter[0], 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) c->counter[0] = 0u;
    __threadfence();
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < B) {
    double bn = 0.0;
    for (unsigned cta = 0; cta < gridDim.x; ++cta) bn += __ldcg(part + (size_t)cta * B + threadIdx.x);
    c->thr[threadIdx.x] = rtol * rtol * bn;
    c->bn[threadIdx.x] = bn;
    c->rz[threadIdx.x] = 0.0;
    c->alpha[threadIdx.x] = 0.0;
    c->beta[threadIdx.x] = 0.0;
    c->active[threadIdx.x] = 1;
  }
  if (threadIdx.x == 0) {
    c->n_active = B;
    c->done = 0;
    c->it = 0;
    c->counter[1] = 0u;
  }
}

// ---------------------------------------------------------------------------------------
// iteration kernel
// ---------------------------------------------------------------------------------------
struct EnsPatch {
  int nchunks, mcap, halo_cap, nstages;
  unsigned stage_bytes;
  const int* __restrict__ halo_ptr;
  const int* __restrict__ halo_idx;
  const int* __restrict__ rowptr_pad;
  const int* __restrict__ lc_off;
  const unsigned short* __restrict__ lcol;
  const double2* __restrict__ bs;
};

// Stage layout (bytes, every block a multiple of 16):
//   bs[mcap] double2 | x[EP] | z[EP] | w[EP] | (base0_ii, S0_ii)[2][R] | p[EP] | halo p[halo_cap*B] | lcol[mcap] u16 | rowptr[R+4] i32
template <int LB>
__device__ __forceinline__ void ens_issue_chunk(const EnsPatch& P, int ch, int k0, int k1, int lc0, unsigned char* st,
                                                const double* x, const double* zo, const double* wo, const double* drow,
                                                size_t rows, const double* po, unsigned long long* bar) {
  constexpr int B = 1 << LB;
  constexpr int R = HF_EPAIRS / B;
  constexpr unsigned vb = HF_EPAIRS * 8u;
  unsigned char* sx = st + (size_t)P.mcap * 16;
  unsigned char* slc = sx + 4 * (size_t)vb + 2 * (size_t)R * 8 + (size_t)P.halo_cap * B * 8;
  unsigned char* srow = slc + (size_t)P.mcap * 2;
  const unsigned n = (unsigned)(k1 - k0);
  const unsigned ncol = (n + 7u) & ~7u;
  unsigned long long keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was last touched by ordinary loads/stores
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ens_smem_u32(bar)),
               "r"(n * 16u + ncol * 2u + 4u * vb + 2u * R * 8u + (R + 4) * 4u)
               : "memory");
  const size_t lo = (size_t)ch * HF_EPAIRS;
  ens_bulk(sx, x + lo, vb, bar, keep);
  ens_bulk(sx + vb, zo + lo, vb, bar, keep);
  ens_bulk(sx + 2 * vb, wo + lo, vb, bar, keep);
  ens_bulk(sx + 3 * vb, drow + (size_t)ch * R, R * 8u, bar, keep);              // base0_ii of the chunk's rows
  ens_bulk(sx + 3 * vb + R * 8u, drow + rows + (size_t)ch * R, R * 8u, bar, keep);   // S0_ii
  ens_bulk(sx + 3 * vb + 2 * R * 8u, po + lo, vb, bar, keep);
  ens_bulk(srow, P.rowptr_pad + (size_t)ch * R, (R + 4) * 4u, bar, keep);
  if (ncol) ens_bulk(slc, P.lcol + lc0, ncol * 2u, bar, keep);
  for (unsigned off = 0; off < n * 16u; off += 16384u)
    ens_bulk(st + off, reinterpret_cast<const unsigned char*>(P.bs + k0) + off, min(16384u, n * 16u - off), bar, keep);
}

template <int LB>
__global__ void __launch_bounds__(HF_ENT, HF_EMINB)
k_ens_iter(EnsPatch P, int par, const double* __restrict__ ks, const double* __restrict__ drow, size_t rows, double* __restrict__ x,
           double* __restrict__ zb0, double* __restrict__ zb1, double* __restrict__ pb0, double* __restrict__ pb1,
           double* __restrict__ wb0, double* __restrict__ wb1, double* __restrict__ part, EnsCtrl* __restrict__ c) {
  constexpr int B = 1 << LB;
  constexpr int R = HF_EPAIRS / B;
  constexpr int NW = HF_ENT / 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ double sh[NW * 32];
  __shared__ double s_tot[4][HF_EB];
  __shared__ double s_al[HF_EB], s_be[HF_EB];
  __shared__ int s_ctl[2];
  __shared__ __align__(8) unsigned long long full[4];
  const int tid = threadIdx.x;
  const int b = tid & (B - 1);
  const int G = gridDim.x;
  const double* __restrict__ zo = par ? zb1 : zb0;
  const double* __restrict__ po = par ? pb1 : pb0;
  const double* __restrict__ wo = par ? wb1 : wb0;
  double* __restrict__ zn = par ? zb0 : zb1;
  double* __restrict__ pn = par ? pb0 : pb1;
  double* __restrict__ wn = par ? wb0 : wb1;
  const int nloc = (P.nchunks - (int)blockIdx.x + G - 1) / G;
  // ---- control block: ONE reader per CTA and variant (the whole grid polls these L2 lines)
  if (tid < B) {
    s_al[tid] = *(volatile double*)&c->alpha[tid];
    s_be[tid] = *(volatile double*)&c->beta[tid];
  }
  if (tid == 0) {
    const int d = *(volatile int*)&c->done;
    s_ctl[0] = d;
    s_ctl[1] = *(volatile int*)&c->it;
    if (d == 0) {
      for (int st = 0; st < P.nstages; ++st) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ens_smem_u32(&full[st])) : "memory");
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      int e0[4], e1[4], el[4];          // all extents first: one round trip instead of one per stage
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = min((int)blockIdx.x + j * G, P.nchunks - 1);
        e0[j] = P.rowptr_pad[ch * R];
        e1[j] = P.rowptr_pad[(ch + 1) * R];
        el[j] = P.lc_off[ch];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < min(P.nstages, nloc))
          ens_issue_chunk<LB>(P, blockIdx.x + j * G, e0[j], e1[j], el[j], smraw + (size_t)j * P.stage_bytes, x, zo, wo, drow, rows, po,
                              &full[j]);
    }
  }
  // ---- halo pipeline prologue: extents of chunks 0..2, node indices of chunks 0..1, values of chunk 0
  // (list begin, list end) per chunk; the subtraction happens where the count is used, one iteration after
  // the loads were issued, so that no load latency is exposed inside the chunk loop
  int hp_a = 0, he_a = 0, hp_b = 0, he_b = 0, hp_c = 0, he_c = 0;
  if (nloc > 0) {
    hp_a = P.halo_ptr[blockIdx.x];
    he_a = P.halo_ptr[blockIdx.x + 1];
  }
  if (nloc > 1) {
    hp_b = P.halo_ptr[blockIdx.x + G];
    he_b = P.halo_ptr[blockIdx.x + G + 1];
  }
  if (nloc > 2) {
    hp_c = P.halo_ptr[blockIdx.x + 2 * G];
    he_c = P.halo_ptr[blockIdx.x + 2 * G + 1];
  }
  const int nh_b0 = he_b - hp_b;
  int nh_a = he_a - hp_a;
  int g_b[HF_EHPT];
  double hz[HF_EHPT], hpv[HF_EHPT], hw[HF_EHPT];
#pragma unroll
  for (int t = 0; t < HF_EHPT; ++t) {
    const int h = (t * HF_ENT + tid) >> LB;
    g_b[t] = (h < nh_b0) ? P.halo_idx[hp_b + h] : -1;
    hz[t] = hpv[t] = hw[t] = 0.0;
    if (h < nh_a) {
      const size_t g = ((size_t)P.halo_idx[hp_a + h] << LB) + b;
      hz[t] = zo[g];
      hpv[t] = po[g];
      hw[t] = wo[g];
    }
  }
  const double kb = ks[b];
  __syncthreads();                      // control block and mbarrier inits visible to the CTA
  const int done = s_ctl[0], it = s_ctl[1];
  const double alpha = s_al[b], beta = s_be[b];
  if (done) return;
  double l_rr = 0.0, l_pq = 0.0, l_rq = 0.0, l_qq = 0.0;
  for (int j = 0; j < nloc; ++j) {
    const int ch = blockIdx.x + j * G;
    const int stg = j % P.nstages;
    const unsigned parity = (unsigned)(j / P.nstages) & 1u;
    unsigned char* st = smraw + (size_t)stg * P.stage_bytes;
    const double2* sbs = reinterpret_cast<const double2*>(st);
    double* sx = reinterpret_cast<double*>(st + (size_t)P.mcap * 16);
    double* sz = sx + HF_EPAIRS;
    double* sw = sz + HF_EPAIRS;
    double* sd = sw + HF_EPAIRS;
    double* sp = sd + 2 * R;            // own pairs, halo pairs follow at sp[HF_EPAIRS + ...]
    const unsigned short* slc = reinterpret_cast<const unsigned short*>(sp + HF_EPAIRS + (size_t)P.halo_cap * B);
    const int* srow = reinterpret_cast<const int*>(slc + P.mcap);
    const size_t g0 = (size_t)ch * HF_EPAIRS;
    // thread 0: extents of the chunk that will refill this stage (needed only after phase 2)
    int nx_k0 = 0, nx_k1 = 0, nx_lc = 0;
    if (tid == 0 && j + P.nstages < nloc) {
      const int chn = ch + P.nstages * G;
      nx_k0 = P.rowptr_pad[chn * R];
      nx_k1 = P.rowptr_pad[(chn + 1) * R];
      nx_lc = P.lc_off[chn];
    }
    ens_mbar_wait(&full[stg], parity);
    // ---- phase 1: finish iteration n-1 on the own pairs (from the stage) and the halo pairs (registers)
    double dinv[HF_ERPT];
#pragma unroll
    for (int t = 0; t < HF_ERPT; ++t) {
      const int i = t * HF_ENT + tid;
      const double d = fma(kb, sd[R + (i >> LB)], sd[i >> LB]);   // diagonal of A_b from the row's (base0_ii, S0_ii): 16 B per row instead of 8 B per pair
      dinv[t] = (d > 0.0) ? __drcp_rn(d) : 0.0;
      double z_new = sz[i], p_new = z_new;
      if (it > 0) {
        const double pv = sp[i];
        if (alpha != 0.0) x[g0 + i] = fma(alpha, pv, sx[i]);
        z_new = fma(-alpha, sw[i], z_new);
        p_new = fma(beta, pv, z_new);
      }
      zn[g0 + i] = z_new;
      pn[g0 + i] = p_new;
      sp[i] = p_new;
      sz[i] = z_new;
      l_rr = fma(z_new * d, z_new, l_rr);
    }
#pragma unroll
    for (int t = 0; t < HF_EHPT; ++t) {
      const int hidx = t * HF_ENT + tid;
      if ((hidx >> LB) < nh_a) sp[HF_EPAIRS + hidx] = (it > 0) ? fma(beta, hpv[t], fma(-alpha, hw[t], hz[t])) : hz[t];
    }
    for (int hidx = HF_EHPT * HF_ENT + tid; (hidx >> LB) < nh_a; hidx += HF_ENT) {   // oversized halos: synchronous
      const size_t g = ((size_t)P.halo_idx[hp_a + (hidx >> LB)] << LB) + b;
      const double z_old = zo[g];
      sp[HF_EPAIRS + hidx] = (it > 0) ? fma(beta, po[g], fma(-alpha, wo[g], z_old)) : z_old;
    }
    __syncthreads();
    // ---- advance the halo pipeline (loads are consumed one iteration later)
#pragma unroll
    for (int t = 0; t < HF_EHPT; ++t) {
      if (g_b[t] >= 0) {
        const size_t g = ((size_t)g_b[t] << LB) + b;
        hz[t] = zo[g];
        hpv[t] = po[g];
        hw[t] = wo[g];
      }
      const int h = (t * HF_ENT + tid) >> LB;
      g_b[t] = (h < he_c - hp_c) ? P.halo_idx[hp_c + h] : -1;
    }
    hp_a = hp_b;
    nh_a = he_b - hp_b;
    hp_b = hp_c;
    he_b = he_c;
    if (j + 3 < nloc) {
      hp_c = P.halo_ptr[ch + 3 * G];
      he_c = P.halo_ptr[ch + 3 * G + 1];
    } else {
      hp_c = he_c = 0;
    }
    // ---- phase 2: q = sum_j (base0 + k_b S0)_ij p_j, everything from shared memory; w = q / d
    const int k00 = srow[0];
#pragma unroll
    for (int t = 0; t < HF_ERPT; ++t) {
      const int i = t * HF_ENT + tid;
      const int il = i >> LB;
      double q = 0.0;
      if (dinv[t] != 0.0) {
        const int ka = srow[il] - k00, kz = srow[il + 1] - k00;
        double a0 = 0.0, a1 = 0.0;
        int k = ka;
        for (; k + 2 <= kz; k += 2) {
          const double2 m0 = sbs[k], m1 = sbs[k + 1];
          const int c0 = slc[k], c1 = slc[k + 1];
          a0 = fma(fma(kb, m0.y, m0.x), sp[(c0 << LB) + b], a0);
          a1 = fma(fma(kb, m1.y, m1.x), sp[(c1 << LB) + b], a1);
        }
        if (k < kz) {
          const double2 m0 = sbs[k];
          a0 = fma(fma(kb, m0.y, m0.x), sp[((int)slc[k] << LB) + b], a0);
        }
        q = a0 + a1;
      }
      wn[g0 + i] = q * dinv[t];
      l_pq = fma(sp[i], q, l_pq);
      l_rq = fma(sz[i], q, l_rq);
      l_qq = fma(q * dinv[t], q, l_qq);
    }
    __syncthreads();                    // the stage is free again
    if (tid == 0 && j + P.nstages < nloc)
      ens_issue_chunk<LB>(P, ch + P.nstages * G, nx_k0, nx_k1, nx_lc, st, x, zo, wo, drow, rows, po, &full[stg]);
  }
  // ---- per-CTA, per-variant partials; the last CTA finalises the iteration
  const size_t GB = (size_t)G * B;
  ens_block_sum<LB, HF_ENT>(l_rr, sh, s_tot[0]);
  ens_block_sum<LB, HF_ENT>(l_pq, sh, s_tot[1]);
  ens_block_sum<LB, HF_ENT>(l_rq, sh, s_tot[2]);
  ens_block_sum<LB, HF_ENT>(l_qq, sh, s_tot[3]);
  if (tid >= 32) return;
  if (tid < B) {
#pragma unroll
    for (int a = 0; a < 4; ++a) __stcg(part + a * GB + (size_t)blockIdx.x * B + tid, s_tot[a][tid]);
    __threadfence();
  }
  __syncwarp();
  int last = 0;
  if (tid == 0) {
    const unsigned ticket = atomicAdd(&c->counter[1], 1u);
    last = (ticket == (unsigned)G - 1u);
    if (last) {
      c->counter[1] = 0u;
      __threadfence();
    }
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  int act = 0;
  if (tid < B) {
    double t4[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double v = 0.0;
      for (int cta = 0; cta < G; ++cta) v += __ldcg(part + a * GB + (size_t)cta * B + tid);   // fixed order
      t4[a] = v;
    }
    const double rr = t4[0];
    c->rz[tid] = rr;
    double al = 0.0, be = 0.0;
    act = c->active[tid];
    if (act) {
      if (!(rr > c->thr[tid])) {
        act = 0;                          // frozen from now on: x, z stay, p = z
      } else {
        al = rr / t4[1];
        const double rr_next = fma(al * al, t4[3], fma(-2.0 * al, t4[2], rr));
        be = fmax(rr_next, 0.0) / rr;
      }
    }
    c->alpha[tid] = al;
    c->beta[tid] = be;
    c->active[tid] = act;
  }
  const unsigned any = __ballot_sync(0xffffffffu, act != 0);
  if (tid == 0) {
    c->n_active = __popc(any);
    if (any == 0u) c->done = 1;
    else c->it = it + 1;
  }
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
// recycled initial guess (see hf_recycle.cu), one basis per variant.  All arrays use the [row, variant]
// layout, so element idx belongs to variant idx & (B-1); inner products are taken in the A_b inner
// product through the weight dg (the solver works with z = D^-1 r and w = D^-1 A p).
// ---------------------------------------------------------------------------------------
#define ERC_SEG 1024
#define ERC_WARPS 8

// parts[(k * nseg + seg) * B + b] = sum over the elements of variant b in segment seg of V[k] * v * wgt
template <int LB>
__global__ void __launch_bounds__(ERC_WARPS * 32)
k_erc_dots(int m, int nseg, size_t ld, const double* __restrict__ V, const double* __restrict__ v,
           const double* __restrict__ wgt, double* __restrict__ parts) {
  constexpr int B = 1 << LB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x;
  const size_t base = (size_t)seg * ERC_SEG + 2 * lane;
  double2 vv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const double2 a = *reinterpret_cast<const double2*>(v + base + 64 * j);
    const double2 g = *reinterpret_cast<const double2*>(wgt + base + 64 * j);
    vv[j] = make_double2(a.x * g.x, a.y * g.y);
  }
  for (int k = warp; k < m; k += ERC_WARPS) {
    const double* p = V + (size_t)k * ld + base;
    double2 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = __ldcs(reinterpret_cast<const double2*>(p + 64 * j));
    double s0 = 0.0, s1 = 0.0;          // variants (2 lane) & (B-1) and the next one: 64 j and 1024 seg are multiples of B
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s0 = fma(a[j].x, vv[j].x, s0);
      s1 = fma(a[j].y, vv[j].y, s1);
    }
#pragma unroll
    for (int o = 16; o >= B / 2; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane < B / 2) {
      double* out = parts + ((size_t)k * nseg + seg) * B + 2 * lane;
      out[0] = s0;
      out[1] = s1;
    }
  }
}

// coef[k, b] = sign * inv[k, b] * sum_seg parts[k, seg, b]   (one warp per k, fixed order)
template <int LB>
__global__ void __launch_bounds__(ERC_WARPS * 32)
k_erc_coef(int m, int nseg, const double* __restrict__ parts, const double* __restrict__ inv, double sign,
           double* __restrict__ coef) {
  constexpr int B = 1 << LB;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * ERC_WARPS + (threadIdx.x >> 5);
  if (k >= m) return;
  const int b = lane & (B - 1);
  double s = 0.0;
  for (int sg = lane >> LB; sg < nseg; sg += 32 >> LB) s += __ldcg(parts + ((size_t)k * nseg + sg) * B + b);
#pragma unroll
  for (int o = 16; o >= B; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane < B) coef[(size_t)k * B + lane] = sign * inv[(size_t)k * B + lane] * s;
}

// GS = false: a = x0, b = z0:  a += W c ; b -= AW c ; outW = a (null once the basis is frozen)
// GS = true : a = d,  b = Ad, coef = -h:  outW = a + W coef ; outAW = b + AW coef ; per-variant partial
//             sums of outW . outAW . wgt -> part[blockIdx.x * B + b]
template <bool GS, int LB>
__global__ void __launch_bounds__(HF_ET)
k_erc_update(int m, size_t n, size_t ld, const double* __restrict__ W, const double* __restrict__ AW,
             const double* __restrict__ coef, double* __restrict__ a, double* __restrict__ bv, double* __restrict__ outW,
             double* __restrict__ outAW, const double* __restrict__ wgt, double* __restrict__ part) {
  constexpr int B = 1 << LB;
  extern __shared__ double s_coef[];
  __shared__ double sh[HF_ET];
  __shared__ double s_out[HF_EB];
  for (int k = threadIdx.x; k < m * B; k += HF_ET) s_coef[k] = coef[k];
  __syncthreads();
  const int b = threadIdx.x & (B - 1);
  double local = 0.0;
  for (size_t i = (size_t)blockIdx.x * HF_ET + threadIdx.x; i < n; i += (size_t)gridDim.x * HF_ET) {
    double ca0 = 0.0, ca1 = 0.0, cb0 = 0.0, cb1 = 0.0;
    const double* w = W + i;
    const double* aw = AW + i;
    int k = 0;
    for (; k + 4 <= m; k += 4) {
      const double w0 = hf_ld_stream(w + (size_t)k * ld), w1 = hf_ld_stream(w + (size_t)(k + 1) * ld),
                   w2 = hf_ld_stream(w + (size_t)(k + 2) * ld), w3 = hf_ld_stream(w + (size_t)(k + 3) * ld);
      const double z0 = hf_ld_stream(aw + (size_t)k * ld), z1 = hf_ld_stream(aw + (size_t)(k + 1) * ld),
                   z2 = hf_ld_stream(aw + (size_t)(k + 2) * ld), z3 = hf_ld_stream(aw + (size_t)(k + 3) * ld);
      const double c0 = s_coef[k * B + b], c1 = s_coef[(k + 1) * B + b], c2 = s_coef[(k + 2) * B + b], c3 = s_coef[(k + 3) * B + b];
      ca0 = fma(c0, w0, ca0);
      ca1 = fma(c1, w1, ca1);
      ca0 = fma(c2, w2, ca0);
      ca1 = fma(c3, w3, ca1);
      cb0 = fma(c0, z0, cb0);
      cb1 = fma(c1, z1, cb1);
      cb0 = fma(c2, z2, cb0);
      cb1 = fma(c3, z3, cb1);
    }
    for (; k < m; ++k) {
      const double ck = s_coef[k * B + b];
      ca0 = fma(ck, hf_ld_stream(w + (size_t)k * ld), ca0);
      cb0 = fma(ck, hf_ld_stream(aw + (size_t)k * ld), cb0);
    }
    const double ca = ca0 + ca1, cb = cb0 + cb1;
    if (GS) {
      const double d = a[i] + ca, ad = bv[i] + cb;
      outW[i] = d;
      outAW[i] = ad;
      local = fma(d * wgt[i], ad, local);
    } else {
      const double x = a[i] + ca;
      a[i] = x;
      bv[i] -= cb;
      if (outW) outW[i] = x;
    }
  }
  if (GS) {
    ens_block_sum<LB, HF_ET>(local, sh, s_out);
    if (threadIdx.x < B) part[(size_t)blockIdx.x * B + threadIdx.x] = s_out[threadIdx.x];
  }
}

// d = x - x0 and ad = D^-1 A_b d on free rows (0 on Dirichlet rows), one thread per (row, variant)
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_erc_spmv(int N, const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ base0,
           const double* __restrict__ s0, const unsigned char* __restrict__ bcflag, const double* __restrict__ ks,
           const double* __restrict__ dg, const double* __restrict__ x, const double* __restrict__ x0,
           double* __restrict__ d, double* __restrict__ ad) {
  constexpr int B = 1 << LB;
  const size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x;
  const int i = (int)(idx >> LB), b = (int)(idx & (B - 1));
  if (i >= N) return;
  double dv = 0.0, av = 0.0;
  if (!bcflag[i]) {
    const double kb = ks[b];
    double a0 = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const size_t j = ((size_t)col[k] << LB) + b;
      a0 = fma(fma(kb, s0[k], base0[k]), x[j] - x0[j], a0);
    }
    dv = x[idx] - x0[idx];
    av = a0 / dg[idx];
  }
  d[idx] = dv;
  ad[idx] = av;
}

template <int LB>
__global__ void k_erc_norm(int nparts, const double* __restrict__ part, double* __restrict__ inv, int slot) {
  constexpr int B = 1 << LB;
  const int b = threadIdx.x;
  if (b >= B) return;
  double nn = 0.0;
  for (int i = 0; i < nparts; ++i) nn += __ldcg(part + (size_t)i * B + b);
  inv[(size_t)slot * B + b] = (nn > 0.0 && isfinite(nn) && isfinite(1.0 / nn)) ? 1.0 / nn : 0.0;
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

// assemble fm * M(rho_c) + K(f_others * kappa on non-sample cells, f_sample on sample cells)
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
  e->LB = 2;                            // tiles of fewer than 4 variants are padded
  while ((1 << e->LB) < batch) ++e->LB;
  e->B = 1 << e->LB;
  const int B = e->B, N = c->N;
  e->R = HF_EPAIRS / B;
  e->nchunks = (N + e->R - 1) / e->R;
  const size_t nb = (size_t)e->nchunks * HF_EPAIRS;      // rows padded to whole chunks (padding stays zero)
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
  for (DevBuf<double>* v : {&e->dg, &e->g, &e->u, &e->uprev, &e->x, &e->z, &e->z1, &e->p0, &e->p1, &e->w, &e->w1})
    HF_TRY(v->alloc(nb, c->stream));
  HF_TRY(e->drow.alloc((size_t)2 * e->nchunks * e->R, c->stream));
  const size_t blocks = ((size_t)N * B + HF_ET - 1) / HF_ET;
  e->grid = (int)std::max<size_t>(1, std::min<size_t>(blocks, (size_t)c->sm_count * 8));
  e->igrid = std::min(e->nchunks, c->sm_count * HF_EMINB);
  HF_TRY(e->part.alloc((size_t)4 * std::max(e->grid, e->igrid) * B, c->stream));
  // ---- patch decomposition: halo lists, chunk-padded local columns, padded row pointers
  {
    std::vector<int> hptr, hidx;
    std::vector<unsigned short> lcol;
    HF_TRY(hf_build_patches(c, e->R, hptr, hidx, lcol, &e->halo_max));
    if (hidx.empty()) hidx.push_back(0);
    std::vector<int> lc_off(e->nchunks + 1, 0), rp((size_t)e->nchunks * e->R + 5, (int)c->nnz);
    std::copy(c->h_rowptr.begin(), c->h_rowptr.end(), rp.begin());
    for (int ch = 0; ch < e->nchunks; ++ch) {
      const int lo = ch * e->R, hi = std::min(lo + e->R, N);
      const int n = c->h_rowptr[hi] - c->h_rowptr[lo];
      e->mcap = std::max(e->mcap, n);
      lc_off[ch + 1] = lc_off[ch] + ((n + 7) & ~7);
    }
    e->mcap = (e->mcap + 7) & ~7;
    std::vector<unsigned short> lpad((size_t)lc_off[e->nchunks] + 8, 0);
    for (int ch = 0; ch < e->nchunks; ++ch) {
      const int lo = ch * e->R, hi = std::min(lo + e->R, N);
      std::copy(lcol.begin() + c->h_rowptr[lo], lcol.begin() + c->h_rowptr[hi], lpad.begin() + lc_off[ch]);
    }
    HF_TRY(e->halo_ptr.upload(hptr.data(), hptr.size(), c->stream));
    HF_TRY(e->halo_idx.upload(hidx.data(), hidx.size(), c->stream));
    HF_TRY(e->lcol.upload(lpad.data(), lpad.size(), c->stream));
    HF_TRY(e->lc_off.upload(lc_off.data(), lc_off.size(), c->stream));
    HF_TRY(e->rowptr_pad.upload(rp.data(), rp.size(), c->stream));
    HF_TRY(e->bs.alloc(c->nnz + 1, c->stream));
    k_ens_pack<<<(unsigned)((c->nnz + 255) / 256), 256, 0, c->stream>>>(c->nnz, e->base0.p, e->s0.p, e->bs.p);
    HF_CUDA(cudaGetLastError());
    e->halo_cap = (e->halo_max + 1) & ~1;
    e->stage_bytes = ((size_t)e->mcap * 18 + sizeof(double) * ((size_t)4 * HF_EPAIRS + 2 * (size_t)e->R + (size_t)e->halo_cap * B) + 4 * (e->R + 4) + 127) &
                     ~(size_t)127;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    int per_sm = 0;
    cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device);
    // HF_EMINB co-resident CTAs share the SM (1 KB per CTA is reserved by the system); 8 KB static shared memory
    const size_t avail = std::min<size_t>((size_t)max_smem, (size_t)per_sm / HF_EMINB - 1024) - 8 * 1024;
    e->nstages = (int)std::min<size_t>(4, avail / e->stage_bytes);
    if (const char* env = getenv("HF_STAGES")) e->nstages = std::max(1, std::min(e->nstages, atoi(env)));
    if (e->nstages < 1)
      return hf_fail(HF_ERR_STATE, "ensemble tile does not fit in shared memory with this node ordering (halo of " +
                                       std::to_string(e->halo_max) + " rows); use hf_set_ordering(ctx, 2) or a smaller batch");
    e->iter_smem = e->stage_bytes * e->nstages;
    HF_CUDA(cudaStreamSynchronize(c->stream));           // host staging vectors go out of scope
  }
  HF_TRY(e->ctrl.alloc(1, c->stream));
  HF_CUDA(cudaMallocHost(&e->h_ctrl, sizeof(EnsCtrl)));
  DevBuf<int> bad;
  HF_TRY(bad.alloc(1, c->stream));
  ENS_DISPATCH(e->LB, k_ens_diag<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->rowptr.p, c->col.p, e->base0.p, e->s0.p,
                                                                                c->bcflag.p, e->ks.p, e->dg.p, e->drow.p,
                                                                                (size_t)e->nchunks * e->R, bad.p));
  ENS_DISPATCH(e->LB, k_ens_bcast<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->u.p, e->u.p));
  ENS_DISPATCH(e->LB, k_ens_bcast<LB><<<(unsigned)blocks, HF_ET, 0, c->stream>>>(N, c->gfull.p, e->g.p));
  HF_CUDA(cudaGetLastError());
  int hbad = 0;
  HF_TRY(bad.download(&hbad, 1, c->stream));
  if (hbad) return hf_fail(HF_ERR_STATE, "ensemble operator has a non-positive diagonal at row " + std::to_string(hbad - 1));
  // meshes that fit on chip: one cooperative launch per solve for the whole tile (hf_enspatch.cu)
  const int mode = c->force_mode >= 0 ? c->force_mode : c->mode;
  if (mode == 0 || mode >= 3) HF_TRY(hf_ens_oc_plan(c, e));
  return HF_OK;
}

static const int kEnsChunk[3] = {8, 32, 128};

static int ens_set_smem(EnsState* e) {
  const int sm = (int)e->iter_smem;
  ENS_DISPATCH(e->LB, HF_CUDA(cudaFuncSetAttribute(k_ens_iter<LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm)));
  return HF_OK;
}

static void ens_launch_iteration(hf_ctx* c, EnsState* e, int par) {
  const EnsPatch P{e->nchunks,    e->mcap,         e->halo_cap, e->nstages, (unsigned)e->stage_bytes, e->halo_ptr.p,
                   e->halo_idx.p, e->rowptr_pad.p, e->lc_off.p, e->lcol.p,  e->bs.p};
  ENS_DISPATCH(e->LB, k_ens_iter<LB><<<e->igrid, HF_ENT, e->iter_smem, c->stream>>>(P, par, e->ks.p, e->drow.p, (size_t)e->nchunks * e->R, e->x.p, e->z.p, e->z1.p,
                                                                                   e->p0.p, e->p1.p, e->w.p, e->w1.p, e->part.p,
                                                                                   e->ctrl.p));
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

// Projection of the initial guess of all variants onto their recycled bases (hf_recycle.cu: hf_rc_project).
static int ens_rc_project(hf_ctx* c, EnsState* e, size_t nb) {
  if (e->rc_cap == 0) return HF_OK;
  const int m = std::min(e->rc_count, e->rc_cap);
  double* slotW = (m < e->rc_cap) ? e->rc_W.p + (size_t)m * nb : nullptr;
  if (m == 0) {
    HF_CUDA(cudaMemcpyAsync(slotW, e->x.p, sizeof(double) * nb, cudaMemcpyDeviceToDevice, c->stream));
    return HF_OK;
  }
  const int B = e->B;
  ENS_DISPATCH(e->LB, k_erc_dots<LB><<<e->rc_nseg, ERC_WARPS * 32, 0, c->stream>>>(m, e->rc_nseg, nb, e->rc_W.p, e->z.p, e->dg.p,
                                                                                  e->rc_parts.p));
  ENS_DISPATCH(e->LB, k_erc_coef<LB><<<(m + ERC_WARPS - 1) / ERC_WARPS, ERC_WARPS * 32, 0, c->stream>>>(m, e->rc_nseg, e->rc_parts.p,
                                                                                                       e->rc_inv.p, 1.0, e->rc_coef.p));
  ENS_DISPATCH(e->LB, k_erc_update<false, LB><<<e->grid, HF_ET, sizeof(double) * m * B, c->stream>>>(
                          m, nb, nb, e->rc_W.p, e->rc_AW.p, e->rc_coef.p, e->x.p, e->z.p, slotW, nullptr, e->dg.p, nullptr));
  c->stat_launches += 3;
  HF_CUDA(cudaGetLastError());
  return HF_OK;
}

// Store the A_b-orthogonalised correction of the solve that just finished (hf_recycle.cu: hf_rc_store).
static int ens_rc_store(hf_ctx* c, EnsState* e, size_t nb) {
  if (e->rc_cap == 0 || e->rc_count >= e->rc_cap) return HF_OK;
  const int m = e->rc_count, B = e->B, N = c->N;
  double* slotW = e->rc_W.p + (size_t)m * nb;
  double* slotAW = e->rc_AW.p + (size_t)m * nb;
  const unsigned blocks = (unsigned)(((size_t)N * B + HF_ET - 1) / HF_ET);
  ENS_DISPATCH(e->LB, k_erc_spmv<LB><<<blocks, HF_ET, 0, c->stream>>>(N, c->rowptr.p, c->col.p, e->base0.p, e->s0.p, c->bcflag.p,
                                                                     e->ks.p, e->dg.p, e->x.p, slotW, e->rc_d.p, e->rc_ad.p));
  if (m > 0) {
    // h = (A_b w_k) . d = sum AW_k * d * dg
    ENS_DISPATCH(e->LB, k_erc_dots<LB><<<e->rc_nseg, ERC_WARPS * 32, 0, c->stream>>>(m, e->rc_nseg, nb, e->rc_AW.p, e->rc_d.p, e->dg.p,
                                                                                    e->rc_parts.p));
    ENS_DISPATCH(e->LB, k_erc_coef<LB><<<(m + ERC_WARPS - 1) / ERC_WARPS, ERC_WARPS * 32, 0, c->stream>>>(
                            m, e->rc_nseg, e->rc_parts.p, e->rc_inv.p, -1.0, e->rc_coef.p));
    c->stat_launches += 2;
  }
  ENS_DISPATCH(e->LB, k_erc_update<true, LB><<<e->grid, HF_ET, sizeof(double) * m * B, c->stream>>>(
                          m, nb, nb, e->rc_W.p, e->rc_AW.p, e->rc_coef.p, e->rc_d.p, e->rc_ad.p, slotW, slotAW, e->dg.p, e->rc_part_nn.p));
  ENS_DISPATCH(e->LB, k_erc_norm<LB><<<1, 32, 0, c->stream>>>(e->grid, e->rc_part_nn.p, e->rc_inv.p, m));
  c->stat_launches += 3;
  HF_CUDA(cudaGetLastError());
  e->rc_count += 1;
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
  const size_t nbp = (size_t)e->nchunks * HF_EPAIRS;     // padded vector length (a multiple of ERC_SEG)
  if (c->rc.cap > 0 && e->rc_cap == 0) {
    size_t free_b = 0, total_b = 0;
    HF_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t fit = (free_b / 2) / (16 * nbp);         // W + AW, at most half of the free memory
    e->rc_cap = (int)std::min<size_t>({(size_t)c->rc.cap, fit, (size_t)(40960 / (8 * B))});   // coefficients live in shared memory
    if (e->rc_cap > 0) {
      e->rc_nseg = (int)(nbp / ERC_SEG);
      HF_TRY(e->rc_W.alloc((size_t)e->rc_cap * nbp, c->stream));
      HF_TRY(e->rc_AW.alloc((size_t)e->rc_cap * nbp, c->stream));
      HF_TRY(e->rc_inv.alloc((size_t)e->rc_cap * B, c->stream));
      HF_TRY(e->rc_coef.alloc((size_t)e->rc_cap * B, c->stream));
      HF_TRY(e->rc_parts.alloc((size_t)e->rc_cap * e->rc_nseg * B, c->stream));
      HF_TRY(e->rc_part_nn.alloc((size_t)e->grid * B, c->stream));
      HF_TRY(e->rc_d.alloc(nbp, c->stream));
      HF_TRY(e->rc_ad.alloc(nbp, c->stream));
      e->rc_count = 0;
    }
  }
  bool on_chip = e->oc_ok;
  if (on_chip) {
    // asynchronous solves: iteration counts and failures are read once at the end; the state at the start of the
    // run is kept so that a failed run can be repeated with the host-polled streaming kernels (as hf_run does)
    if (e->oc_iters.n < (size_t)std::max(1, n_steps)) HF_TRY(e->oc_iters.alloc(std::max(1, n_steps), c->stream));
    HF_CUDA(cudaMemsetAsync(e->oc_fail.p, 0, sizeof(int), c->stream));
    if (e->oc_u0.n != 2 * nbp) HF_TRY(e->oc_u0.alloc(2 * nbp, c->stream));
    HF_CUDA(cudaMemcpyAsync(e->oc_u0.p, e->u.p, sizeof(double) * nbp, cudaMemcpyDeviceToDevice, c->stream));
    HF_CUDA(cudaMemcpyAsync(e->oc_u0.p + nbp, e->uprev.p, sizeof(double) * nbp, cudaMemcpyDeviceToDevice, c->stream));
  }
  const bool had_prev = e->have_prev;
  c->stat_solve_ms = 0.0;
  c->stat_solve_launches = 0;
  if (c->profile)                                           // hf_set_profile: CUDA events around every solve
    while (c->prof_ev.size() < (size_t)2 * n_steps) {
      cudaEvent_t ev;
      HF_CUDA(cudaEventCreate(&ev));
      c->prof_ev.push_back(ev);
    }
  HF_CUDA(cudaEventRecord(c->ev0, c->stream));
run_again:
  for (int s = 0; s < n_steps; ++s) {
    if (c->n_gauss) {
      const int n = c->n_gauss * B;
      ENS_DISPATCH(e->LB, k_ens_gauss<LB><<<(n + 255) / 256, 256, 0, c->stream>>>(c->n_gauss, c->gauss_dof.p, c->gauss_r.p, amp[s], t_ic,
                                                                                 e->coeff.p, e->g.p));
      c->stat_launches += 1;
    }
    ENS_DISPATCH(e->LB, k_ens_init<LB><<<e->grid, HF_ET, 0, c->stream>>>(N, c->rowptr.p, c->col.p, c->valM.p, e->base0.p, e->s0.p,
                                                                        c->bcflag.p, e->ks.p, e->g.p, e->u.p, e->uprev.p,
                                                                        e->have_prev ? c->warm : 0.0, e->dg.p, e->x.p, e->z.p,
                                                                        e->part.p, e->ctrl.p, c->rtol));
    c->stat_launches += 1;
    HF_CUDA(cudaGetLastError());
    HF_TRY(ens_rc_project(c, e, nbp));
    int it = 0;
    const bool prof = c->profile && (size_t)(2 * s + 1) < c->prof_ev.size();
    if (prof) HF_CUDA(cudaEventRecord(c->prof_ev[2 * s], c->stream));
    if (on_chip) HF_TRY(hf_ens_oc_solve_async(c, e, s));
    else HF_TRY(ens_solve(c, e, &it));
    if (prof) HF_CUDA(cudaEventRecord(c->prof_ev[2 * s + 1], c->stream));
    HF_TRY(ens_rc_store(c, e, nbp));
    if (iters && !on_chip) iters[s] = it;
    if (n_watch) {
      const int n = n_watch * B;
      ENS_DISPATCH(e->LB, k_ens_sample<LB><<<(n + 255) / 256, 256, 0, c->stream>>>(n_watch, n_steps, s, e->watch.p, e->x.p, e->hist.p));
      c->stat_launches += 1;
    }
    std::swap(e->u.p, e->uprev.p);      // u_{n-1} <- u_n (neither array is an argument of the captured graphs)
    HF_CUDA(cudaMemcpyAsync(e->u.p, e->x.p, sizeof(double) * nb, cudaMemcpyDeviceToDevice, c->stream));
    e->have_prev = true;
  }
  if (on_chip && n_steps) {
    int nfail = 0;
    HF_TRY(e->oc_fail.download(&nfail, 1, c->stream));
    if (nfail) {
      // an on-chip solve hit the iteration cap or left the fixed-point range of its reduction: the state it left is
      // not trustworthy, the whole run is repeated from its initial state with the streaming kernels (still the GPU)
      c->stat_retries += 1;
      if (getenv("HF_DEBUG")) {
        HF_CUDA(cudaMemcpyAsync(e->h_ctrl, e->ctrl.p, sizeof(EnsCtrl), cudaMemcpyDeviceToHost, c->stream));
        HF_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<int> hit(n_steps);
        e->oc_iters.download(hit.data(), n_steps, c->stream);
        fprintf(stderr, "[hf debug] on-chip ensemble run failed: fail=%d last it=%d done=%d rz=%g %g %g %g thr=%g active=%d%d%d%d iters:",
                nfail, e->h_ctrl->it, e->h_ctrl->done, e->h_ctrl->rz[0], e->h_ctrl->rz[1], e->h_ctrl->rz[2], e->h_ctrl->rz[3],
                e->h_ctrl->thr[0], e->h_ctrl->active[0], e->h_ctrl->active[1], e->h_ctrl->active[2], e->h_ctrl->active[3]);
        for (int s = 0; s < n_steps; ++s) fprintf(stderr, " %d", hit[s]);
        fprintf(stderr, "\n");
      }
      // the failed run may have left non-finite values behind, also in the padding rows that the recycle kernels sum over
      for (DevBuf<double>* v : {&e->x, &e->z, &e->z1, &e->p0, &e->p1, &e->w, &e->w1, &e->rc_d, &e->rc_ad})
        if (v->n) HF_CUDA(cudaMemsetAsync(v->p, 0, v->n * sizeof(double), c->stream));
      HF_CUDA(cudaMemcpyAsync(e->u.p, e->oc_u0.p, sizeof(double) * nbp, cudaMemcpyDeviceToDevice, c->stream));
      HF_CUDA(cudaMemcpyAsync(e->uprev.p, e->oc_u0.p + nbp, sizeof(double) * nbp, cudaMemcpyDeviceToDevice, c->stream));
      e->have_prev = had_prev;
      e->rc_count = 0;                                       // the recycled basis of the failed run is dropped as well
      on_chip = false;
      goto run_again;
    }
    std::vector<int> hit(n_steps);
    HF_TRY(e->oc_iters.download(hit.data(), n_steps, c->stream));
    for (int s = 0; s < n_steps; ++s) {
      if (iters) iters[s] = hit[s];
      c->stat_iters += (unsigned long long)hit[s];
    }
    e->last_iters = hit[n_steps - 1];
  }
  HF_CUDA(cudaEventRecord(c->ev1, c->stream));
  if (n_watch && n_steps)   // device layout [B, S, W]; only the first nb variants are real
    HF_CUDA(cudaMemcpyAsync(hist, e->hist.p, sizeof(double) * (size_t)e->nb * n_steps * n_watch, cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  HF_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->stat_run_ms = ms;
  if (c->profile)
    for (int s = 0; s < n_steps; ++s) {
      float t = 0.f;
      HF_CUDA(cudaEventElapsedTime(&t, c->prof_ev[2 * s], c->prof_ev[2 * s + 1]));
      c->stat_solve_ms += t;
    }
  return HF_OK;
}

extern "C" int hf_ens_get_state(hf_ctx* c, double* u) {
  if (!c || !c->ens || !u) return hf_fail(HF_ERR_STATE, "hf_ens_get_state: no ensemble / null argument");
  cudaSetDevice(c->device);
  EnsState* e = c->ens;
  const size_t nb = (size_t)c->N * e->B;
  if (e->stage.n < nb) HF_TRY(e->stage.alloc(nb, c->stream));
  ENS_DISPATCH(e->LB, k_ens_transpose<LB><<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(c->N, c->permuted ? c->rank_d.p : nullptr,
                                                                                            e->u.p, e->stage.p));
  HF_CUDA(cudaGetLastError());
  return e->stage.download(u, (size_t)c->N * e->nb, c->stream);
}

extern "C" int hf_ens_get_path(hf_ctx* c) {
  if (!c || !c->ens) return hf_fail(HF_ERR_STATE, "hf_ens_get_path: no ensemble");
  return c->ens->oc_ok ? 5 : 1;
}

extern "C" int hf_ens_destroy(hf_ctx* c) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hf_ens_free(c);
  return HF_OK;
}
