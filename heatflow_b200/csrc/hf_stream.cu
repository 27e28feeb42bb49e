// heatflow_b200 - Jacobi-preconditioned CG, persistent streaming kernel: ONE cooperative launch per solve
// (sm_100a).
//
// Replaces KSP PREONLY + PC LU / MUMPS (reference: run_with_diamond.py:389-394, :480) for meshes whose
// operator does not fit on chip.  Same algorithm, same patch decomposition, same shared-memory stages and the
// same arithmetic - bit for bit - as the one-launch-per-iteration kernel k_pcg_iter (hf_pcg.cu); what changes
// is who drives the iteration loop:
//   * the loop over the PCG iterations runs inside the kernel; iterations are separated by ONE grid barrier
//     (release / acquire on a monotonic 64-bit arrival counter) that doubles as the reduction: every CTA
//     publishes its four partial sums before it arrives and adds all CTAs' partials, in CTA order, after it
//     has passed - so alpha, beta and the stopping test are the same bits in every CTA and nothing is
//     broadcast.  No host polling, no launches queued past convergence, no control-block reads;
//   * the operator is iteration invariant, so the operator part of the first chunks of iteration n+1 streams
//     into the shared-memory stages that free up at the end of iteration n, ACROSS the barrier; only the
//     32 B/row of vector data wait for it.  Each stage's mbarrier therefore takes two arrivals per use
//     (operator part, vector part).
// Data written by other SMs is read through the L2: TMA for the own rows (ordered after the writers' generic
// stores by fence.proxy.async + the gpu-scope barrier), ld.global.cg for the halo gathers.
#include <algorithm>
#include <cmath>

#include "hf_ctx.cuh"
#include "hf_tma.cuh"

struct StreamArgs {
  PatchView A;
  IterStage S;
  double* x;
  double* rb0;
  double* rb1;
  double* pb0;
  double* pb1;
  double* qb0;
  double* qb1;
  double* parts;                 // [2][4][G] per-CTA partial sums, double buffered by iteration parity
  HfCtrl* c;
  unsigned long long* bar;       // [0] arrival counter (monotonic), [1] its value when the last launch ended
  int* iters_out;
  int* fail;
  int max_it, nparts;
  double rtol;
};

__device__ __forceinline__ void hf_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hf_smem_u32(bar)) : "memory");
}

// operator part of a chunk: values, 16-bit local columns, slice pointers
template <int R>
__device__ __forceinline__ void hf_issue_operator(const PatchView& A, int ch, int e0, int e1, unsigned char* st, int mat_cap,
                                                  int halo_cap, unsigned long long* bar) {
  constexpr int SPC = R / HF_SLICE;
  const unsigned n = (unsigned)(e1 - e0);
  unsigned char* scol = st + (size_t)mat_cap * 8 + 4 * (size_t)R * 8 + (size_t)halo_cap * 8;
  unsigned char* sptr = scol + (size_t)mat_cap * 2;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was last touched by ordinary loads/stores
  hf_mbar_expect(bar, n * 10u + (SPC + 4) * 4u);
  const unsigned long long stream = hf_policy_evict_first();
  hf_bulk_g2s(sptr, A.slice_ptr + (size_t)ch * SPC, (SPC + 4) * 4u, bar);
  for (unsigned off = 0; off < n * 8u; off += HF_BULK_PIECE)
    hf_bulk_g2s_hint(st + off, reinterpret_cast<const unsigned char*>(A.val + e0) + off, min(HF_BULK_PIECE, n * 8u - off), bar, stream);
  for (unsigned off = 0; off < n * 2u; off += HF_BULK_PIECE)
    hf_bulk_g2s_hint(scol + off, reinterpret_cast<const unsigned char*>(A.lcol + e0) + off, min(HF_BULK_PIECE, n * 2u - off), bar,
                     stream);
}

// vector part of a chunk: x, r, q, p on the own rows
template <int R>
__device__ __forceinline__ void hf_issue_vectors(int ch, unsigned char* st, int mat_cap, const double* x, const double* ro,
                                                 const double* po, const double* qo, unsigned long long* bar) {
  const unsigned vb = (unsigned)R * 8u;
  unsigned char* sx = st + (size_t)mat_cap * 8;
  const size_t lo = (size_t)ch * R;
  const unsigned long long keep = hf_policy_evict_last();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  hf_mbar_expect(bar, 4u * vb);
  hf_bulk_g2s_hint(sx, x + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + vb, ro + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + 2 * vb, qo + lo, vb, bar, keep);
  hf_bulk_g2s_hint(sx + 3 * vb, po + lo, vb, bar, keep);
}

__device__ __forceinline__ unsigned long long hf_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <int R>
__global__ void __launch_bounds__(HF_IT, HF_IT_MINB) k_pcg_stream(StreamArgs P) {
  constexpr int SPC = R / HF_SLICE;
  constexpr int NW = HF_IT / 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ double sh4[4][NW];
  __shared__ double s_tot[4];
  __shared__ __align__(8) unsigned long long full[4];
  const PatchView& A = P.A;
  const IterStage& S = P.S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, bid = blockIdx.x;
  const int nloc = (A.nchunks - bid + G - 1) / G;                 // chunks of this CTA (>= 1: the host launches G <= nchunks)
  const int nst = min(S.nstages, nloc);                           // a stage is refilled at most one iteration ahead
  // ---- threshold and ||r_0||^2: the same sums, in the same order, as k_pcg_ctrl_init
  double thr, rr;
  {
    double* red = &sh4[0][0];
    if (P.nparts > 0) {
      const double bn2 = hf_sum_parts(P.c->part_bn, P.nparts, red);
      rr = hf_sum_parts(P.c->part_rr[0], P.nparts, red);
      thr = P.rtol * P.rtol * bn2;
      if (bid == 0 && tid == 0) {
        P.c->bn2 = bn2;
        P.c->thr = thr;
      }
    } else {
      thr = P.c->thr;
      rr = P.c->rr;
    }
    __syncthreads();
  }
  bool done = !(rr > thr);
  int it = 0;
  unsigned long long target = 0ull;                               // thread 0: arrival count that completes the next barrier
  if (!done && P.max_it > 0) {
    if (tid == 0) {
      target = P.bar[1];
      for (int s = 0; s < nst; ++s) hf_mbar_init(&full[s], 2);
      int e0[4], e1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = min(bid + j * G, A.nchunks - 1);
        e0[j] = A.slice_ptr[ch * SPC];
        e1[j] = A.slice_ptr[(ch + 1) * SPC];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nst) {
          unsigned char* st = smraw + (size_t)j * S.stage_bytes;
          hf_issue_operator<R>(A, bid + j * G, e0[j], e1[j], st, S.mat_cap, S.halo_cap, &full[j]);
          hf_issue_vectors<R>(bid + j * G, st, S.mat_cap, P.x, P.rb0, P.pb0, P.qb0, &full[j]);
        }
    }
    // ---- halo pipeline (iteration invariant part): list extents of chunks 0..2, node indices of chunks 0..1.
    // The pipeline wraps around at the end of an iteration, so these are loaded once per solve.
    int hp_a = A.halo_ptr[bid], he_a = A.halo_ptr[bid + 1];
    const int c1 = bid + (1 % nloc) * G, c2 = bid + (2 % nloc) * G;
    int hp_b = A.halo_ptr[c1], he_b = A.halo_ptr[c1 + 1];
    int hp_c = A.halo_ptr[c2], he_c = A.halo_ptr[c2 + 1];
    int nh_a = he_a - hp_a;
    int g_a = (tid < nh_a) ? A.halo_idx[hp_a + tid] : -1;        // indices of the chunk whose values are loaded next
    int g_b = (tid < he_b - hp_b) ? A.halo_idx[hp_b + tid] : -1;
    int jn3 = 3 % nloc;                                           // local index of the chunk whose extents are loaded next
    unsigned v = 0;                                               // visits (chunks processed) so far, over all iterations
    double alpha = 0.0, beta = 0.0;
    int par = 0;
    __syncthreads();                                              // mbarrier inits visible
    while (true) {
      const double* __restrict__ ro = par ? P.rb1 : P.rb0;
      const double* __restrict__ po = par ? P.pb1 : P.pb0;
      const double* __restrict__ qo = par ? P.qb1 : P.qb0;
      double* __restrict__ rn = par ? P.rb0 : P.rb1;
      double* __restrict__ pn = par ? P.pb0 : P.pb1;
      double* __restrict__ qn = par ? P.qb0 : P.qb1;
      // halo values of the first chunk (written by other CTAs in the previous iteration: through L2)
      double hr = 0.0, hp = 0.0, hq = 0.0;
      if (g_a >= 0) {
        hr = __ldcg(ro + g_a);
        hp = __ldcg(po + g_a);
        hq = __ldcg(qo + g_a);
      }
      double l_rr = 0.0, l_pq = 0.0, l_rq = 0.0, l_qq = 0.0;
      for (int j = 0; j < nloc; ++j, ++v) {
        const int ch = bid + j * G;
        const int stg = (int)(v % (unsigned)nst);
        const unsigned parity = (v / (unsigned)nst) & 1u;
        unsigned char* st = smraw + (size_t)stg * S.stage_bytes;
        const double* sval = reinterpret_cast<const double*>(st);
        double* sx = reinterpret_cast<double*>(st + (size_t)S.mat_cap * 8);
        double* sr = sx + R;
        double* sq = sr + R;
        double* sp = sq + R;                // own rows, halo follows at sp[R + h]
        const unsigned short* scol = reinterpret_cast<const unsigned short*>(sp + R + S.halo_cap);
        const int* sptr = reinterpret_cast<const int*>(scol + S.mat_cap);
        const int lo = ch * R;
        // thread 0: operator extent of the chunk that will refill this stage (needed only after phase 2)
        const int jr = (j + nst < nloc) ? j + nst : j + nst - nloc;   // local index of that chunk (next iteration when wrapped)
        int nx_e0 = 0, nx_e1 = 0;
        if (tid == 0) {
          const int chn = bid + jr * G;
          nx_e0 = A.slice_ptr[chn * SPC];
          nx_e1 = A.slice_ptr[(chn + 1) * SPC];
        }
        hf_mbar_wait(&full[stg], parity);
        // ---- phase 1: finish iteration n-1 on the own rows (from the stage) and on the halo (from registers)
        if (tid < R) {
          const double rv = sr[tid];
          double r_new = rv, p_new = rv;
          if (it > 0) {
            const double pv = sp[tid];
            P.x[lo + tid] = fma(alpha, pv, sx[tid]);
            r_new = fma(-alpha, sq[tid], rv);
            p_new = fma(beta, pv, r_new);
          }
          rn[lo + tid] = r_new;
          pn[lo + tid] = p_new;
          sp[tid] = p_new;
          sr[tid] = r_new;
          l_rr = fma(r_new, r_new, l_rr);
        }
        if (tid < nh_a) sp[R + tid] = (it > 0) ? fma(beta, hp, fma(-alpha, hq, hr)) : hr;
        for (int h = HF_IT + tid; h < nh_a; h += HF_IT) {       // oversized halos (poor node order): synchronous
          const int g = A.halo_idx[hp_a + h];
          const double r_old = __ldcg(ro + g);
          sp[R + h] = (it > 0) ? fma(beta, __ldcg(po + g), fma(-alpha, __ldcg(qo + g), r_old)) : r_old;
        }
        __syncthreads();
        // ---- advance the halo pipeline (all loads are consumed one chunk later); the values of the first chunk
        // of the next iteration are loaded after the barrier
        if (j + 1 < nloc) {
          if (g_b >= 0) {
            hr = __ldcg(ro + g_b);
            hp = __ldcg(po + g_b);
            hq = __ldcg(qo + g_b);
          }
        } else {
          g_a = g_b;
        }
        hp_a = hp_b;
        nh_a = he_b - hp_b;
        g_b = (tid < he_c - hp_c) ? A.halo_idx[hp_c + tid] : -1;
        hp_b = hp_c;
        he_b = he_c;
        {
          const int cn = bid + jn3 * G;
          hp_c = A.halo_ptr[cn];
          he_c = A.halo_ptr[cn + 1];
          jn3 = (jn3 + 1 == nloc) ? 0 : jn3 + 1;
        }
        // ---- phase 2: q = Ahat p, everything from shared memory
        if (warp < SPC) {
          const int e0 = sptr[0];
          const int base = sptr[warp] - e0;
          const int w = (sptr[warp + 1] - e0 - base) >> 5;     // 0 for the padding slices behind the last row
          const unsigned short* cp = scol + base + lane;
          const double* vp = sval + base + lane;
          double acc0 = 0.0, acc1 = 0.0;
          int k = 0;
          for (; k + 4 <= w; k += 4) {
            const int c0 = cp[k * 32], c1_ = cp[(k + 1) * 32], c2_ = cp[(k + 2) * 32], c3 = cp[(k + 3) * 32];
            const double v0 = vp[k * 32], v1 = vp[(k + 1) * 32], v2 = vp[(k + 2) * 32], v3 = vp[(k + 3) * 32];
            acc0 = fma(v0, sp[c0], acc0);
            acc1 = fma(v1, sp[c1_], acc1);
            acc0 = fma(v2, sp[c2_], acc0);
            acc1 = fma(v3, sp[c3], acc1);
          }
          for (; k < w; ++k) acc0 = fma(vp[k * 32], sp[cp[k * 32]], acc0);
          const double acc = acc0 + acc1;
          const int i = warp * HF_SLICE + lane;
          qn[lo + i] = acc;
          l_pq = fma(sp[i], acc, l_pq);
          l_rq = fma(sr[i], acc, l_rq);
          l_qq = fma(acc, acc, l_qq);
        }
        __syncthreads();                    // the stage is free again
        if (tid == 0) {
          const int chn = bid + jr * G;
          hf_issue_operator<R>(A, chn, nx_e0, nx_e1, st, S.mat_cap, S.halo_cap, &full[stg]);
          if (j + nst < nloc) hf_issue_vectors<R>(chn, st, S.mat_cap, P.x, ro, po, qo, &full[stg]);
        }
      }
      // ---- per-CTA partials, then the grid barrier
      {
        const double v0 = hf_warp_sum(l_rr), v1 = hf_warp_sum(l_pq), v2 = hf_warp_sum(l_rq), v3 = hf_warp_sum(l_qq);
        if (lane == 0) {
          sh4[0][warp] = v0;
          sh4[1][warp] = v1;
          sh4[2][warp] = v2;
          sh4[3][warp] = v3;
        }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");   // this thread's x r p q stores before any later TMA read
      __syncthreads();
      double* part = P.parts + (size_t)(it & 1) * 4 * G;
      if (warp == 0) {
        if (lane < 4) {
          double t = 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) t += sh4[lane][i];
          __stcg(part + lane * G + bid, t);
        }
        __syncwarp();
        if (lane == 0) {
          target += (unsigned long long)G;
          __threadfence();                  // the CTA's stores (ordered by the barrier above) and the partials before the arrival
          atomicAdd(P.bar, 1ull);
          while (hf_ld_acquire(P.bar) < target) {}
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
      }
      __syncthreads();
      // every CTA adds the partials in CTA order: same bits everywhere
      if (warp < 4) {
        double s = 0.0;
        for (int i = lane; i < G; i += 32) s += __ldcg(part + warp * G + i);
        s = hf_warp_sum(s);
        if (lane == 0) s_tot[warp] = s;
      }
      __syncthreads();
      rr = s_tot[0];
      if (!(rr > thr)) {                    // also stops on NaN
        done = true;
        break;
      }
      {
        const double al = rr / s_tot[1];
        const double rr_next = fma(al * al, s_tot[3], fma(-2.0 * al, s_tot[2], rr));
        alpha = al;
        beta = fmax(rr_next, 0.0) / rr;
      }
      ++it;
      par ^= 1;
      if (it >= P.max_it) break;
      // vector parts of the first chunks of the next iteration (their operator parts are already in flight)
      if (tid == 0) {
        const double* ro2 = par ? P.rb1 : P.rb0;
        const double* po2 = par ? P.pb1 : P.pb0;
        const double* qo2 = par ? P.qb1 : P.qb0;
        for (int j = 0; j < nst; ++j) {
          const int stg = (int)((v + (unsigned)j) % (unsigned)nst);
          hf_issue_vectors<R>(bid + j * G, smraw + (size_t)stg * S.stage_bytes, S.mat_cap, P.x, ro2, po2, qo2, &full[stg]);
        }
      }
    }
    // ---- drain the operator parts that were prefetched for an iteration that will not run
    if (tid == 0) {
      for (int j = 0; j < nst; ++j) {
        const unsigned vv = v + (unsigned)j;
        const int stg = (int)(vv % (unsigned)nst);
        hf_mbar_arrive(&full[stg]);
        hf_mbar_wait(&full[stg], (vv / (unsigned)nst) & 1u);
      }
      if (bid == 0) P.bar[1] = target;
    }
  }
  if (bid == 0 && tid == 0) {
    P.c->rr = rr;
    P.c->itA = it;
    P.c->done = done ? 1 : 0;
    if (P.iters_out) *P.iters_out = it;
    if (!done || !isfinite(rr)) atomicAdd(P.fail, 1);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static const void* stream_kernel(const SellOp&) { return (const void*)k_pcg_stream<256>; }

// Largest co-resident grid of the kernel for this operator's stage layout; 0 = cooperative launch not possible.
int hf_stream_plan(hf_ctx* c, SellOp& op) {
  op.st_grid = 0;
  if (op.R != 256 || op.nchunks == 0 || op.iter_smem == 0) return HF_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  if (!coop) return HF_OK;
  const void* fn = stream_kernel(op);
  HF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.iter_smem));
  int per_sm = 0;
  HF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, HF_IT, op.iter_smem));
  per_sm = std::min(per_sm, HF_IT_MINB);
  if (per_sm < 1) return HF_OK;
  op.st_grid = std::min(op.nchunks, c->sm_count * per_sm);
  PcgWork& w = c->ws;
  if (w.bar.n == 0) HF_TRY(w.bar.alloc(2, c->stream));
  if (w.fail.n == 0) HF_TRY(w.fail.alloc(1, c->stream));
  if (w.sparts.n < (size_t)8 * op.st_grid) HF_TRY(w.sparts.alloc((size_t)8 * op.st_grid, c->stream));
  return HF_OK;
}

int hf_stream_solve_async(hf_ctx* c, const SellOp& op, int step_slot, bool sum_parts) {
  PcgWork& w = c->ws;
  if (!op.st_grid) return hf_fail(HF_ERR_STATE, "persistent streaming PCG kernel is not available for this operator");
  StreamArgs a;
  a.A = op.patch();
  a.S = IterStage{op.mat_cap, op.halo_cap, op.nstages, (unsigned)op.stage_bytes};
  a.x = w.x.p;
  a.rb0 = w.r.p;
  a.rb1 = w.r1.p;
  a.pb0 = w.p0.p;
  a.pb1 = w.p1.p;
  a.qb0 = w.q.p;
  a.qb1 = w.q1.p;
  a.parts = w.sparts.p;
  a.c = w.ctrl.p;
  a.bar = w.bar.p;
  a.iters_out = (step_slot >= 0 && (size_t)step_slot < w.step_iters.n) ? w.step_iters.p + step_slot : nullptr;
  a.fail = w.fail.p;
  a.max_it = c->max_iters;
  a.nparts = sum_parts ? w.grid : 0;
  a.rtol = c->rtol;
  void* args[] = {&a};
  const void* fn = stream_kernel(op);
  HF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.iter_smem));   // per function, not per operator
  HF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(op.st_grid), dim3(HF_IT), args, op.iter_smem, c->stream));
  c->stat_launches += 1;
  return HF_OK;
}
