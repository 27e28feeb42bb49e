// heatflow_b200 - persistent cooperative Jacobi-PCG: one launch per linear solve (sm_100a).
//
// For the reference's own problem sizes (1e5 - 1.5e5 dofs) a PCG iteration moves only ~20 MB,
// so the streaming two-kernel iteration (hf_pcg.cu) is bound by launch and dependent-load
// latency, not by HBM.  Measured on B200: a grid-wide reduction through L2 costs ~1.5 us however
// it is organised, a gpu-scope fence ~0.4 us.  This kernel therefore runs the whole solve in
// one launch with ONE fence-free grid reduction per iteration:
//   * one CTA of 256 threads per SM; every thread owns SPW = 4 rows (consecutive sliced-ELL slices)
//     whose x lives in registers; the CTA's slice of the scaled operator sits in shared memory,
//   * ghost zones: every CTA keeps r and p for the whole contiguous column range its rows touch
//     (own rows + halo) in shared memory and applies the CG vector updates to the halo
//     redundantly, bit-identically to the owner.  The only vector exchanged is q = A p, written
//     as flag-with-data packets {lo32, gen, hi32, gen} (16-byte stores, the generation makes a
//     packet self-validating, so no fence and no barrier orders the exchange),
//   * the three dot products (p,q), (r,q), (q,q) of an iteration go through one grid reduction:
//     alpha = rr/(p,q) and ||r_new||^2 = rr - 2 alpha (r,q) + alpha^2 (q,q) (exact identity for
//     r_new = r - alpha q), hence beta, without a second reduction.  The recurrence carries an
//     absolute error ~eps * rr_ref, so it is replaced by the directly summed ||r||^2 every
//     HF_RR_CHECK iterations, whenever rr fell by 1e-4 since the last direct value, and before
//     convergence is accepted (CPU emulation: same iteration counts as textbook CG),
//   * reductions use per-CTA slots of flag-with-data packets; CTA 0 adds the partials in slot
//     order and broadcasts, so every CTA sees identical bits (bit-reproducible, no atomics).
// Launched with cudaLaunchCooperativeKernel (all CTAs co-resident).
#include <algorithm>
#include <cmath>

#include "hf_ctx.cuh"

struct PersistArgs {
  SellView A;
  double* x;            // in: xhat_0, out: xhat
  const double* r;      // in: rhat_0
  uint4* qpk;           // [2][Npad] q packets
  HfCtrl* c;            // thr, rr0 in; rr, itA, done out
  uint4* slots;         // reduction slots (HF_RED_MODE 0)
  unsigned long long* acc;       // [2 sets][8 replicas][16 words] fixed-point accumulators (HF_RED_MODE 1)
  unsigned long long* acc_prev;  // their values when the previous launch ended
  unsigned* gen;        // packet generation base, monotonic across launches
  const int2* cta_range;
  int* iters_out;       // may be null
  int* fail;            // incremented when max_it is hit
  int max_it;
  int npad;
  int nparts;           // > 0: thr and ||r_0||^2 are summed here from nparts partials of the control block
  double rtol;          //      (saves the k_pcg_ctrl_init launch in front of every solve)
  int mat_cap;          // elements of the CTA operator slice held in shared memory
  int sz_cap;           // elements of the ghost range
};

#include "hf_persist.cuh"


template <int SPW>
__global__ void __launch_bounds__(HF_PT, 1) k_pcg_persist(PersistArgs P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);
  double* sp = sval + P.mat_cap;       // p over the ghost range
  double* sr = sp + P.sz_cap;          // r over the ghost range
  double* sqh = sr + P.sz_cap;         // validated halo q values (nh <= sz_cap)
  double* red = sqh + P.sz_cap;        // reduction scratch: 2 x (HF_PW*3 + 3)
  int* scol = reinterpret_cast<int*>(red + 2 * (HF_PW * 3 + 3) + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, nsl = P.A.nslices;
  const int f = blockIdx.x * HF_PW * SPW;
  const int fend = min(f + HF_PW * SPW, nsl);
  const int e0 = P.A.slice_ptr[f], e1 = P.A.slice_ptr[fend];
  const int2 rng = P.cta_range[blockIdx.x];
  const int lo = rng.x, nr = rng.y - rng.x;
  for (int i = tid; i < e1 - e0; i += HF_PT) {
    sval[i] = P.A.val[e0 + i];
    scol[i] = P.A.col[e0 + i] - lo;    // columns relative to the ghost range
  }
  for (int i = tid; i < nr; i += HF_PT) {
    const double rv = P.r[lo + i];
    sr[i] = rv;
    sp[i] = rv;                        // p_0 = r_0
  }
  const int own0 = f * 32 - lo, own1 = fend * 32 - lo;   // own rows occupy [own0, own1) of the range
  const int nh = nr - (own1 - own0);                     // halo entries

  int base[SPW], wid[SPW], idx[SPW];
  double x[SPW], q[SPW];
#pragma unroll
  for (int k = 0; k < SPW; ++k) {
    const int s = f + warp * SPW + k;
    base[k] = 0;
    wid[k] = -1;                       // marks "no slice"
    idx[k] = 0;
    x[k] = q[k] = 0.0;
    if (s < nsl) {
      const int b0 = P.A.slice_ptr[s];
      base[k] = b0 - e0 + lane;
      wid[k] = (P.A.slice_ptr[s + 1] - b0) >> 5;
      idx[k] = s * 32 + lane - lo;
      x[k] = P.x[s * 32 + lane];
    }
  }
  double thr, rr;                      // stopping threshold and ||r_0||^2
  if (P.nparts > 0) {                  // same sums, in the same order, as k_pcg_ctrl_init (identical in every CTA)
    const double bn2 = hf_sum_parts(P.c->part_bn, P.nparts, red);
    rr = hf_sum_parts(P.c->part_rr[0], P.nparts, red);
    thr = P.rtol * P.rtol * bn2;
    if (blockIdx.x == 0 && tid == 0) {
      P.c->bn2 = bn2;
      P.c->thr = thr;
    }
  } else {
    thr = P.c->thr;
    rr = P.c->rr;
  }
  unsigned gen = *P.gen;
  __syncthreads();
  // the first HF_WR entries of every own row are cached in registers (zero-padded)
  double mv[SPW][HF_WR];
  int mc[SPW][HF_WR];
#pragma unroll
  for (int k = 0; k < SPW; ++k)
#pragma unroll
    for (int kk = 0; kk < HF_WR; ++kk) {
      const bool in = kk < wid[k];
      mv[k][kk] = in ? sval[base[k] + kk * 32] : 0.0;
      mc[k][kk] = in ? scol[base[k] + kk * 32] : idx[k];
    }

  int it = 0;
  int since_check = 0;
  double rr_ref = rr;
  bool done = !(rr > thr);
#if HF_RED_MODE == 1
  // magnitude estimates for the fixed-point reductions, identical in every CTA: ||p||^2 follows
  // pp_n = rr_n + beta_n^2 pp_{n-1} (r_n is orthogonal to p_{n-1}); |p.q| <= L pp, q.q <= L^2 pp,
  // |r.q| <= L sqrt(rr pp) with L = 16 >= ||Ahat||_inf (|ahat_ij| <= 1, at most 16 entries per row)
  double pp = rr;
  FxState fx;
  hf_fx_load_state(fx, P.acc_prev);
#endif
  while (!done && it < P.max_it) {
    ++gen;
    uint4* qout = P.qpk + (size_t)(it & 1) * P.npad;   // double buffered by iteration parity
    // ---- q = A p on the own rows; publish q packets; partial dots
    double d[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
      if (wid[k] >= 0) {
        double acc = 0.0;
        const int w = wid[k], b = base[k];
#pragma unroll
        for (int kk = 0; kk < HF_WR; ++kk) acc = fma(mv[k][kk], sp[mc[k][kk]], acc);
        for (int kk = HF_WR; kk < w; ++kk) acc = fma(sval[b + kk * 32], sp[scol[b + kk * 32]], acc);
        q[k] = acc;
        hf_pkt_store(qout + lo + idx[k], acc, gen);
        const double pv = sp[idx[k]], rv = sr[idx[k]];
        d[0] = fma(pv, acc, d[0]);
        d[1] = fma(rv, acc, d[1]);
        d[2] = fma(acc, acc, d[2]);
      }
    }
#if HF_RED_MODE == 1
    const int e_pp = hf_exp2(pp), e_rr = hf_exp2(rr);
    const int eb3[3] = {hf_clamp_exp(e_pp + 4 + HF_FX_MARGIN), hf_clamp_exp((e_rr + e_pp + 1) / 2 + 5 + HF_FX_MARGIN),
                        hf_clamp_exp(e_pp + 8 + HF_FX_MARGIN)};
    hf_fx_arrive<3>(d, eb3, P.acc, gen, red);
#else
    hf_grid_arrive<3>(d, P.slots, gen, red);
#endif
    // ---- halo q packets: warps >= 3 fetch them (all loads of a round in flight together),
    // re-poll until the generation matches and park the values in shared memory; this overlaps
    // the reduction, which warps 0..2 poll
    if (warp >= 3) {
      const int t0h = tid - 3 * 32;
      for (int h0 = t0h; h0 < nh; h0 += 6 * (HF_PT - 96)) {
        uint4 hq[6];
        bool need[6];
#pragma unroll
        for (int t = 0; t < 6; ++t) need[t] = (h0 + t * (HF_PT - 96)) < nh;
        bool pending;
        do {
          pending = false;
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            const int h = h0 + t * (HF_PT - 96);
            if (need[t]) hq[t] = hf_pkt_load(qout + lo + ((h < own0) ? h : h + (own1 - own0)));
          }
#pragma unroll
          for (int t = 0; t < 6; ++t)
            if (need[t]) {
              if (hf_pkt_ok(hq[t], gen)) {
                need[t] = false;
                sqh[h0 + t * (HF_PT - 96)] = hf_pkt_val(hq[t]);
              } else {
                pending = true;
              }
            }
        } while (pending);
      }
    }
    double tot[3];
#if HF_RED_MODE == 1
    hf_fx_wait<3>(tot, eb3, P.acc, G, gen, red, fx);
#else
    hf_grid_wait<3>(tot, P.slots, G, gen, red);
#endif
    const double alpha = rr / tot[0];
    double rr_new = fma(alpha * alpha, tot[2], fma(-2.0 * alpha, tot[1], rr));
    ++since_check;
    // the recurrence carries an absolute error ~eps * rr_ref: recompute directly every HF_RR_CHECK
    // iterations, whenever rr has dropped by 1e-4 since the last direct value, and before
    // convergence is accepted
    const bool check = (since_check >= HF_RR_CHECK) || !(rr_new > thr) || (rr_new < 1e-4 * rr_ref);
    const double beta = check ? 0.0 : rr_new / rr;
    // ---- own rows: x += alpha p ; r -= alpha q ; p = r + beta p (deferred when beta is not final)
    double dd[1] = {0.0};
#pragma unroll
    for (int k = 0; k < SPW; ++k) {
      if (wid[k] >= 0) {
        const int i = idx[k];
        const double pv = sp[i];
        x[k] = fma(alpha, pv, x[k]);
        const double rv = fma(-alpha, q[k], sr[i]);
        sr[i] = rv;
        dd[0] = fma(rv, rv, dd[0]);
        if (!check) sp[i] = fma(beta, pv, rv);
      }
    }
    // ---- halo rows: the same update with the neighbours' q
    for (int h = tid; h < nh; h += HF_PT) {
      const int i = (h < own0) ? h : h + (own1 - own0);
      const double rv = fma(-alpha, sqh[h], sr[i]);
      sr[i] = rv;
      if (!check) sp[i] = fma(beta, sp[i], rv);
    }
    if (check) {
      // replace the recurrence value by the directly summed ||r||^2, then finish the p update
      ++gen;
      double t1[1];
#if HF_RED_MODE == 1
      // ||r - alpha q||^2 <= 2 (rr + alpha^2 256 pp)
      const int eb1[1] = {hf_clamp_exp(max(e_rr, 2 * hf_exp2(fabs(alpha)) + 8 + e_pp) + 2 + HF_FX_MARGIN)};
      hf_fx_arrive<1>(dd, eb1, P.acc, gen, red);
      hf_fx_wait<1>(t1, eb1, P.acc, G, gen, red, fx);
#else
      hf_grid_arrive<1>(dd, P.slots, gen, red);
      hf_grid_wait<1>(t1, P.slots, G, gen, red);
#endif
      rr_new = t1[0];
      rr_ref = rr_new;
      since_check = 0;
      const double b2 = rr_new / rr;
      for (int i = tid; i < nr; i += HF_PT) sp[i] = fma(b2, sp[i], sr[i]);
      done = !(rr_new > thr);
#if HF_RED_MODE == 1
      pp = fma(b2 * b2, pp, fabs(rr_new));
#endif
    }
#if HF_RED_MODE == 1
    else pp = fma(beta * beta, pp, fabs(rr_new));
#endif
    rr = rr_new;
    ++it;
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < SPW; ++k)
    if (wid[k] >= 0) P.x[lo + idx[k]] = x[k];
#if HF_RED_MODE == 1
  hf_fx_store_state(fx, P.acc_prev);   // accumulator values the next launch starts from
#endif
  if (blockIdx.x == 0 && tid == 0) {
    P.c->rr = rr;
    P.c->itA = it;
    P.c->done = done ? 1 : 0;
    if (P.iters_out) *P.iters_out = it;
    if (!done || !isfinite(rr)) atomicAdd(P.fail, 1);   // iteration cap, or a non-finite residual / reduction overflow
    *P.gen = gen;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static size_t persist_smem_bytes(int mat_cap, int sz_cap) {
  return sizeof(double) * ((size_t)mat_cap + 3 * (size_t)sz_cap + 2 * (HF_PW * 3 + 3) + 2) + sizeof(int) * (size_t)mat_cap;
}

// Decide SPW / grid / shared-memory layout for an operator; spw = 0 when the mesh does not fit.
int hf_persist_plan(hf_ctx* c, SellOp& op) {
  op.p_spw = 0;
  const int nsl = op.nslices;
  if (nsl == 0 || c->h_rowptr.empty()) return HF_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  if (!coop) return HF_OK;
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  std::vector<int> sp(nsl + 1);
  HF_CUDA(cudaMemcpyAsync(sp.data(), op.slice_ptr.p, sizeof(int) * (nsl + 1), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  for (int spw = HF_SPW; spw <= HF_SPW; spw += 2) {
    const int per = HF_PW * spw;
    const int G = (nsl + per - 1) / per;
    if (G > c->sm_count || G > HF_MAX_GRID) continue;
    int mat_cap = 0, sz_cap = 0;
    std::vector<int2> range(G);
    for (int b = 0; b < G; ++b) {
      const int f = b * per, fe = std::min(f + per, nsl);
      mat_cap = std::max(mat_cap, sp[fe] - sp[f]);
      int lo = f * 32, hi = fe * 32;             // own (padded) rows are always inside the range
      const int r0 = f * 32, r1 = std::min(fe * 32, c->N);
      if (r0 < r1) {
        for (int k = c->h_rowptr[r0]; k < c->h_rowptr[r1]; ++k) {
          lo = std::min(lo, c->h_col[k]);
          hi = std::max(hi, c->h_col[k] + 1);
        }
      }
      range[b] = make_int2(lo, hi);
      sz_cap = std::max(sz_cap, hi - lo);
    }
    mat_cap = (mat_cap + 1) & ~1;
    sz_cap = (sz_cap + 1) & ~1;
    const size_t bytes = persist_smem_bytes(mat_cap, sz_cap);
    if (bytes > (size_t)max_smem) continue;
    op.p_spw = spw;
    op.p_grid = G;
    op.p_mat_cap = mat_cap;
    op.p_sz_cap = sz_cap;
    op.p_smem = bytes;
    HF_TRY(op.p_range.upload(range.data(), G, c->stream));
    HF_CUDA(cudaFuncSetAttribute(k_pcg_persist<HF_SPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    PcgWork& w = c->ws;
    if (w.slots.n < (size_t)2 * (HF_MAX_GRID + 1) * HF_SLOT_STRIDE) {
      HF_TRY(w.slots.alloc((size_t)2 * (HF_MAX_GRID + 1) * HF_SLOT_STRIDE, c->stream));
      HF_TRY(w.gen.alloc(1, c->stream));
      HF_TRY(w.fail.alloc(1, c->stream));
      HF_TRY(w.acc.alloc((size_t)2 * HF_NREP * HF_ACC_LINE, c->stream));
      HF_TRY(w.acc_prev.alloc((size_t)2 * HF_NREP * HF_ACC_LINE, c->stream));
    }
    if (w.qpk.n < (size_t)2 * c->Npad) HF_TRY(w.qpk.alloc((size_t)2 * c->Npad, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    break;
  }
  return HF_OK;
}

// Launch the whole solve; no host synchronisation.  step_slot >= 0 stores the iteration count
// in ws.step_iters[step_slot].  On entry ws.x = xhat_0, ws.r = rhat_0 and the control block holds
// thr and rr (= ||rhat_0||^2) from k_pcg_ctrl_init.
int hf_pcg_solve_async(hf_ctx* c, const SellOp& op, int step_slot, bool sum_parts) {
  PcgWork& w = c->ws;
  if (!op.p_spw) return hf_fail(HF_ERR_STATE, "persistent PCG kernel is not available for this mesh size");
  PersistArgs a;
  a.A = op.view();
  a.x = w.x.p;
  a.r = w.r.p;
  a.qpk = w.qpk.p;
  a.c = w.ctrl.p;
  a.slots = w.slots.p;
  a.acc = w.acc.p;
  a.acc_prev = w.acc_prev.p;
  a.gen = w.gen.p;
  a.cta_range = op.p_range.p;
  a.iters_out = (step_slot >= 0 && (size_t)step_slot < w.step_iters.n) ? w.step_iters.p + step_slot : nullptr;
  a.fail = w.fail.p;
  a.max_it = c->max_iters;
  a.nparts = sum_parts ? w.grid : 0;
  a.rtol = c->rtol;
  a.npad = c->Npad;
  a.mat_cap = op.p_mat_cap;
  a.sz_cap = op.p_sz_cap;
  void* args[] = {&a};
  const void* fn = (const void*)k_pcg_persist<HF_SPW>;
  // the attribute is per function, not per operator: other operators / contexts may have lowered it
  HF_CUDA(cudaFuncSetAttribute(k_pcg_persist<HF_SPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.p_smem));
  HF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(op.p_grid), dim3(HF_PT), args, op.p_smem, c->stream));
  c->stat_launches += 1;
  return HF_OK;
}
