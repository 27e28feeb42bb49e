// heatflow_b200 - persistent cooperative Jacobi-PCG on compact patches (sm_100a).
//
// One launch per solve; neighbouring CTAs exchange q = Ahat p on their halo rows as flag-with-data packets and
// there is one exact fixed-point grid reduction per iteration (hf_persist.cuh).  With the
// Hilbert node order a CTA's R = 256 x RPT rows form a compact 2-D patch whose halo is a short list
// (~4 sqrt(R) rows instead of two full mesh rows), so per row the CTA keeps
//   operator   8 B value + 2 B local column per stored entry       (~85 B, sliced-ELL padding included)
//   vectors    r, p on own + halo rows, q on halo rows              (~18 B)
// on chip: x and the first K entries of every row in registers (the 256 KB register file is otherwise idle),
// the remaining entries and the vectors in shared memory - up to R = 3584 rows per SM.
// The streaming kernel (hf_pcg.cu) needs ~17-21 us per iteration at these sizes (L2-resident, bound
// by launch and chunk latency); this kernel needs ~4-6 us.
// Only rows that appear in another CTA's halo list publish their q packet.
#include <algorithm>
#include <cmath>

#include "hf_ctx.cuh"
#include "hf_persist.cuh"

struct PatchArgs {
  int nslices;
  const int* slice_ptr;          // sliced-ELL offsets (SellOp)
  const double* val;             // sliced-ELL values (SellOp)
  const unsigned short* lcol;    // local columns for chunks of R rows: [0, R) own, R + h halo
  const int* halo_ptr;           // [G + 1]
  const int* halo_idx;           // global rows of the halo entries
  const unsigned char* pub;      // [Npad] 1 = the row is in some CTA's halo list
  double* x;                     // in: xhat_0, out: xhat
  const double* r;               // in: rhat_0
  uint4* qpk;                    // [2][Npad] q packets
  HfCtrl* c;
  unsigned long long* acc;       // fixed-point accumulators (shared with k_pcg_persist)
  unsigned long long* acc_prev;
  unsigned* gen;
  int* iters_out;
  int* fail;
  int max_it, npad, nparts;
  double rtol;
  int mat_cap, halo_cap;
  int eb_shift;                  // diagnostics (hf_debug_fx_shift): subtracted from the fixed-point exponent bounds
  long long* phase;              // diagnostics (-DHF_PHASE_TIMING): [G][2][8] clock64 cycles per phase, warps 0 and 1
};

#ifdef HF_PHASE_TIMING
#define HF_PT_DECL long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_t0 = clock64();
#define HF_PT_MARK(i)                                   \
  do {                                                  \
    const long long t1__ = clock64();                   \
    pt_acc[i] += t1__ - pt_t0;                          \
    pt_t0 = t1__;                                       \
  } while (0)
#define HF_PT_STORE                                                                         \
  if (P.phase && lane == 0 && warp < 2)                                                     \
    for (int i__ = 0; i__ < 8; ++i__) P.phase[((size_t)blockIdx.x * 2 + warp) * 8 + i__] = pt_acc[i__];
#else
#define HF_PT_DECL
#define HF_PT_MARK(i)
#define HF_PT_STORE
#endif

// MINB = CTAs per SM the kernel is compiled for.  1: the whole register file of the SM caches operator rows
// (fastest single solve).  2: half the registers and at most half the shared memory, so that the cooperative
// launches of two independent solves (two contexts, two streams: parameter sweeps) are co-resident and hide
// each other's reduction latency - 1.44 x the sweep throughput of back-to-back single solves (measured).
template <int RPT, int K, int MINB>
__global__ void __launch_bounds__(HF_PT, MINB) k_pcg_patch(PatchArgs P) {
  constexpr int R = HF_PT * RPT;
  constexpr int NSL = R / 32;                    // slices per chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);
  double* sp = sval + P.mat_cap;                 // p: own rows [0, R), halo [R, R + nh)
  double* sr = sp + R + P.halo_cap;              // r, same layout
  double* sqh = sr + R + P.halo_cap;             // validated halo q values
  double* red = sqh + P.halo_cap;                // reduction scratch: 2 x (HF_PW*3 + 3)
  int* shal = reinterpret_cast<int*>(red + 2 * (HF_PW * 3 + 3) + 2);   // halo rows (global indices)
  int* sbase = shal + P.halo_cap;                // [NSL + 1] offsets of the slices' shared-memory parts
  unsigned short* scol = reinterpret_cast<unsigned short*>(sbase + ((NSL + 4) & ~3));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, nsl = P.nslices;
  const int f = blockIdx.x * (R / 32);
  const int fend = min(f + R / 32, nsl);
  const int lo = blockIdx.x * R;
  const int hp = P.halo_ptr[blockIdx.x];
  const int nh = P.halo_ptr[blockIdx.x + 1] - hp;
  // shared-memory offsets of the slices: entries K.. of every row (the first K live in registers)
  if (warp == 0) {
    constexpr int PER = (NSL + 31) / 32;
    int cnt[PER], tot = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int sl = lane * PER + j, s = f + sl;
      cnt[j] = (sl < NSL && s < nsl) ? max(((P.slice_ptr[s + 1] - P.slice_ptr[s]) >> 5) - K, 0) * 32 : 0;
      tot += cnt[j];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - tot;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int sl = lane * PER + j;
      if (sl <= NSL) sbase[sl] = run;
      run += cnt[j];
    }
    if (lane == 31 && PER * 32 == NSL) sbase[NSL] = run;
  }
  for (int i = tid; i < R; i += HF_PT) {
    const double rv = (lo + i < P.npad) ? P.r[lo + i] : 0.0;
    sr[i] = rv;
    sp[i] = rv;                                  // p_0 = r_0
  }
  for (int h = tid; h < nh; h += HF_PT) {
    const int g = P.halo_idx[hp + h];
    shal[h] = g;
    const double rv = P.r[g];
    sr[R + h] = rv;
    sp[R + h] = rv;
  }
  __syncthreads();                               // sbase
  int base[RPT], wid[RPT];
  double x[RPT], q[RPT];
  double mv[RPT][K > 0 ? K : 1];
  int mc[RPT][K > 0 ? K : 1];
  unsigned pubmask = 0u;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int sl = warp * RPT + k, s = f + sl;
    base[k] = 0;
    wid[k] = -1;                                 // marks "no slice"
    x[k] = q[k] = 0.0;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
      mv[k][kk] = 0.0;
      mc[k][kk] = sl * 32 + lane;
    }
    if (s < nsl) {
      const int b0 = P.slice_ptr[s];
      const int w = (P.slice_ptr[s + 1] - b0) >> 5;
      wid[k] = w;
      base[k] = sbase[sl] + lane;
      x[k] = P.x[s * 32 + lane];
      if (P.pub[s * 32 + lane]) pubmask |= 1u << k;
#pragma unroll
      for (int kk = 0; kk < K; ++kk)
        if (kk < w) {
          mv[k][kk] = P.val[b0 + kk * 32 + lane];
          mc[k][kk] = P.lcol[b0 + kk * 32 + lane];
        }
      for (int kk = K; kk < w; ++kk) {           // the warp's own slices: coalesced
        sval[base[k] + (kk - K) * 32] = P.val[b0 + kk * 32 + lane];
        scol[base[k] + (kk - K) * 32] = P.lcol[b0 + kk * 32 + lane];
      }
    }
  }
  double thr, rr;                                // stopping threshold and ||r_0||^2
  if (P.nparts > 0) {                            // same sums, in the same order, as k_pcg_ctrl_init
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

  int it = 0;
  int since_check = 0;
  double rr_ref = rr;
  bool done = !(rr > thr);
  double pp = rr;                                // magnitude estimate of ||p||^2, see hf_persist.cu
  FxState fx;
  hf_fx_load_state(fx, P.acc_prev);
  while (!done && it < P.max_it) {
    ++gen;
    uint4* qout = P.qpk + (size_t)(it & 1) * P.npad;   // double buffered by iteration parity
    // ---- q = A p on the own rows (operator and p from shared memory); publish boundary rows; partial dots
    double d[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (wid[k] >= 0) {
        const int w = wid[k] - K, b = base[k];    // entries beyond the register-cached ones
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
          if (kk & 1) a1 = fma(mv[k][kk], sp[mc[k][kk]], a1);
          else a0 = fma(mv[k][kk], sp[mc[k][kk]], a0);
        }
        int kk = 0;
        for (; kk + 2 <= w; kk += 2) {
          const int c0 = scol[b + kk * 32], c1 = scol[b + (kk + 1) * 32];
          const double v0 = sval[b + kk * 32], v1 = sval[b + (kk + 1) * 32];
          a0 = fma(v0, sp[c0], a0);
          a1 = fma(v1, sp[c1], a1);
        }
        if (kk < w) a0 = fma(sval[b + kk * 32], sp[scol[b + kk * 32]], a0);
        const double acc = a0 + a1;
        q[k] = acc;
        const int i = (warp * RPT + k) * 32 + lane;
        if (pubmask & (1u << k)) hf_pkt_store(qout + lo + i, acc, gen);
        const double pv = sp[i], rv = sr[i];
        d[0] = fma(pv, acc, d[0]);
        d[1] = fma(rv, acc, d[1]);
        d[2] = fma(acc, acc, d[2]);
      }
    }
    const int e_pp = hf_exp2(pp), e_rr = hf_exp2(rr);
    const int eb3[3] = {hf_clamp_exp(e_pp + 4 + HF_FX_MARGIN - P.eb_shift), hf_clamp_exp((e_rr + e_pp + 1) / 2 + 5 + HF_FX_MARGIN),
                        hf_clamp_exp(e_pp + 8 + HF_FX_MARGIN)};
    hf_fx_arrive<3>(d, eb3, P.acc, gen, red, P.fail);
    // ---- halo q packets: warps >= 1 fetch them while warp 0 polls the reduction
    if (warp >= 1) {
      constexpr int NP = HF_PT - 32;
      for (int h0 = tid - 32; h0 < nh; h0 += 4 * NP) {
        uint4 hq[4];
        bool need[4];
        int g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          need[t] = (h0 + t * NP) < nh;
          g[t] = need[t] ? shal[h0 + t * NP] : 0;
        }
        bool pending;
        int spins = 0;
        do {
          pending = false;
          if (++spins > HF_SPIN_MAX) {                // a neighbour never published: give up loudly instead of hanging
            atomicAdd(P.fail, 1);
            break;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (need[t]) hq[t] = hf_pkt_load(qout + g[t]);
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (need[t]) {
              if (hf_pkt_ok(hq[t], gen)) {
                need[t] = false;
                sqh[h0 + t * NP] = hf_pkt_val(hq[t]);
              } else {
                pending = true;
              }
            }
        } while (pending);
      }
    }
    double tot[3];
    hf_fx_wait<3>(tot, eb3, P.acc, G, gen, red, fx, P.fail);
    const double alpha = rr / tot[0];
    double rr_new = fma(alpha * alpha, tot[2], fma(-2.0 * alpha, tot[1], rr));
    ++since_check;
    const bool check = (since_check >= HF_RR_CHECK) || !(rr_new > thr) || (rr_new < 1e-4 * rr_ref);
    const double beta = check ? 0.0 : rr_new / rr;
    // ---- own rows: x += alpha p ; r -= alpha q ; p = r + beta p (deferred when beta is not final)
    double dd[1] = {0.0};
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (wid[k] >= 0) {
        const int i = (warp * RPT + k) * 32 + lane;
        const double pv = sp[i];
        x[k] = fma(alpha, pv, x[k]);
        const double rv = fma(-alpha, q[k], sr[i]);
        sr[i] = rv;
        dd[0] = fma(rv, rv, dd[0]);
        if (!check) sp[i] = fma(beta, pv, rv);
      }
    }
    // ---- halo rows: the same update with the neighbours' q (bit-identical to the owner's)
    for (int h = tid; h < nh; h += HF_PT) {
      const double rv = fma(-alpha, sqh[h], sr[R + h]);
      sr[R + h] = rv;
      if (!check) sp[R + h] = fma(beta, sp[R + h], rv);
    }
    if (check) {
      ++gen;
      double t1[1];
      const int eb1[1] = {hf_clamp_exp(max(e_rr, 2 * hf_exp2(fabs(alpha)) + 8 + e_pp) + 2 + HF_FX_MARGIN)};
      hf_fx_arrive<1>(dd, eb1, P.acc, gen, red, P.fail);
      hf_fx_wait<1>(t1, eb1, P.acc, G, gen, red, fx, P.fail);
      rr_new = t1[0];
      rr_ref = rr_new;
      since_check = 0;
      const double b2 = rr_new / rr;
      for (int i = tid; i < R + nh; i += HF_PT) sp[i] = fma(b2, sp[i], sr[i]);
      done = !(rr_new > thr);
      pp = fma(b2 * b2, pp, fabs(rr_new));
    } else {
      pp = fma(beta * beta, pp, fabs(rr_new));
    }
    rr = rr_new;
    ++it;
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < RPT; ++k)
    if (wid[k] >= 0) P.x[lo + (warp * RPT + k) * 32 + lane] = x[k];
  hf_fx_store_state(fx, P.acc_prev);   // accumulator values the next launch starts from
  if (blockIdx.x == 0 && tid == 0) {
    P.c->rr = rr;
    P.c->itA = it;
    P.c->done = done ? 1 : 0;
    if (P.iters_out) *P.iters_out = it;
    if (!done || !isfinite(rr)) atomicAdd(P.fail, 1);
    *P.gen = gen;
  }
}

// ---------------------------------------------------------------------------------------
// Pipelined variant (Ghysels & Vanroose, "Hiding global synchronization latency in the preconditioned Conjugate
// Gradient algorithm", 2014, unpreconditioned form on the Jacobi-scaled operator): the ONE grid reduction of an
// iteration (gamma = r.r, delta = w.r with w = Ahat r) is started BEFORE the iteration's SpMV q = Ahat w and
// collected after it, so the ~1.4 us store->load trip through L2 overlaps the SpMV and the halo exchange
// instead of following them:
//     gamma_i = r.r ; delta_i = w.r                     -> arrive
//     q = Ahat w ; publish / fetch halo q               (while the reduction is in flight)
//     beta = gamma_i / gamma_{i-1} ; alpha = gamma_i / (delta_i - beta gamma_i / alpha_{i-1})      <- wait
//     z = q + beta z ; s = w + beta s ; p = r + beta p ; x += alpha p ; r -= alpha s ; w -= alpha z
// Same Krylov iterates as classic CG in exact arithmetic; the stopping test is on gamma, the directly summed
// ||r||^2 of the iterate x holds.  Measured (clock64 per phase, N = 1.4e5, 138 CTAs, cycles per iteration): arrive
// 1000, SpMV 940, halo fetch 1440 (hidden), wait for the sum 530 after the fetch, update 575: 2.42 us per iteration
// against 2.56 us for the classic kernel - the reduction trip (~2900 cycles from arrival to result) is now the
// critical path by itself; with 6 rows per thread the smaller register cache costs more than the overlap gains.  Measured against the LU oracle on the benchmark operators (numpy prototype,
// 3 sweep corner variants x 100 steps): same iteration counts (+-1 %) and the same 1e-12 .. 1.5e-11 agreement
// as classic CG - no residual replacement needed at rtol 1e-14.  Per row the CTA keeps x r p s z w in
// registers (own rows), w on own + halo rows and z on halo rows in shared memory.
template <int RPT, int K, int MINB>
__global__ void __launch_bounds__(HF_PT, MINB) k_pcg_pipe(PatchArgs P) {
  constexpr int R = HF_PT * RPT;
  constexpr int NSL = R / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sval = reinterpret_cast<double*>(smem_raw);
  double* sw = sval + P.mat_cap;                 // w: own rows [0, R), halo [R, R + nh)
  double* szh = sw + R + P.halo_cap;             // z on the halo rows
  double* sqh = szh + P.halo_cap;                // validated halo q values
  double* red = sqh + P.halo_cap;                // reduction scratch: 2 x (HF_PW*3 + 3)
  int* shal = reinterpret_cast<int*>(red + 2 * (HF_PW * 3 + 3) + 2);
  int* sbase = shal + P.halo_cap;
  unsigned short* scol = reinterpret_cast<unsigned short*>(sbase + ((NSL + 4) & ~3));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, nsl = P.nslices;
  const int f = blockIdx.x * (R / 32);
  const int lo = blockIdx.x * R;
  const int hp = P.halo_ptr[blockIdx.x];
  const int nh = P.halo_ptr[blockIdx.x + 1] - hp;
  if (warp == 0) {
    constexpr int PER = (NSL + 31) / 32;
    int cnt[PER], tot = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int sl = lane * PER + j, s = f + sl;
      cnt[j] = (sl < NSL && s < nsl) ? max(((P.slice_ptr[s + 1] - P.slice_ptr[s]) >> 5) - K, 0) * 32 : 0;
      tot += cnt[j];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - tot;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int sl = lane * PER + j;
      if (sl <= NSL) sbase[sl] = run;
      run += cnt[j];
    }
    if (lane == 31 && PER * 32 == NSL) sbase[NSL] = run;
  }
  // r_0 on own + halo rows goes through the w array first: w_0 = Ahat r_0
  for (int i = tid; i < R; i += HF_PT) sw[i] = (lo + i < P.npad) ? P.r[lo + i] : 0.0;
  for (int h = tid; h < nh; h += HF_PT) {
    const int g = P.halo_idx[hp + h];
    shal[h] = g;
    sw[R + h] = P.r[g];
    szh[h] = 0.0;
  }
  __syncthreads();
  int base[RPT], wid[RPT];
  double x[RPT], r[RPT], p[RPT], s[RPT], z[RPT], w[RPT];
  double mv[RPT][K > 0 ? K : 1];
  int mc[RPT][K > 0 ? K : 1];
  unsigned pubmask = 0u;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int sl = warp * RPT + k, sg = f + sl;
    base[k] = 0;
    wid[k] = -1;
    x[k] = r[k] = p[k] = s[k] = z[k] = w[k] = 0.0;
#pragma unroll
    for (int kk = 0; kk < K; ++kk) {
      mv[k][kk] = 0.0;
      mc[k][kk] = sl * 32 + lane;
    }
    if (sg < nsl) {
      const int b0 = P.slice_ptr[sg];
      const int wd = (P.slice_ptr[sg + 1] - b0) >> 5;
      wid[k] = wd;
      base[k] = sbase[sl] + lane;
      x[k] = P.x[sg * 32 + lane];
      r[k] = sw[sl * 32 + lane];
      if (P.pub[sg * 32 + lane]) pubmask |= 1u << k;
#pragma unroll
      for (int kk = 0; kk < K; ++kk)
        if (kk < wd) {
          mv[k][kk] = P.val[b0 + kk * 32 + lane];
          mc[k][kk] = P.lcol[b0 + kk * 32 + lane];
        }
      for (int kk = K; kk < wd; ++kk) {
        sval[base[k] + (kk - K) * 32] = P.val[b0 + kk * 32 + lane];
        scol[base[k] + (kk - K) * 32] = P.lcol[b0 + kk * 32 + lane];
      }
    }
  }
  double thr, rr;
  if (P.nparts > 0) {
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
  // one SpMV from shared memory: out[k] = sum_j Ahat_kj v_j with v on own + halo rows in sw
  // pub_buf >= 0: the result of a boundary row is published (packet buffer pub_buf, tag pub_tag) as soon as it is
  // computed - hf_set_mesh puts the boundary rows into the first slice of every warp, so the packets leave after a
  // quarter of the SpMV
  auto spmv = [&](double (&out)[RPT], const int pub_buf, const unsigned pub_tag) {
    uint4* qpub = P.qpk + (size_t)(pub_buf < 0 ? 0 : pub_buf) * P.npad;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      out[k] = 0.0;
      if (wid[k] >= 0) {
        const int wd = wid[k] - K, b = base[k];
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
          if (kk & 1) a1 = fma(mv[k][kk], sw[mc[k][kk]], a1);
          else a0 = fma(mv[k][kk], sw[mc[k][kk]], a0);
        }
        int kk = 0;
        for (; kk + 2 <= wd; kk += 2) {
          const int c0 = scol[b + kk * 32], c1 = scol[b + (kk + 1) * 32];
          const double v0 = sval[b + kk * 32], v1 = sval[b + (kk + 1) * 32];
          a0 = fma(v0, sw[c0], a0);
          a1 = fma(v1, sw[c1], a1);
        }
        if (kk < wd) a0 = fma(sval[b + kk * 32], sw[scol[b + kk * 32]], a0);
        out[k] = a0 + a1;
        if (pub_buf >= 0 && (pubmask & (1u << k))) hf_pkt_store_nb(qpub + lo + (warp * RPT + k) * 32 + lane, out[k], pub_tag);
      }
    }
  };
  // fetch the halo rows of the vector the neighbours published as packets tagged `tag` in buffer `buf` into dst[0 .. nh)
  auto exchange = [&](int buf, unsigned tag, double* dst, bool skip_warp0) {
    uint4* qout = P.qpk + (size_t)buf * P.npad;
    if (skip_warp0 && warp == 0) return;
    const int first = skip_warp0 ? 32 : 0;
    const int NP = HF_PT - first;
    for (int h0 = tid - first; h0 < nh; h0 += 4 * NP) {
      uint4 hq[4];
      bool need[4];
      int g[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        need[t] = (h0 + t * NP) < nh;
        g[t] = need[t] ? shal[h0 + t * NP] : 0;
      }
      bool pending;
      int spins = 0;
      do {
        pending = false;
        if (++spins > HF_SPIN_MAX) {                  // a neighbour never published: give up loudly instead of hanging
          atomicAdd(P.fail, 1);
          break;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (need[t]) hq[t] = hf_pkt_load(qout + g[t]);
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (need[t]) {
            if (hf_pkt_ok(hq[t], tag)) {
              need[t] = false;
              dst[h0 + t * NP] = hf_pkt_val(hq[t]);
            } else {
              pending = true;
            }
          }
      } while (pending);
    }
  };

  int it = 0;
  bool done = !(rr > thr);
  double inv_gam_old = 1.0, inv_alpha = 1.0, alpha = 1.0, gam_est = rr;
  FxState fx;
  hf_fx_load_state(fx, P.acc_prev);
  if (!done && P.max_it > 0) {
    // ---- w_0 = Ahat r_0 on the own rows, then on the halo rows (packets in buffer 1, tag gen + 1)
    spmv(w, 1, gen + 1u);
    exchange(1, gen + 1u, sqh, false);
    __syncthreads();                               // every thread has read r_0 from sw
#pragma unroll
    for (int k = 0; k < RPT; ++k)
      if (wid[k] >= 0) sw[(warp * RPT + k) * 32 + lane] = w[k];
    for (int h = tid; h < nh; h += HF_PT) sw[R + h] = sqh[h];
    __syncthreads();
  }
  HF_PT_DECL
  while (!done && it < P.max_it) {
    ++gen;
    HF_PT_MARK(7);
    // ---- gamma = r.r, delta = w.r on the own rows: the reduction starts before the SpMV
    double d[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      d[0] = fma(r[k], r[k], d[0]);
      d[1] = fma(w[k], r[k], d[1]);
    }
    const int e_g = hf_exp2(gam_est);
    const int eb2[2] = {hf_clamp_exp(e_g + HF_FX_MARGIN - P.eb_shift), hf_clamp_exp(e_g + 4 + HF_FX_MARGIN)};
    hf_fx_arrive_warp<2>(d, eb2, P.acc, gen, P.fail);
    HF_PT_MARK(0);
    // ---- q = Ahat w while the reduction is in flight; halo q packets
    double q[RPT];
    spmv(q, it & 1, gen);
    HF_PT_MARK(1);
    exchange(it & 1, gen, sqh, true);
    HF_PT_MARK(2);
    double tot[2];
    hf_fx_wait<2, 0, HF_PW>(tot, eb2, P.acc, G, gen, red, fx, P.fail);   // no poll delay: the SpMV has already covered the trip
    HF_PT_MARK(3);
    const double gam = tot[0], delta = tot[1];
    rr = gam;
    if (!(gam > thr)) {                            // also stops on NaN; x is the iterate gamma belongs to
      done = true;
      break;
    }
    // beta = gamma_i / gamma_{i-1}, alpha = gamma_i / (delta_i - beta gamma_i / alpha_{i-1}) with the reciprocals of
    // gamma_{i-1} and alpha_{i-1} carried over: one division on the critical path (two independent ones per iteration)
    const double beta = (it > 0) ? gam * inv_gam_old : 0.0;
    const double den = (it > 0) ? fma(-beta * gam, inv_alpha, delta) : delta;
    alpha = gam / den;
    inv_alpha = den / gam;
    inv_gam_old = 1.0 / gam;
    // ---- own rows: z = q + beta z ; s = w + beta s ; p = r + beta p ; x += alpha p ; r -= alpha s ; w -= alpha z
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      if (wid[k] >= 0) {
        z[k] = fma(beta, z[k], q[k]);
        s[k] = fma(beta, s[k], w[k]);
        p[k] = fma(beta, p[k], r[k]);
        x[k] = fma(alpha, p[k], x[k]);
        r[k] = fma(-alpha, s[k], r[k]);
        w[k] = fma(-alpha, z[k], w[k]);
        sw[(warp * RPT + k) * 32 + lane] = w[k];
      }
    }
    // ---- halo rows: the same z and w updates with the neighbours' q (bit-identical to the owner's)
    for (int h = tid; h < nh; h += HF_PT) {
      const double zh = fma(beta, szh[h], sqh[h]);
      szh[h] = zh;
      sw[R + h] = fma(-alpha, zh, sw[R + h]);
    }
    gam_est = gam;
    ++it;
    HF_PT_MARK(4);
    __syncthreads();
    HF_PT_MARK(5);
  }
  HF_PT_STORE
#pragma unroll
  for (int k = 0; k < RPT; ++k)
    if (wid[k] >= 0) P.x[lo + (warp * RPT + k) * 32 + lane] = x[k];
  hf_fx_store_state(fx, P.acc_prev);
  if (blockIdx.x == 0 && tid == 0) {
    P.c->rr = rr;
    P.c->itA = it;
    P.c->done = done ? 1 : 0;
    if (P.iters_out) *P.iters_out = it;
    if (!done || !isfinite(rr)) atomicAdd(P.fail, 1);
    *P.gen = gen;
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// SELL-ordered 16-bit local columns for chunks of R rows from the CSR-ordered ones
__global__ void k_patch_lcol(int N, int Npad, int R, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr,
                             const unsigned short* __restrict__ lcol_csr, unsigned short* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const int s = i / HF_SLICE, lane = i % HF_SLICE;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) / HF_SLICE;
  int len = 0, r0 = 0;
  if (i < N) {
    r0 = rowptr[i];
    len = rowptr[i + 1] - r0;
  }
  for (int k = 0; k < w; ++k) out[base + k * HF_SLICE + lane] = (k < len) ? lcol_csr[r0 + k] : (unsigned short)(i % R);
}

static size_t patch_smem_bytes(int R, int mat_cap, int halo_cap) {
  return sizeof(double) * ((size_t)mat_cap + 2 * ((size_t)R + halo_cap) + halo_cap + 2 * (HF_PW * 3 + 3) + 2) +
         sizeof(int) * ((size_t)halo_cap + ((R / 32 + 4) & ~3)) + sizeof(unsigned short) * (size_t)mat_cap;
}

// (rows per thread, register-cached entries per row): the cache is sized to the register budget of one
// CTA of 256 threads per SM (255 registers per thread)
static const int kPatchRpt[6] = {4, 6, 8, 10, 12, 14};
static const int kPatchK[2][6] = {{8, 6, 4, 4, 3, 2}, {2, 1, 0, 0, 0, 0}};   // [share - 1][rows-per-thread index]
// the pipelined variant keeps six vectors of the own rows in registers, so it caches fewer operator entries (-1: no such
// kernel); the shared-memory operator part is sized for the smaller of the two caches
static const int kPipeK[2][6] = {{8, 4, 2, -1, -1, -1}, {-1, -1, -1, -1, -1, -1}};

static const void* patch_kernel(int rpt, int share) {
  if (share == 2) {
    switch (rpt) {
      case 4: return (const void*)k_pcg_patch<4, 2, 2>;
      case 6: return (const void*)k_pcg_patch<6, 1, 2>;
      case 8: return (const void*)k_pcg_patch<8, 0, 2>;
      case 10: return (const void*)k_pcg_patch<10, 0, 2>;
      case 12: return (const void*)k_pcg_patch<12, 0, 2>;
      default: return (const void*)k_pcg_patch<14, 0, 2>;
    }
  }
  switch (rpt) {
    case 4: return (const void*)k_pcg_patch<4, 8, 1>;
    case 6: return (const void*)k_pcg_patch<6, 6, 1>;
    case 8: return (const void*)k_pcg_patch<8, 4, 1>;
    case 10: return (const void*)k_pcg_patch<10, 4, 1>;
    case 12: return (const void*)k_pcg_patch<12, 3, 1>;
    default: return (const void*)k_pcg_patch<14, 2, 1>;
  }
}
// pipelined variant: x r p s z w of the own rows live in registers, so it exists for the small rows-per-thread counts
static const void* pipe_kernel(int rpt, int share) {
  // two co-resident solves (hf_set_sharing(2)) already hide each other's reduction trip, and with half the register
  // file the six own-row vectors of the pipelined kernel leave room for one cached operator entry only: measured
  // 28.7 simulations/s with k_pcg_pipe<4, 1, 2> against 32.4 with the classic k_pcg_patch<4, 2, 2> (two host threads)
  if (share == 2) return nullptr;
  switch (rpt) {
    case 4: return (const void*)k_pcg_pipe<4, 8, 1>;
    case 6: return (const void*)k_pcg_pipe<6, 4, 1>;       // meshes of 1.5e5 - 2.3e5 dofs (the reference's own gmsh meshes start there)
    case 8: return (const void*)k_pcg_pipe<8, 2, 1>;       // up to 3.0e5 dofs
    default: return nullptr;
  }
}
bool hf_patch_pipelined(const hf_ctx* c, const SellOp& vals) {
  const SellOp& op = vals.plan_from ? *vals.plan_from : vals;
  const int mode = c->force_mode >= 0 ? c->force_mode : c->mode;
  return op.pp_rpt != 0 && mode != 4 && c->op_transient && pipe_kernel(op.pp_rpt, op.pp_share) != nullptr;
}
static int patch_set_smem_rpt(int rpt, int share, size_t bytes) {
  HF_CUDA(cudaFuncSetAttribute(patch_kernel(rpt, share), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (const void* fn = pipe_kernel(rpt, share)) HF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return HF_OK;
}

// Decide rows per thread / grid / shared-memory layout; pp_rpt = 0 when the mesh does not fit.
int hf_patch_plan(hf_ctx* c, SellOp& op) {
  op.pp_rpt = 0;
  const int nsl = op.nslices;
  if (nsl == 0 || c->h_rowptr.empty()) return HF_OK;
  int coop = 0, max_smem = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  if (!coop) return HF_OK;
  std::vector<int> sp(nsl + 1);
  HF_CUDA(cudaMemcpyAsync(sp.data(), op.slice_ptr.p, sizeof(int) * (nsl + 1), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  const int share = c->share;
  // two co-resident CTAs per SM: each gets half of the SM's shared memory (1 KB per CTA is reserved by the system)
  if (share == 2) {
    int per_sm = 0;
    cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c->device);
    max_smem = std::min(max_smem, per_sm / 2 - 1024);
  }
  for (int t = 0; t < 6; ++t) {
    const int rpt = kPatchRpt[t], R = HF_PT * rpt;
    const int K = kPipeK[share - 1][t] >= 0 ? std::min(kPatchK[share - 1][t], kPipeK[share - 1][t]) : kPatchK[share - 1][t];
    const int G = (c->Npad + R - 1) / R;
    if (G > c->sm_count || G > HF_MAX_GRID) continue;
    int mat_cap = 0;                              // shared-memory part of a chunk's operator: entries K.. of every row
    for (int b = 0; b < G; ++b) {
      const int f = std::min(b * (R / 32), nsl), fe = std::min(f + R / 32, nsl);
      int n = 0;
      for (int sl = f; sl < fe; ++sl) n += std::max(((sp[sl + 1] - sp[sl]) >> 5) - K, 0) * 32;
      mat_cap = std::max(mat_cap, n);
    }
    std::vector<int> hptr, hidx;
    std::vector<unsigned short> lcol;
    int halo_max = 0;
    if (hf_build_patches(c, R, hptr, hidx, lcol, &halo_max) != HF_OK) continue;
    mat_cap = (mat_cap + 7) & ~7;
    const int halo_cap = (halo_max + 3) & ~3;
    const size_t bytes = patch_smem_bytes(R, mat_cap, halo_cap);
    if (bytes > (size_t)max_smem) continue;
    // rows that some other chunk reads publish their q packets
    std::vector<unsigned char> pub(c->Npad, 0);
    for (int g : hidx) pub[g] = 1;
    if (hidx.empty()) hidx.push_back(0);
    hptr.resize(G + 1, hptr.empty() ? 0 : hptr.back());
    DevBuf<unsigned short> lcol_csr;
    HF_TRY(lcol_csr.upload(lcol.data(), lcol.size(), c->stream));
    HF_TRY(op.pp_lcol.alloc(op.padded_nnz, c->stream));
    k_patch_lcol<<<(c->Npad + 255) / 256, 256, 0, c->stream>>>(c->N, c->Npad, R, c->rowptr.p, op.slice_ptr.p, lcol_csr.p, op.pp_lcol.p);
    HF_CUDA(cudaGetLastError());
    HF_TRY(op.pp_halo_ptr.upload(hptr.data(), hptr.size(), c->stream));
    HF_TRY(op.pp_halo_idx.upload(hidx.data(), hidx.size(), c->stream));
    HF_TRY(op.pp_pub.upload(pub.data(), pub.size(), c->stream));
    op.pp_rpt = rpt;
    op.pp_grid = G;
    op.pp_mat_cap = mat_cap;
    op.pp_halo_cap = halo_cap;
    op.pp_smem = bytes;
    op.pp_share = share;
    HF_TRY(patch_set_smem_rpt(rpt, share, bytes));
    PcgWork& w = c->ws;
    if (w.acc.n == 0) {
      HF_TRY(w.gen.alloc(1, c->stream));
      HF_TRY(w.fail.alloc(1, c->stream));
      HF_TRY(w.acc.alloc((size_t)2 * HF_NREP * HF_ACC_LINE, c->stream));
      HF_TRY(w.acc_prev.alloc((size_t)2 * HF_NREP * HF_ACC_LINE, c->stream));
    }
    if (w.qpk.n < (size_t)2 * c->Npad) HF_TRY(w.qpk.alloc((size_t)2 * c->Npad, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));           // lcol_csr goes out of scope
    break;
  }
  return HF_OK;
}

int hf_patch_solve_async(hf_ctx* c, const SellOp& vals, int step_slot, bool sum_parts) {
  PcgWork& w = c->ws;
  const SellOp& op = vals.plan_from ? *vals.plan_from : vals;     // plan (patches) may be borrowed; values are the operator's own
  if (!op.pp_rpt) return hf_fail(HF_ERR_STATE, "patch PCG kernel is not available for this mesh size");
  PatchArgs a;
  a.nslices = vals.nslices;
  a.slice_ptr = vals.slice_ptr.p;
  a.val = vals.val.p;
  a.lcol = op.pp_lcol.p;
  a.halo_ptr = op.pp_halo_ptr.p;
  a.halo_idx = op.pp_halo_idx.p;
  a.pub = op.pp_pub.p;
  a.x = w.x.p;
  a.r = w.r.p;
  a.qpk = w.qpk.p;
  a.c = w.ctrl.p;
  a.acc = w.acc.p;
  a.acc_prev = w.acc_prev.p;
  a.gen = w.gen.p;
  a.iters_out = (step_slot >= 0 && (size_t)step_slot < w.step_iters.n) ? w.step_iters.p + step_slot : nullptr;
  a.fail = w.fail.p;
  a.max_it = c->max_iters;
  a.npad = c->Npad;
  a.nparts = sum_parts ? w.grid : 0;
  a.rtol = c->rtol;
  a.mat_cap = op.pp_mat_cap;
  a.halo_cap = op.pp_halo_cap;
  a.eb_shift = c->debug_fx_shift;
  a.phase = c->debug_phase.n ? c->debug_phase.p : nullptr;
  void* args[] = {&a};
  const void* fn = hf_patch_pipelined(c, vals) ? pipe_kernel(op.pp_rpt, op.pp_share) : patch_kernel(op.pp_rpt, op.pp_share);
  HF_TRY(patch_set_smem_rpt(op.pp_rpt, op.pp_share, op.pp_smem));   // per function, not per operator
  HF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(op.pp_grid), dim3(HF_PT), args, op.pp_smem, c->stream));
  c->stat_launches += 1;
  return HF_OK;
}
