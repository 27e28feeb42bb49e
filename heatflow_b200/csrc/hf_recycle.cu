// heatflow_b200 - initial guess of the Jacobi-PCG solve from the previous time steps (sm_100a).
//
// The reference solves every time step with one LU factorisation (run_with_diamond.py:389-394,
// :480).  An iterative solver pays for every step from scratch unless it reuses what the earlier
// solves found: all steps share the operator  Ahat = D^-1/2 (M + dt K) D^-1/2,  and the solutions
// of a diffusion problem driven by one scalar amplitude stay close to a low-dimensional space.
// This file keeps an Ahat-orthogonal basis  W = [w_1 .. w_m]  of the corrections the first m solves
// computed, together with  inv_k = 1 / (w_k . Ahat w_k):
//   before the solve   c_k = inv_k (w_k . r0),  x0 <- x0 + W c,  r0 <- r0 - Ahat (W c)
//                      (Galerkin projection: the error of x0 becomes Ahat-orthogonal to span W; the new residual comes
//                      from ONE SpMV with the increment W c, so it is the true residual of the new x0 to rounding);
//   after the solve    d = x - x0,  Ad = Ahat d (one SpMV),  h_k = inv_k (w_k . Ad),  w_{m+1} = d - W h
//                      (one Gram-Schmidt pass in the Ahat inner product; without it the computed corrections are only
//                      orthogonal to ~1e-4 and the projection stops paying).
// Ahat W is never stored: w_k . Ahat d = w_k . Ad needs W and the fresh Ad only, and the residual update needs
// Ahat (W c), one SpMV.  The Gram-Schmidt pass is lazy and fused with the next projection: ONE pass over W computes
// both w_k . r0 (new step) and w_k . Ad (pending correction), a second pass applies h (finalising w_{m+1}) and c, so a
// step streams 2 m N doubles (round 1: 4 m N with a stored Ahat W, 6 m N before the lazy update) plus two SpMVs.  The
// new vector's coefficient follows from the raw dot products:
//   w.r0 = d.r0 - sum h_k (w_k.r0),   w.Ahat w = d.Ad - sum h_k^2 / inv_k.
// The solver itself is unchanged (Jacobi-PCG to the same tolerance on the same system): only the
// starting point moves, so the converged answer is the same to the solver tolerance.  Measured on
// geballe_with_diamond (N = 1.4e5, 100 steps): 129 instead of 620 iterations per step.
// All kernels are stream ordered; reductions add per-warp / per-CTA partials in a fixed order.
#include <algorithm>
#include <cmath>

#include "hf_ctx.cuh"

#ifndef RC_SEG
#define RC_SEG 512            // rows per CTA of the dot-product kernel (two CTAs per SM at 1.4e5 dofs: their load /
#endif                        // reduce phases overlap; 1024 rows per CTA was 3 % slower end to end)
#define RC_NV (RC_SEG / 64)    // 16-byte loads per lane and basis vector
#define RC_WARPS 8

// parts[k * nseg + seg] = sum over the RC_SEG rows of segment `seg` of V[k][i] * v[i], k < m (vector m_extra is read
// from `extra` instead of its slot), and, with TWO, parts2[k * nseg + seg] = the same with u instead of v.
// One CTA per segment; every lane keeps its RC_SEG / 32 values of v (and u) in registers (consecutive-pair loads) and
// the warps share the basis vectors (warp w takes k = w, w + 8, ...), so each warp has RC_SEG / 64 independent 512-byte
// loads in flight per basis vector and no barrier is needed.
template <bool TWO>
__global__ void __launch_bounds__(RC_WARPS * 32, 2)
k_rc_dots(int m, int nseg, size_t ld, const double* __restrict__ V, const double* __restrict__ v, const double* __restrict__ u,
          double* __restrict__ parts, double* __restrict__ parts2, int m_extra, const double* __restrict__ extra) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x;
  const size_t base = (size_t)seg * RC_SEG + 2 * lane;
  double2 vv[RC_NV], uu[TWO ? RC_NV : 1];
#pragma unroll
  for (int j = 0; j < RC_NV; ++j) {
    vv[j] = *reinterpret_cast<const double2*>(v + base + 64 * j);
    if (TWO) uu[j] = *reinterpret_cast<const double2*>(u + base + 64 * j);
  }
  for (int k = warp; k < m; k += RC_WARPS) {
    const double* p = ((k == m_extra) ? extra : V + (size_t)k * ld) + base;   // the pending raw correction is not in a slot yet
    double2 a[RC_NV];
#pragma unroll
    for (int j = 0; j < RC_NV; ++j) a[j] = __ldcs(reinterpret_cast<const double2*>(p + 64 * j));
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < RC_NV; ++j) {
      s0 = fma(a[j].x, vv[j].x, s0);
      s1 = fma(a[j].y, vv[j].y, s1);
      if (TWO) {
        t0 = fma(a[j].x, uu[j].x, t0);
        t1 = fma(a[j].y, uu[j].y, t1);
      }
    }
    double sv = s0 + s1, tv = t0 + t1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sv += __shfl_xor_sync(0xffffffffu, sv, o);
      if (TWO) tv += __shfl_xor_sync(0xffffffffu, tv, o);
    }
    if (lane == 0) {
      parts[(size_t)k * nseg + seg] = sv;
      if (TWO) parts2[(size_t)k * nseg + seg] = tv;
    }
  }
}

// Projection coefficients, one CTA.  t_k = sum_seg parts[k][seg] = w_k . r0 (one warp per k, fixed order);
// c_k = inv_k t_k for the m finalised vectors.  With a pending raw correction (index m; parts2[k] = w_k . Ad,
// d.Ad partials in part_nn):  hn_k = -inv_k (w_k . Ad),
//   w.r0 = t_m + sum hn_k t_k,  w.Ahat w = d.Ad - sum hn_k^2 / inv_k,  inv_m = 1 / (w.Ahat w),  c_m = inv_m (w.r0)
// (inv_m = 0 when the correction is empty or not finite: the slot stays inert).
#define RC_CT 1024
__global__ void __launch_bounds__(RC_CT)
k_rc_coef_project(int m, int pending, int nseg, const double* __restrict__ parts, const double* __restrict__ parts2,
                  double* __restrict__ inv, double* __restrict__ hn, int nparts_nn, const double* __restrict__ part_nn,
                  double* __restrict__ coef) {
  __shared__ double s_t[RC_CT];        // m + pending <= cap + 1 <= RC_CT is checked on the host
  __shared__ double s_h[RC_CT];
  __shared__ double s_red[3][RC_CT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mm = m + pending;
  for (int k = warp; k < mm; k += RC_CT / 32) {
    double s = 0.0, g = 0.0;
    for (int i = lane; i < nseg; i += 32) {
      s += __ldcg(parts + (size_t)k * nseg + i);
      if (pending && k < m) g += __ldcg(parts2 + (size_t)k * nseg + i);
    }
    s = hf_warp_sum(s);
    g = hf_warp_sum(g);
    if (lane == 0) {
      s_t[k] = s;
      s_h[k] = (pending && k < m) ? -inv[k] * g : 0.0;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < m; k += RC_CT) {
    coef[k] = inv[k] * s_t[k];
    if (pending) hn[k] = s_h[k];
  }
  if (!pending) return;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int k = threadIdx.x; k < m; k += RC_CT) {
    const double h = s_h[k], iv = inv[k];
    s1 = fma(h, s_t[k], s1);
    if (iv != 0.0) s2 = fma(h, h / iv, s2);
  }
  for (int i = threadIdx.x; i < nparts_nn; i += RC_CT) s3 += __ldcg(part_nn + i);
  s1 = hf_warp_sum(s1);
  s2 = hf_warp_sum(s2);
  s3 = hf_warp_sum(s3);
  if (lane == 0) {
    s_red[0][warp] = s1;
    s_red[1][warp] = s2;
    s_red[2][warp] = s3;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int w = 0; w < RC_CT / 32; ++w) {
      a1 += s_red[0][w];
      a2 += s_red[1][w];
      a3 += s_red[2][w];
    }
    const double nn = a3 - a2;
    const double iv = (nn > 0.0 && isfinite(nn) && isfinite(1.0 / nn)) ? 1.0 / nn : 0.0;
    inv[m] = iv;
    coef[m] = iv * (s_t[m] + a1);
  }
}

// Fused pass over W before the solve:  a = x0.
//   PENDING: finalise the raw correction d of the previous solve as vector m:  w = d + W hn  -> slot m of W
//   inc = W c (+ c_m w) ;  a += inc ;  x0save = a (null once the basis is frozen)
template <bool PENDING>
__global__ void __launch_bounds__(HF_BLOCK)
k_rc_apply(int m, int n, size_t ld, double* __restrict__ W, const double* __restrict__ coef, const double* __restrict__ hn,
           const double* __restrict__ d, double* __restrict__ a, double* __restrict__ inc, double* __restrict__ x0save) {
  extern __shared__ double s_c[];      // c[0..m] then hn[0..m)
  double* s_h = s_c + m + 1;
  for (int k = threadIdx.x; k <= m; k += HF_BLOCK) s_c[k] = (k < m || PENDING) ? coef[k] : 0.0;
  if (PENDING)
    for (int k = threadIdx.x; k < m; k += HF_BLOCK) s_h[k] = hn[k];
  __syncthreads();
  // two rows per thread (16-byte loads; n and ld are even), four basis vectors per pass: 8 independent loads in flight
  // per thread.  Per row the sums run over k in the same order as a one-row loop would.
  for (int i = 2 * (blockIdx.x * HF_BLOCK + threadIdx.x); i < n; i += 2 * gridDim.x * HF_BLOCK) {
    double2 ca0 = {0.0, 0.0}, ca1 = {0.0, 0.0}, ga0 = {0.0, 0.0}, ga1 = {0.0, 0.0};
    const double* w = W + i;
    auto ld2 = [](const double* q) {
      double2 v;
      asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(q));
      return v;
    };
    auto acc = [](double c, const double2& v, double2& s) {
      s.x = fma(c, v.x, s.x);
      s.y = fma(c, v.y, s.y);
    };
    int k = 0;
    for (; k + 4 <= m; k += 4) {
      const double2 w0 = ld2(w + (size_t)k * ld), w1 = ld2(w + (size_t)(k + 1) * ld), w2 = ld2(w + (size_t)(k + 2) * ld),
                    w3 = ld2(w + (size_t)(k + 3) * ld);
      acc(s_c[k], w0, ca0);
      acc(s_c[k + 1], w1, ca1);
      acc(s_c[k + 2], w2, ca0);
      acc(s_c[k + 3], w3, ca1);
      if (PENDING) {
        acc(s_h[k], w0, ga0);
        acc(s_h[k + 1], w1, ga1);
        acc(s_h[k + 2], w2, ga0);
        acc(s_h[k + 3], w3, ga1);
      }
    }
    for (; k < m; ++k) {
      const double2 wk = ld2(w + (size_t)k * ld);
      acc(s_c[k], wk, ca0);
      if (PENDING) acc(s_h[k], wk, ga0);
    }
    double2 ca = {ca0.x + ca1.x, ca0.y + ca1.y};
    if (PENDING) {
      const double2 dv = *reinterpret_cast<const double2*>(d + i);
      const double2 wm = {dv.x + (ga0.x + ga1.x), dv.y + (ga0.y + ga1.y)};
      *reinterpret_cast<double2*>(W + (size_t)m * ld + i) = wm;
      acc(s_c[m], wm, ca);
    }
    const double2 av = *reinterpret_cast<const double2*>(a + i);
    const double2 x = {av.x + ca.x, av.y + ca.y};
    *reinterpret_cast<double2*>(a + i) = x;
    *reinterpret_cast<double2*>(inc + i) = ca;
    if (x0save) *reinterpret_cast<double2*>(x0save + i) = x;
  }
}

// r -= Ahat inc and the partial sums of r.r -> part[blockIdx.x] (sliced-ELL, one warp per slice, grid-stride over the
// slices so that the solver kernels find exactly gridDim.x partial sums; the gathers hit L2)
__global__ void __launch_bounds__(HF_BLOCK)
k_rc_resid(SellView A, const double* __restrict__ inc, double* __restrict__ r, double* __restrict__ part) {
  __shared__ double sh[HF_BLOCK / 32];
  const int lane = threadIdx.x & 31;
  double local = 0.0;
  for (int s = blockIdx.x * (HF_BLOCK / 32) + (threadIdx.x >> 5); s < A.nslices; s += gridDim.x * (HF_BLOCK / 32)) {
    const int b0 = A.slice_ptr[s];
    const int w = (A.slice_ptr[s + 1] - b0) >> 5;
    const int* cp = A.col + b0 + lane;
    const double* vp = A.val + b0 + lane;
    double acc0 = 0.0, acc1 = 0.0;
    int k = 0;
    for (; k + 2 <= w; k += 2) {
      const int c0 = cp[k * 32], c1 = cp[(k + 1) * 32];
      acc0 = fma(vp[k * 32], inc[c0], acc0);
      acc1 = fma(vp[(k + 1) * 32], inc[c1], acc1);
    }
    if (k < w) acc0 = fma(vp[k * 32], inc[cp[k * 32]], acc0);
    const int i = s * HF_SLICE + lane;
    const double rv = r[i] - (acc0 + acc1);
    r[i] = rv;
    local = fma(rv, rv, local);
  }
  const double tot = hf_block_sum(local, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// d = x - x0, ad = Ahat d and the partial sums of d.ad in one pass (sliced-ELL, one warp per slice;
// the gathers hit L2)
__global__ void __launch_bounds__(HF_BLOCK)
k_rc_spmv(SellView A, const double* __restrict__ x, const double* __restrict__ x0, double* __restrict__ d,
          double* __restrict__ ad, double* __restrict__ part_nn) {
  __shared__ double sh[HF_BLOCK / 32];
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (HF_BLOCK / 32) + (threadIdx.x >> 5);
  double local = 0.0;
  if (s < A.nslices) {
    const int b0 = A.slice_ptr[s];
    const int w = (A.slice_ptr[s + 1] - b0) >> 5;
    const int* cp = A.col + b0 + lane;
    const double* vp = A.val + b0 + lane;
    double acc0 = 0.0, acc1 = 0.0;
    int k = 0;
    for (; k + 2 <= w; k += 2) {
      const int c0 = cp[k * 32], c1 = cp[(k + 1) * 32];
      const double v0 = vp[k * 32], v1 = vp[(k + 1) * 32];
      acc0 = fma(v0, x[c0] - x0[c0], acc0);
      acc1 = fma(v1, x[c1] - x0[c1], acc1);
    }
    if (k < w) {
      const int c0 = cp[k * 32];
      acc0 = fma(vp[k * 32], x[c0] - x0[c0], acc0);
    }
    const int i = s * HF_SLICE + lane;
    const double di = x[i] - x0[i], adi = acc0 + acc1;
    d[i] = di;
    ad[i] = adi;
    local = di * adi;
  }
  const double tot = hf_block_sum(local, sh);
  if (threadIdx.x == 0) part_nn[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
void hf_rc_reset(hf_ctx* c) {
  c->rc.count = 0;
  c->rc.pending = false;
}

static int rc_alloc(hf_ctx* c) {
  Recycle& rc = c->rc;
  const size_t ld = ((size_t)c->Npad + RC_SEG - 1) / RC_SEG * RC_SEG;
  if (rc.ld == ld && rc.W.n == (size_t)rc.cap * ld) return HF_OK;
  rc.ld = ld;
  rc.nseg = (int)(ld / RC_SEG);
  rc.nn_parts = (c->Npad / HF_SLICE + HF_BLOCK / 32 - 1) / (HF_BLOCK / 32);
  HF_TRY(rc.W.alloc((size_t)rc.cap * ld, c->stream));
  HF_TRY(rc.inv.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.coef.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.hn.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.parts.alloc((size_t)2 * (rc.cap + 1) * rc.nseg, c->stream));   // w_k . r0 | w_k . Ad
  HF_TRY(rc.part_nn.alloc(rc.nn_parts, c->stream));
  HF_TRY(rc.d.alloc(ld, c->stream));
  HF_TRY(rc.ad.alloc(ld, c->stream));
  HF_TRY(rc.inc.alloc(ld, c->stream));
  HF_TRY(rc.x0.alloc(ld, c->stream));
  hf_rc_reset(c);
  return HF_OK;
}

extern "C" int hf_set_recycle(hf_ctx* c, int32_t max_vectors) {
  if (!c || max_vectors < 0 || max_vectors >= RC_CT) return hf_fail(HF_ERR_ARG, "hf_set_recycle: max_vectors must be in [0, 1023]");
  cudaSetDevice(c->device);
  Recycle& rc = c->rc;
  if (max_vectors != rc.cap) {
    rc.W.release();
    rc.ld = 0;
  }
  rc.cap = max_vectors;
  hf_rc_reset(c);
  return HF_OK;
}

// ws.x = x0, ws.r = r0 = bhat - Ahat x0 and ctrl.part_rr[0][0 .. ws.grid) hold the caller's values;
// on return they hold the projected ones and rc.x0 keeps x0.  A correction left pending by the
// previous solve is finalised on the way (it becomes vector `count`).
// Once `cap` corrections are stored the basis is frozen: every vector is Ahat-orthogonal to the others
// and carries a direction the later solutions keep using (the low-order Krylov vectors of the time
// stepping), so evicting the oldest one costs a full-length solve per step (measured); the late
// corrections are the small ones and are simply not recorded any more.
int hf_rc_project(hf_ctx* c) {
  Recycle& rc = c->rc;
  if (rc.cap == 0) return HF_OK;
  HF_TRY(rc_alloc(c));
  PcgWork& w = c->ws;
  const int m = rc.count, pend = rc.pending ? 1 : 0;
  const bool record = m + pend < rc.cap;          // this solve's correction will be stored
  if (m + pend == 0) {
    HF_CUDA(cudaMemcpyAsync(rc.x0.p, w.x.p, sizeof(double) * c->Npad, cudaMemcpyDeviceToDevice, c->stream));
    return HF_OK;
  }
  double* parts2 = rc.parts.p + (size_t)(rc.cap + 1) * rc.nseg;
  if (pend)
    k_rc_dots<true><<<rc.nseg, RC_WARPS * 32, 0, c->stream>>>(m + 1, rc.nseg, rc.ld, rc.W.p, w.r.p, rc.ad.p, rc.parts.p, parts2, m, rc.d.p);
  else
    k_rc_dots<false><<<rc.nseg, RC_WARPS * 32, 0, c->stream>>>(m, rc.nseg, rc.ld, rc.W.p, w.r.p, nullptr, rc.parts.p, parts2, -1, nullptr);
  k_rc_coef_project<<<1, RC_CT, 0, c->stream>>>(m, pend, rc.nseg, rc.parts.p, parts2, rc.inv.p, rc.hn.p, rc.nn_parts, rc.part_nn.p,
                                                rc.coef.p);
  HfCtrl* ctl = w.ctrl.p;
  const size_t smem = sizeof(double) * (2 * (size_t)m + 1);
  if (pend)
    k_rc_apply<true><<<w.grid, HF_BLOCK, smem, c->stream>>>(m, c->Npad, rc.ld, rc.W.p, rc.coef.p, rc.hn.p, rc.d.p, w.x.p, rc.inc.p,
                                                            record ? rc.x0.p : nullptr);
  else
    k_rc_apply<false><<<w.grid, HF_BLOCK, smem, c->stream>>>(m, c->Npad, rc.ld, rc.W.p, rc.coef.p, rc.hn.p, rc.d.p, w.x.p, rc.inc.p,
                                                             record ? rc.x0.p : nullptr);
  k_rc_resid<<<w.grid, HF_BLOCK, 0, c->stream>>>(c->opA.view(), rc.inc.p, w.r.p, &ctl->part_rr[0][0]);
  c->stat_launches += 4;
  HF_CUDA(cudaGetLastError());
  rc.count = m + pend;
  rc.pending = false;
  return HF_OK;
}

// ws.x = converged xhat: d = x - x0, Ad and the Gram-Schmidt coefficients against the stored basis;
// the vector itself is finalised by the next hf_rc_project.
int hf_rc_store(hf_ctx* c, const SellOp& op) {
  Recycle& rc = c->rc;
  if (rc.cap == 0 || rc.count >= rc.cap) return HF_OK;
  PcgWork& w = c->ws;
  const int m = rc.count;
  k_rc_spmv<<<rc.nn_parts, HF_BLOCK, 0, c->stream>>>(op.view(), w.x.p, rc.x0.p, rc.d.p, rc.ad.p, rc.part_nn.p);
  c->stat_launches += 1;
  (void)m;                             // the Gram-Schmidt coefficients are taken by the next hf_rc_project (w_k . Ad)
  HF_CUDA(cudaGetLastError());
  rc.pending = true;
  return HF_OK;
}
