// heatflow_b200 - initial guess of the Jacobi-PCG solve from the previous time steps (sm_100a).
//
// The reference solves every time step with one LU factorisation (run_with_diamond.py:389-394,
// :480).  An iterative solver pays for every step from scratch unless it reuses what the earlier
// solves found: all steps share the operator  Ahat = D^-1/2 (M + dt K) D^-1/2,  and the solutions
// of a diffusion problem driven by one scalar amplitude stay close to a low-dimensional space.
// This file keeps an Ahat-orthogonal basis  W = [w_1 .. w_m]  of the corrections the first m solves
// computed, together with  AW = Ahat W  and  inv_k = 1 / (w_k . Ahat w_k):
//   before the solve   x0 <- x0 + W c,  r0 <- r0 - AW c,  c_k = inv_k (w_k . r0)
//                      (Galerkin projection: the error of x0 becomes Ahat-orthogonal to span W);
//   after the solve    d = x - x0,  Ad = Ahat d (one SpMV, so that AW = Ahat W holds to rounding),
//                      h_k = inv_k (Ahat w_k . d),  w_{m+1} = d - W h,  Ahat w_{m+1} = Ad - AW h
//                      (one Gram-Schmidt pass in the Ahat inner product; without it the computed
//                      corrections are only orthogonal to ~1e-4 and the projection stops paying).
// The Gram-Schmidt update is lazy: the pass over (W, AW) that applies h is the same pass that applies
// c at the next solve, so a step streams 4 m N doubles (AW for h, W for c, W and AW for the fused
// update) instead of 6 m N.  The new vector's coefficient follows from the raw dot products:
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

// parts[k * nseg + seg] = sum over the RC_SEG rows of segment `seg` of V[k][i] * v[i], k < m
// (vector m_extra is read from `extra` instead of its slot).
// One CTA per segment; every lane keeps its RC_SEG / 32 values of v in registers (consecutive-pair
// loads) and the warps share the basis vectors (warp w takes k = w, w + 8, ...), so each warp has
// RC_SEG / 64 independent 512-byte loads in flight per basis vector and no barrier is needed.
__global__ void __launch_bounds__(RC_WARPS * 32, 2)
k_rc_dots(int m, int nseg, size_t ld, const double* __restrict__ V, const double* __restrict__ v,
          double* __restrict__ parts, int m_extra, const double* __restrict__ extra) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seg = blockIdx.x;
  const size_t base = (size_t)seg * RC_SEG + 2 * lane;
  double2 vv[RC_NV];
#pragma unroll
  for (int j = 0; j < RC_NV; ++j) vv[j] = *reinterpret_cast<const double2*>(v + base + 64 * j);
  // two basis vectors per pass: 2 x RC_NV independent 16-byte loads per lane in flight (8 KB per warp) before the
  // first use, and the two shuffle trees overlap - the one-vector loop left the SM at 12 % warps active / 2.0 TB/s
  for (int k = warp; k < m; k += 2 * RC_WARPS) {
    const int k2 = k + RC_WARPS;
    const bool two = k2 < m;
    const double* p = ((k == m_extra) ? extra : V + (size_t)k * ld) + base;   // the pending raw correction is not in a slot yet
    const double* p2 = ((k2 == m_extra || !two) ? (two ? extra : p) : V + (size_t)k2 * ld) + (two ? base : 0);
    double2 a[RC_NV], b[RC_NV];
#pragma unroll
    for (int j = 0; j < RC_NV; ++j) a[j] = __ldcs(reinterpret_cast<const double2*>(p + 64 * j));
#pragma unroll
    for (int j = 0; j < RC_NV; ++j) b[j] = __ldcs(reinterpret_cast<const double2*>(p2 + 64 * j));
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < RC_NV; ++j) {
      s0 = fma(a[j].x, vv[j].x, s0);
      s1 = fma(a[j].y, vv[j].y, s1);
      t0 = fma(b[j].x, vv[j].x, t0);
      t1 = fma(b[j].y, vv[j].y, t1);
    }
    double s = s0 + s1, t = t0 + t1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    if (lane == 0) {
      parts[(size_t)k * nseg + seg] = s;
      if (two) parts[(size_t)k2 * nseg + seg] = t;
    }
  }
}

// coef[k] = sign * inv[k] * sum_seg parts[k][seg]   (one warp per k, fixed order)
__global__ void __launch_bounds__(RC_WARPS * 32)
k_rc_coef(int m, int nseg, const double* __restrict__ parts, const double* __restrict__ inv, double sign,
          double* __restrict__ coef) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * RC_WARPS + (threadIdx.x >> 5);
  if (k >= m) return;
  double s = 0.0;
  for (int i = lane; i < nseg; i += 32) s += __ldcg(parts + (size_t)k * nseg + i);
  s = hf_warp_sum(s);
  if (lane == 0) coef[k] = sign * inv[k] * s;
}

// Projection coefficients, one CTA.  t_k = sum_seg parts[k][seg] (one warp per k, fixed order);
// c_k = inv_k t_k for the m finalised vectors.  With a pending raw correction (index m, Gram-Schmidt
// coefficients hn_k = -h_k from the store phase, d.Ad partials in part_nn):
//   w.r0 = t_m + sum hn_k t_k,  w.Ahat w = d.Ad - sum hn_k^2 / inv_k,  inv_m = 1 / (w.Ahat w),  c_m = inv_m (w.r0)
// (inv_m = 0 when the correction is empty or not finite: the slot stays inert).
#define RC_CT 1024
__global__ void __launch_bounds__(RC_CT)
k_rc_coef_project(int m, int pending, int nseg, const double* __restrict__ parts, double* __restrict__ inv,
                  const double* __restrict__ hn, int nparts_nn, const double* __restrict__ part_nn, double* __restrict__ coef) {
  __shared__ double s_t[RC_CT];        // m + pending <= cap + 1 <= RC_CT is checked on the host
  __shared__ double s_red[3][RC_CT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mm = m + pending;
  for (int k = warp; k < mm; k += RC_CT / 32) {
    double s = 0.0;
    for (int i = lane; i < nseg; i += 32) s += __ldcg(parts + (size_t)k * nseg + i);
    s = hf_warp_sum(s);
    if (lane == 0) s_t[k] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < m; k += RC_CT) coef[k] = inv[k] * s_t[k];
  if (!pending) return;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int k = threadIdx.x; k < m; k += RC_CT) {
    const double h = hn[k], iv = inv[k];
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

// Fused pass over (W, AW) before the solve:  a = x0, b = r0.
//   PENDING: finalise the raw correction (d, ad) of the previous solve as vector m:
//            w = d + W hn,  aw = ad + AW hn  -> slot m of W / AW
//   a += W c (+ c_m w) ;  b -= AW c (+ c_m aw) ;  x0save = a (null once the basis is frozen) ;
//   partial sums of b.b -> part[blockIdx.x]
template <bool PENDING>
__global__ void __launch_bounds__(HF_BLOCK)
k_rc_apply(int m, int n, size_t ld, double* __restrict__ W, double* __restrict__ AW, const double* __restrict__ coef,
           const double* __restrict__ hn, const double* __restrict__ d, const double* __restrict__ ad, double* __restrict__ a,
           double* __restrict__ b, double* __restrict__ x0save, double* __restrict__ part) {
  extern __shared__ double s_c[];      // c[0..m] then hn[0..m)
  __shared__ double sh[HF_BLOCK / 32];
  double* s_h = s_c + m + 1;
  for (int k = threadIdx.x; k <= m; k += HF_BLOCK) s_c[k] = (k < m || PENDING) ? coef[k] : 0.0;
  if (PENDING)
    for (int k = threadIdx.x; k < m; k += HF_BLOCK) s_h[k] = hn[k];
  __syncthreads();
  double local = 0.0;
  // two rows per thread (16-byte loads; n and ld are even), four basis vectors per pass: 8 + 8 independent loads in
  // flight per thread.  Per row the sums run over k in the same order as a one-row loop would.
  for (int i = 2 * (blockIdx.x * HF_BLOCK + threadIdx.x); i < n; i += 2 * gridDim.x * HF_BLOCK) {
    double2 ca0 = {0.0, 0.0}, ca1 = {0.0, 0.0}, cb0 = {0.0, 0.0}, cb1 = {0.0, 0.0};
    double2 ga0 = {0.0, 0.0}, ga1 = {0.0, 0.0}, gb0 = {0.0, 0.0}, gb1 = {0.0, 0.0};
    const double* w = W + i;
    const double* aw = AW + i;
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
      const double2 z0 = ld2(aw + (size_t)k * ld), z1 = ld2(aw + (size_t)(k + 1) * ld), z2 = ld2(aw + (size_t)(k + 2) * ld),
                    z3 = ld2(aw + (size_t)(k + 3) * ld);
      const double c0 = s_c[k], c1 = s_c[k + 1], c2 = s_c[k + 2], c3 = s_c[k + 3];
      acc(c0, w0, ca0);
      acc(c1, w1, ca1);
      acc(c2, w2, ca0);
      acc(c3, w3, ca1);
      acc(c0, z0, cb0);
      acc(c1, z1, cb1);
      acc(c2, z2, cb0);
      acc(c3, z3, cb1);
      if (PENDING) {
        const double h0 = s_h[k], h1 = s_h[k + 1], h2 = s_h[k + 2], h3 = s_h[k + 3];
        acc(h0, w0, ga0);
        acc(h1, w1, ga1);
        acc(h2, w2, ga0);
        acc(h3, w3, ga1);
        acc(h0, z0, gb0);
        acc(h1, z1, gb1);
        acc(h2, z2, gb0);
        acc(h3, z3, gb1);
      }
    }
    for (; k < m; ++k) {
      const double2 wk = ld2(w + (size_t)k * ld), zk = ld2(aw + (size_t)k * ld);
      acc(s_c[k], wk, ca0);
      acc(s_c[k], zk, cb0);
      if (PENDING) {
        acc(s_h[k], wk, ga0);
        acc(s_h[k], zk, gb0);
      }
    }
    double2 ca = {ca0.x + ca1.x, ca0.y + ca1.y}, cb = {cb0.x + cb1.x, cb0.y + cb1.y};
    if (PENDING) {
      const double2 dv = *reinterpret_cast<const double2*>(d + i), adv = *reinterpret_cast<const double2*>(ad + i);
      const double2 wm = {dv.x + (ga0.x + ga1.x), dv.y + (ga0.y + ga1.y)};
      const double2 awm = {adv.x + (gb0.x + gb1.x), adv.y + (gb0.y + gb1.y)};
      *reinterpret_cast<double2*>(W + (size_t)m * ld + i) = wm;
      *reinterpret_cast<double2*>(AW + (size_t)m * ld + i) = awm;
      acc(s_c[m], wm, ca);
      acc(s_c[m], awm, cb);
    }
    const double2 av = *reinterpret_cast<const double2*>(a + i), bv = *reinterpret_cast<const double2*>(b + i);
    const double2 x = {av.x + ca.x, av.y + ca.y}, r = {bv.x - cb.x, bv.y - cb.y};
    *reinterpret_cast<double2*>(a + i) = x;
    *reinterpret_cast<double2*>(b + i) = r;
    if (x0save) *reinterpret_cast<double2*>(x0save + i) = x;
    local = fma(r.x, r.x, local);
    local = fma(r.y, r.y, local);
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
  HF_TRY(rc.AW.alloc((size_t)rc.cap * ld, c->stream));
  HF_TRY(rc.inv.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.coef.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.hn.alloc(rc.cap + 1, c->stream));
  HF_TRY(rc.parts.alloc((size_t)(rc.cap + 1) * rc.nseg, c->stream));
  HF_TRY(rc.part_nn.alloc(rc.nn_parts, c->stream));
  HF_TRY(rc.d.alloc(ld, c->stream));
  HF_TRY(rc.ad.alloc(ld, c->stream));
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
    rc.AW.release();
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
  k_rc_dots<<<rc.nseg, RC_WARPS * 32, 0, c->stream>>>(m + pend, rc.nseg, rc.ld, rc.W.p, w.r.p, rc.parts.p, pend ? m : -1, rc.d.p);
  k_rc_coef_project<<<1, RC_CT, 0, c->stream>>>(m, pend, rc.nseg, rc.parts.p, rc.inv.p, rc.hn.p, rc.nn_parts, rc.part_nn.p,
                                                rc.coef.p);
  HfCtrl* ctl = w.ctrl.p;
  const size_t smem = sizeof(double) * (2 * (size_t)m + 1);
  if (pend)
    k_rc_apply<true><<<w.grid, HF_BLOCK, smem, c->stream>>>(m, c->Npad, rc.ld, rc.W.p, rc.AW.p, rc.coef.p, rc.hn.p, rc.d.p, rc.ad.p,
                                                            w.x.p, w.r.p, record ? rc.x0.p : nullptr, &ctl->part_rr[0][0]);
  else
    k_rc_apply<false><<<w.grid, HF_BLOCK, smem, c->stream>>>(m, c->Npad, rc.ld, rc.W.p, rc.AW.p, rc.coef.p, rc.hn.p, rc.d.p, rc.ad.p,
                                                             w.x.p, w.r.p, record ? rc.x0.p : nullptr, &ctl->part_rr[0][0]);
  c->stat_launches += 3;
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
  if (m > 0) {
    k_rc_dots<<<rc.nseg, RC_WARPS * 32, 0, c->stream>>>(m, rc.nseg, rc.ld, rc.AW.p, rc.d.p, rc.parts.p, -1, nullptr);
    k_rc_coef<<<(m + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, 0, c->stream>>>(m, rc.nseg, rc.parts.p, rc.inv.p, -1.0, rc.hn.p);
    c->stat_launches += 2;
  }
  HF_CUDA(cudaGetLastError());
  rc.pending = true;
  return HF_OK;
}
