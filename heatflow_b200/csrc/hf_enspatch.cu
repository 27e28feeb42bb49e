// heatflow_b200 - batched on-chip Jacobi-PCG for parameter_sweep tiles (sm_100a).
//
// north_star (c) for meshes that fit on chip: B variants of one simulation (sample conductivity k_b, heating width
// fwhm_b; parameter_sweep.py:123-192) are solved by ONE cooperative launch per time step.  Same decomposition as the
// single-simulation kernel (hf_patch.cu): one CTA of 256 threads per SM owns a Hilbert patch of 1024 rows, the
// first EP_K entries of every row live in registers, the rest and the direction vectors in shared memory,
// neighbouring CTAs exchange flag-with-data packets on their halo rows and there is one exact fixed-point grid
// reduction per iteration (hf_persist.cuh) - but a thread carries its rows for all B variants:
//   * the operator (values, 16-bit local columns) is read once per entry for the B variants; variant b's operator is
//     base0 + k_b S0 (hf_ensemble.cu) - S0 entries only exist in the slices that touch sample cells, and when all
//     variants of the tile share k (a sweep sorted by k: 64 heating widths per conductivity) they are folded
//     into the values on the host side of the launch;
//   * the latency of the grid reduction (~1.5 us from the arrival of the partial sums to the result) is paid once per
//     iteration of the whole tile.  NH = 2 (HF_ENS_NH) advances the tile as two half-tiles in lock step, the
//     reduction of one half in flight while the CTA computes the SpMV of the other (software pipelining across
//     variants instead of across iterations: classic CG recurrences, no extra vectors) - measured slower (10.2 us per
//     iteration of the tile against 7.2 us) because the per-half passes repeat the operator decode and the halo fetch
//     of a half only starts after the other half's SpMV; kept as a tested option, NH = 1 is the default.
// Measured (B200, N = 141 783, 139 CTAs, tile of 4 with one conductivity, clock64 per phase, cycles per iteration of
// the tile): SpMV + dot products 6300, arrive 1800, reduction trip 3300 (halo fetch 1750 hidden in it), updates 2200:
// 7.2 - 8.2 us = 1.8 - 2.1 us per variant and iteration against 2.3 us for the single-simulation pipelined kernel.  What
// bounds it is shared-memory bandwidth, not latency: every row moves ~560 B per iteration for the four variants
// (272 B of gathers, the rest own-row p / z / w / 1/d reads and writes) = 4500 cycles at 128 B/clk, and random 16-byte
// gathers reach 54 B/clk (tools/ubench_fp64.cu) - batching shares the operator and the reduction but not the
// vector traffic, so 4 variants cost ~3 x one.  End to end a tile of 4 takes 32 - 33 ms per simulation (100 steps,
// recycled bases) against 36 ms alone and 30 - 31 ms with two simulations sharing the SMs (sweep 'serial' engine), which
// therefore stays the default of parameter_sweep on on-chip meshes; `--mode ensemble --batch 4` selects this kernel.
// Formulation as in hf_ensemble.cu: z = D^-1 r, p, w = D^-1 A p, so halo rows need no per-variant scaling:
//     t = A_b p ; w = t / d ;  (p,t), (z,t), (t,w)  -> one reduction ;  alpha = rz / (p,t)
//     x += alpha p ; z -= alpha w ; rz' = rz - 2 alpha (z,t) + alpha^2 (t,w) ; p = z + (rz'/rz) p
// rz is recomputed directly (sum z^2 d) every HF_RR_CHECK iterations, when it has dropped by 1e4 and before a
// variant is declared converged (same policy as k_pcg_patch).  Converged variants are frozen (alpha = beta = 0).
#include <algorithm>
#include <cmath>

#include "hf_ens.cuh"
#include "hf_persist.cuh"

#define EP_T 256
#define EP_RPT 4
#define EP_R (EP_T * EP_RPT)
#define EP_K 8                 // register-cached operator entries per row
#define EP_NREP 8              // replicated accumulator lines per (half-tile, parity, variant)
#define EP_LINE 80             // 64-bit words between accumulator lines (640 B: neighbouring lines map to other L2 slices)
#define EP_W (EP_T / 32)

struct EnsOcArgs {
  int nslices, nrows, max_it, mat_cap, s0_cap, halo_cap, eb_shift;   // nrows: rows of the mesh (rows beyond it are padding)
  const int* slice_ptr;
  const double* val;             // sliced-ELL values: base0 (+ k S0 for uniform tiles)
  const double* s0;              // sliced-ELL values of S0
  const int* sflag;              // [nslices]
  const unsigned short* lcol;    // local columns for chunks of EP_R rows
  const int* halo_ptr;
  const int* halo_idx;
  const unsigned char* pub;
  const double* ks;              // [B]
  const double* dg;              // [rows, B] diagonal of A_b, 0 on Dirichlet rows
  double* x;                     // [rows, B] in: x0, out: x
  const double* z;               // [rows, B] z0 = D^-1 r0
  uint4* qpk;                    // [2][rows * B]
  size_t pk_stride, vec_len;     // packets per parity buffer; length of the [row, variant] vectors
  EnsCtrl* c;
  unsigned long long* acc;
  int* iters_out;
  int* fail;
  unsigned* gen;                 // [2] packet / reduction generation per half-tile, monotonic across launches: the packet
                                 // buffers still hold the previous solve's packets, whose tags must never match again
  long long* phase;              // diagnostics (-DHF_PHASE_TIMING): [G][2][8] clock64 cycles per phase, warps 0 and 1
};

// ---- exact one-trip grid reduction (see hf_persist.cuh) for NV values per variant of one half-tile -------------
// Accumulator lines: (half g, parity, replica r, variant bl) -> one 128-byte line holding NV chunks of (hi, lo).
template <int BG>
__device__ __forceinline__ unsigned long long* ep_line(unsigned long long* acc, int g, unsigned gen, int r, int bl) {
  return acc + ((size_t)((g * 2 + (int)(gen & 1u)) * EP_NREP + r) * BG + bl) * EP_LINE;
}

// v[bl * NV + j]: this thread's partial of value j of variant bl.  eb_of(bl, j) = exponent bound of that value.
template <int BG, int NV, typename EbFn>
__device__ __forceinline__ void ep_arrive(double (&v)[BG * NV], EbFn eb_of, unsigned long long* acc, int g, unsigned gen,
                                          double* sred /*[2][EP_W][BG*NV]*/, int* fail) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sh = sred + (gen & 1u) * (EP_W * BG * NV);
#pragma unroll
  for (int i = 0; i < BG * NV; ++i) {
    const double t = hf_warp_sum(v[i]);
    if (lane == 0) sh[warp * (BG * NV) + i] = t;
  }
  __syncthreads();
  if (threadIdx.x < BG * NV) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < EP_W; ++w) t += sh[w * (BG * NV) + threadIdx.x];
    const int bl = threadIdx.x / NV, j = threadIdx.x % NV;
    const int e = eb_of(bl, j);
    const double x = hf_scale2(t, 48 - e);
    long long hi;
    unsigned long long lo;
    if (fabs(x) < 281474976710656.0) {
      const double f = floor(x);
      hi = (long long)f;
      lo = (unsigned long long)((x - f) * 281474976710656.0);
    } else {                                               // not finite or out of range: the solve is reported as failed
      hi = 1ll << 54;
      lo = 0ull;
      atomicAdd(fail, 1);
    }
    unsigned long long* line = ep_line<BG>(acc, g, gen, blockIdx.x % EP_NREP, bl) + 2 * j;
    hf_red_add(line, ((unsigned long long)hi << 8) + 1ull);
    hf_red_add(line + 1, (lo << 8) + 1ull);
  }
}

// Warp 0 polls; lane l reads the line of (replica l >> 2, variant l & 3).  On return lanes bl < BG of warp 0 hold
// the NV totals of variant bl in out[]; other threads hold garbage.  prev: [2 parities][32 lanes][2 NV] words.
template <int BG, int NV, int NVMAX, typename EbFn>
__device__ __forceinline__ void ep_wait(double (&out)[NV], EbFn eb_of, unsigned long long* acc, int g, unsigned gen, int G,
                                        unsigned long long* prev, int* fail, int delay) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  const int r = lane >> 2, bl = lane & 3;
  const bool mine = bl < BG && r < G && r < EP_NREP;
  const unsigned long long mem = mine ? (unsigned long long)((G - 1 - r) / EP_NREP + 1) : 0ull;
  const unsigned long long* line = ep_line<BG>(acc, g, gen, mine ? r : 0, mine ? bl : 0);
  unsigned long long* pv = prev + ((size_t)(gen & 1u) * 32 + lane) * (2 * NVMAX);
  unsigned long long w[2 * NV], p[2 * NV];
#pragma unroll
  for (int i = 0; i < 2 * NV; ++i) {
    p[i] = pv[i];
    w[i] = 0ull;
  }
  if (delay > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < delay) {}
  }
  bool timed_out = false;
  for (int spins = 0;; ++spins) {
    bool ok = true;
    if (spins > HF_SPIN_MAX) {
      timed_out = true;
      break;
    }
    if (mine) {
#pragma unroll
      for (int j = 0; j < NV; ++j) hf_ld2(line + 2 * j, w[2 * j], w[2 * j + 1]);
#pragma unroll
      for (int i = 0; i < 2 * NV; ++i) ok = ok && (((w[i] - p[i]) & 0xffull) == mem);
    }
    if (__all_sync(0xffffffffu, ok)) break;
  }
  long long shi[NV];
  unsigned long long slo[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    shi[j] = mine ? ((long long)(w[2 * j] - p[2 * j] - mem) >> 8) : 0ll;
    slo[j] = mine ? ((w[2 * j + 1] - p[2 * j + 1] - mem) >> 8) : 0ull;
  }
  if (mine) {
#pragma unroll
    for (int i = 0; i < 2 * NV; ++i) pv[i] = w[i];
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      shi[j] += __shfl_xor_sync(0xffffffffu, shi[j], o);
      slo[j] += __shfl_xor_sync(0xffffffffu, slo[j], o);
    }
  }
  if (timed_out && lane == 0) atomicAdd(fail, 1);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int e = eb_of(bl < BG ? bl : 0, j);
    const bool bad = timed_out || shi[j] >= (1ll << 53) || shi[j] <= -(1ll << 53);
    const double val = hf_scale2(fma((double)shi[j], 281474976710656.0, (double)slo[j]), e - HF_FX_BITS);
    out[j] = bad ? __longlong_as_double(0x7ff8000000000000ll) : val;
  }
}

// ---------------------------------------------------------------------------------------
// The kernel: a tile of 4 variants in NH half-tiles of BG = 4 / NH variants.
// Shared-memory vectors are stored pair-major: element (row, b) of an array with `rows` rows sits at
// ((b >> 1) * rows + row) * 2 + (b & 1), i.e. one double2 per (variant pair, row).  A gather of column c for a pair is
// one 16-byte load whose bank group is c mod 8 - consecutive columns never collide, whereas the [row][4] layout
// (32-byte stride) leaves half of the banks idle and serialises neighbouring columns two by two.
// ---------------------------------------------------------------------------------------
#ifdef HF_PHASE_TIMING
#define EP_PT_DECL long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_t0 = clock64();
#define EP_PT_MARK(i)                 \
  do {                                \
    const long long t1__ = clock64(); \
    pt_acc[i] += t1__ - pt_t0;        \
    pt_t0 = t1__;                     \
  } while (0)
#define EP_PT_STORE                                                                         \
  if (P.phase && lane == 0 && warp < 2)                                                     \
    for (int i__ = 0; i__ < 8; ++i__) P.phase[((size_t)blockIdx.x * 2 + warp) * 8 + i__] = pt_acc[i__];
#else
#define EP_PT_DECL
#define EP_PT_MARK(i)
#define EP_PT_STORE
#endif

template <int NH>
__global__ void __launch_bounds__(EP_T, 1) k_ens_patch(EnsOcArgs P) {
  constexpr int B = 4;
  constexpr int BG = B / NH;          // variants per half-tile
  constexpr int NPG = BG / 2;         // variant pairs per half-tile
  constexpr int R = EP_R;
  constexpr int NSL = R / 32;
  constexpr int NVM = 3;
  static_assert(NSL == 32, "one slice per lane in the offset scan");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int RH = R + P.halo_cap, HC = P.halo_cap;
  double* sval = reinterpret_cast<double*>(smem_raw);            // operator entries beyond the register cache
  double* ss0 = sval + P.mat_cap;                                // S0 entries of the flagged slices
  double2* sp = reinterpret_cast<double2*>(ss0 + P.s0_cap);      // p: [2 pairs][R + halo]
  double2* szh = sp + 2 * (size_t)RH;                            // z on the halo rows: [2 pairs][halo]
  double2* swh = szh + 2 * (size_t)HC;                           // w on the halo rows (validated packets)
  double2* sdinv = swh + 2 * (size_t)HC;                         // 1 / d on the own rows (0 on Dirichlet rows): [2][R]
  double2* swo = sdinv + 2 * (size_t)R;                          // w on the own rows (phase 1 -> phase 2): [2][R]
  double2* szo = swo + 2 * (size_t)R;                            // z on the own rows: [2][R]
  double* sred = reinterpret_cast<double*>(szo + 2 * (size_t)R);     // [NH][2][EP_W][BG * 3]
  double* s_alpha = sred + NH * 2 * EP_W * BG * NVM;             // per variant: alpha, beta, rz, rz_ref, pp, thr, ks
  double* s_beta = s_alpha + B;
  double* s_rz = s_beta + B;
  double* s_ref = s_rz + B;
  double* s_pp = s_ref + B;
  double* s_thr = s_pp + B;
  double* s_ks = s_thr + B;
  unsigned long long* s_prev = reinterpret_cast<unsigned long long*>(s_ks + B);   // [NH][2][32][2 * 3]
  int* s_act = reinterpret_cast<int*>(s_prev + NH * 2 * 32 * 2 * NVM);           // [B]
  int* s_ctl = s_act + B;                                        // [NH][4]: check, done, since, -
  int* shal = s_ctl + NH * 4;
  int* sbase = shal + HC;                                        // [NSL + 1] offsets into sval / scol
  int* sb0 = sbase + ((NSL + 4) & ~3);                           // [NSL] offsets into ss0, -1 = slice has no S0 entries
  unsigned short* scol = reinterpret_cast<unsigned short*>(sb0 + NSL);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, nsl = P.nslices;
  const int f = blockIdx.x * NSL;
  const int lo = blockIdx.x * R;
  const int hp = P.halo_ptr[blockIdx.x];
  const int nh = P.halo_ptr[blockIdx.x + 1] - hp;
  // ---- slice offsets: overflow entries (beyond EP_K) and S0 blocks
  if (warp == 0) {
    const int s = f + lane;
    const int wdt = (s < nsl) ? ((P.slice_ptr[s + 1] - P.slice_ptr[s]) >> 5) : 0;
    const int cnt = max(wdt - EP_K, 0) * 32;
    const int cnt0 = (s < nsl && P.sflag[s]) ? wdt * 32 : 0;
    int incl = cnt, incl0 = cnt0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o), v0 = __shfl_up_sync(0xffffffffu, incl0, o);
      if (lane >= o) {
        incl += v;
        incl0 += v0;
      }
    }
    sbase[lane] = incl - cnt;
    sb0[lane] = cnt0 ? incl0 - cnt0 : -1;
    if (lane == 31) sbase[32] = incl;
  }
  if (tid < B) {
    s_ks[tid] = P.ks[tid];
    s_thr[tid] = P.c->thr[tid];
    const double rz = P.c->rz[tid];
    s_rz[tid] = s_ref[tid] = s_pp[tid] = rz;
    s_alpha[tid] = s_beta[tid] = 0.0;
    s_act[tid] = (rz > P.c->thr[tid]) ? 1 : 0;
  }
  for (int i = tid; i < NH * 2 * 32 * 2 * NVM; i += EP_T) s_prev[i] = 0ull;
  for (int i = tid; i < 2 * R; i += EP_T) {                      // (pair q, row): 16 contiguous bytes in global memory
    const int q = i / R, row = i - q * R;
    const size_t gidx = ((size_t)lo + row) * B + 2 * q;
    double2 zv = make_double2(0.0, 0.0), dv = zv;
    if (gidx < P.vec_len) {
      zv = *reinterpret_cast<const double2*>(P.z + gidx);        // p_0 = z_0 (padding rows hold zeros)
      dv = *reinterpret_cast<const double2*>(P.dg + gidx);
    }
    sp[(size_t)q * RH + row] = zv;
    szo[(size_t)q * R + row] = zv;
    sdinv[(size_t)q * R + row] = make_double2(dv.x > 0.0 ? 1.0 / dv.x : 0.0, dv.y > 0.0 ? 1.0 / dv.y : 0.0);
  }
  for (int h = tid; h < nh; h += EP_T) shal[h] = P.halo_idx[hp + h];
  __syncthreads();
  for (int i = tid; i < 2 * nh; i += EP_T) {
    const int q = i / nh, h = i - q * nh;
    const double2 zv = *reinterpret_cast<const double2*>(P.z + (size_t)shal[h] * B + 2 * q);
    sp[(size_t)q * RH + R + h] = zv;
    szh[(size_t)q * HC + h] = zv;
  }
  int wid[EP_RPT], base[EP_RPT], b0off[EP_RPT];
  double x[EP_RPT][B];
  double mv[EP_RPT][EP_K];
  unsigned mc[EP_RPT][EP_K / 2];
  unsigned pubmask = 0u;
#pragma unroll
  for (int k = 0; k < EP_RPT; ++k) {
    const int sl = warp * EP_RPT + k, s = f + sl;
    wid[k] = -1;
    base[k] = 0;
    b0off[k] = -1;
#pragma unroll
    for (int b = 0; b < B; ++b) x[k][b] = 0.0;
#pragma unroll
    for (int kk = 0; kk < EP_K; ++kk) mv[k][kk] = 0.0;
#pragma unroll
    for (int kk = 0; kk < EP_K / 2; ++kk) mc[k][kk] = (unsigned)(sl * 32 + lane) * 0x10001u;
    if (s < nsl) {
      const int p0 = P.slice_ptr[s];
      const int wd = (P.slice_ptr[s + 1] - p0) >> 5;
      wid[k] = wd;
      base[k] = sbase[sl] + lane;
      b0off[k] = sb0[sl];
      const size_t gi = ((size_t)s * 32 + lane) * B;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const double2 xv = *reinterpret_cast<const double2*>(P.x + gi + 2 * q);
        x[k][2 * q] = xv.x;
        x[k][2 * q + 1] = xv.y;
      }
      if (P.pub[s * 32 + lane]) pubmask |= 1u << k;
#pragma unroll
      for (int kk = 0; kk < EP_K; ++kk)
        if (kk < wd) {
          mv[k][kk] = P.val[p0 + kk * 32 + lane];
          const unsigned cc = P.lcol[p0 + kk * 32 + lane];
          mc[k][kk >> 1] = (kk & 1) ? ((mc[k][kk >> 1] & 0xffffu) | (cc << 16)) : ((mc[k][kk >> 1] & 0xffff0000u) | cc);
        }
      for (int kk = EP_K; kk < wd; ++kk) {
        sval[base[k] + (kk - EP_K) * 32] = P.val[p0 + kk * 32 + lane];
        scol[base[k] + (kk - EP_K) * 32] = P.lcol[p0 + kk * 32 + lane];
      }
      if (b0off[k] >= 0)
        for (int kk = 0; kk < wd; ++kk) ss0[b0off[k] + kk * 32 + lane] = P.s0[p0 + kk * 32 + lane];
    }
  }
  if (tid < NH * 4) s_ctl[tid] = 0;
  __syncthreads();
  if (tid < NH) {                                                // a half-tile whose variants all start converged is done
    bool any = false;
    for (int bl = 0; bl < BG; ++bl) any = any || s_act[tid * BG + bl];
    s_ctl[tid * 4 + 1] = any ? 0 : 1;
  }
  __syncthreads();
  unsigned gen[NH];
  int it[NH];
  bool done[NH];
#pragma unroll
  for (int g = 0; g < NH; ++g) {
    gen[g] = P.gen[g];
    it[g] = 0;
    done[g] = s_ctl[g * 4 + 1] != 0;
  }
  EP_PT_DECL

  // ---- phase 1 of half-tile g: t = A_b p on the own rows, w = t / d, publish, partial dot products, arrive
  auto phase1 = [&](const int g) {
    ++gen[g];
    uint4* qout = P.qpk + (size_t)(it[g] & 1) * P.pk_stride;
    double d[BG * 3];
#pragma unroll
    for (int i = 0; i < BG * 3; ++i) d[i] = 0.0;
#pragma unroll
    for (int k = 0; k < EP_RPT; ++k) {
      if (wid[k] >= 0) {
        double a[BG];
#pragma unroll
        for (int bl = 0; bl < BG; ++bl) a[bl] = 0.0;
        const int ov = wid[k] - EP_K, bs = base[k];
        if (b0off[k] < 0) {                                      // warp-uniform: the slice has no S0 entries
#pragma unroll
          for (int kk = 0; kk < EP_K; ++kk) {
            const int c = (int)((mc[k][kk >> 1] >> (16 * (kk & 1))) & 0xffffu);
#pragma unroll
            for (int ql = 0; ql < NPG; ++ql) {
              const double2 pv = sp[(size_t)(g * NPG + ql) * RH + c];
              a[2 * ql] = fma(mv[k][kk], pv.x, a[2 * ql]);
              a[2 * ql + 1] = fma(mv[k][kk], pv.y, a[2 * ql + 1]);
            }
          }
          for (int kk = 0; kk < ov; ++kk) {
            const double m = sval[bs + kk * 32];
            const int c = scol[bs + kk * 32];
#pragma unroll
            for (int ql = 0; ql < NPG; ++ql) {
              const double2 pv = sp[(size_t)(g * NPG + ql) * RH + c];
              a[2 * ql] = fma(m, pv.x, a[2 * ql]);
              a[2 * ql + 1] = fma(m, pv.y, a[2 * ql + 1]);
            }
          }
        } else {
          const double* s0p = ss0 + b0off[k] + lane;
#pragma unroll
          for (int kk = 0; kk < EP_K; ++kk) {
            const int c = (int)((mc[k][kk >> 1] >> (16 * (kk & 1))) & 0xffffu);
            const double sv = (kk < wid[k]) ? s0p[kk * 32] : 0.0;
#pragma unroll
            for (int ql = 0; ql < NPG; ++ql) {
              const double2 pv = sp[(size_t)(g * NPG + ql) * RH + c];
              a[2 * ql] = fma(fma(s_ks[g * BG + 2 * ql], sv, mv[k][kk]), pv.x, a[2 * ql]);
              a[2 * ql + 1] = fma(fma(s_ks[g * BG + 2 * ql + 1], sv, mv[k][kk]), pv.y, a[2 * ql + 1]);
            }
          }
          for (int kk = 0; kk < ov; ++kk) {
            const double m = sval[bs + kk * 32], sv = s0p[(kk + EP_K) * 32];
            const int c = scol[bs + kk * 32];
#pragma unroll
            for (int ql = 0; ql < NPG; ++ql) {
              const double2 pv = sp[(size_t)(g * NPG + ql) * RH + c];
              a[2 * ql] = fma(fma(s_ks[g * BG + 2 * ql], sv, m), pv.x, a[2 * ql]);
              a[2 * ql + 1] = fma(fma(s_ks[g * BG + 2 * ql + 1], sv, m), pv.y, a[2 * ql + 1]);
            }
          }
        }
        const int i = (warp * EP_RPT + k) * 32 + lane;
#pragma unroll
        for (int ql = 0; ql < NPG; ++ql) {
          const int q = g * NPG + ql;
          const double2 di = sdinv[(size_t)q * R + i];
          const double2 pv = sp[(size_t)q * RH + i], zv = szo[(size_t)q * R + i];
          const double t0 = a[2 * ql], t1 = a[2 * ql + 1];
          const double w0 = t0 * di.x, w1 = t1 * di.y;
          swo[(size_t)q * R + i] = make_double2(w0, w1);
          if (pubmask & (1u << k)) {
            hf_pkt_store(qout + (size_t)(lo + i) * B + 2 * q, w0, gen[g]);
            hf_pkt_store(qout + (size_t)(lo + i) * B + 2 * q + 1, w1, gen[g]);
          }
          d[(2 * ql) * 3 + 0] = fma(pv.x, t0, d[(2 * ql) * 3 + 0]);
          d[(2 * ql) * 3 + 1] = fma(zv.x, t0, d[(2 * ql) * 3 + 1]);
          d[(2 * ql) * 3 + 2] = fma(t0, w0, d[(2 * ql) * 3 + 2]);
          d[(2 * ql + 1) * 3 + 0] = fma(pv.y, t1, d[(2 * ql + 1) * 3 + 0]);
          d[(2 * ql + 1) * 3 + 1] = fma(zv.y, t1, d[(2 * ql + 1) * 3 + 1]);
          d[(2 * ql + 1) * 3 + 2] = fma(t1, w1, d[(2 * ql + 1) * 3 + 2]);
        }
      }
    }
    EP_PT_MARK(0);
    auto eb_of = [&](int bl, int j) {
      const int e_pp = hf_exp2(s_pp[g * BG + bl]), e_rr = hf_exp2(s_rz[g * BG + bl]);
      const int e = (j == 0) ? e_pp + 4 + HF_FX_MARGIN - P.eb_shift : (j == 1) ? (e_rr + e_pp + 1) / 2 + 5 + HF_FX_MARGIN : e_pp + 8 + HF_FX_MARGIN;
      return hf_clamp_exp(e);
    };
    ep_arrive<BG, 3>(d, eb_of, P.acc, g, gen[g], sred + g * (2 * EP_W * BG * NVM), P.fail);
    EP_PT_MARK(1);
  };

  // ---- phase 2 of half-tile g: halo packets, wait for the sums, updates, direct residual check when due
  auto phase2 = [&](const int g) {
    const uint4* qin = P.qpk + (size_t)(it[g] & 1) * P.pk_stride;
    if (warp >= 1) {
      constexpr int NT = EP_T - 32;
      constexpr int NF = (BG == 4) ? 5 : 3;                    // packets in flight per thread: one round covers ~280 halo rows
      for (int h0 = tid - 32; h0 < nh * BG; h0 += NF * NT) {
        uint4 hq[NF];
        bool need[NF];
        size_t src[NF];
        int dst[NF];
#pragma unroll
        for (int t = 0; t < NF; ++t) {
          const int idx = h0 + t * NT;
          need[t] = idx < nh * BG;
          const int h = need[t] ? idx / BG : 0, b = g * BG + (need[t] ? idx % BG : 0);
          src[t] = (size_t)shal[h] * B + b;
          dst[t] = ((b >> 1) * HC + h) * 2 + (b & 1);
        }
        bool pending;
        int spins = 0;
        do {
          pending = false;
          if (++spins > HF_SPIN_MAX) {
            atomicAdd(P.fail, 1);
            break;
          }
#pragma unroll
          for (int t = 0; t < NF; ++t)
            if (need[t]) hq[t] = hf_pkt_load(qin + src[t]);
#pragma unroll
          for (int t = 0; t < NF; ++t)
            if (need[t]) {
              if (hf_pkt_ok(hq[t], gen[g])) {
                need[t] = false;
                reinterpret_cast<double*>(swh)[dst[t]] = hf_pkt_val(hq[t]);
              } else {
                pending = true;
              }
            }
        } while (pending);
      }
    }
    EP_PT_MARK(2);
    auto eb3 = [&](int bl, int j) {
      const int e_pp = hf_exp2(s_pp[g * BG + bl]), e_rr = hf_exp2(s_rz[g * BG + bl]);
      const int e = (j == 0) ? e_pp + 4 + HF_FX_MARGIN - P.eb_shift : (j == 1) ? (e_rr + e_pp + 1) / 2 + 5 + HF_FX_MARGIN : e_pp + 8 + HF_FX_MARGIN;
      return hf_clamp_exp(e);
    };
    double tot[3];
    ep_wait<BG, 3, NVM>(tot, eb3, P.acc, g, gen[g], G, s_prev + (size_t)g * (2 * 32 * 2 * NVM), P.fail, NH == 1 ? HF_POLL_DELAY : 0);
    if (warp == 0) {                                             // lanes bl < BG: scalars of variant g * BG + bl
      const int b = g * BG + (lane < BG ? lane : 0);
      const double rz = s_rz[b];
      const bool act = s_act[b] != 0;
      double alpha = 0.0, rz_new = rz;
      if (act) {
        alpha = rz / tot[0];
        rz_new = fma(alpha * alpha, tot[2], fma(-2.0 * alpha, tot[1], rz));
      }
      const int since = s_ctl[g * 4 + 2] + 1;
      const bool want = lane < BG && act && (since >= HF_RR_CHECK || !(rz_new > s_thr[b]) || rz_new < 1e-4 * s_ref[b]);
      const bool check = __any_sync(0xffffffffu, want);
      if (lane < BG) {
        const double beta = (act && !check) ? rz_new / rz : 0.0;
        s_alpha[b] = alpha;
        s_beta[b] = beta;
        if (!check) {
          s_pp[b] = fma(beta * beta, s_pp[b], fabs(rz_new));
          s_rz[b] = rz_new;
        }
      }
      if (lane == 0) {
        s_ctl[g * 4 + 0] = check ? 1 : 0;
        s_ctl[g * 4 + 2] = check ? 0 : since;
      }
    }
    __syncthreads();
    EP_PT_MARK(3);
    const bool check = s_ctl[g * 4 + 0] != 0;
    double al[BG], be[BG];
#pragma unroll
    for (int bl = 0; bl < BG; ++bl) {
      al[bl] = s_alpha[g * BG + bl];
      be[bl] = s_beta[g * BG + bl];
    }
    double dd[BG];
#pragma unroll
    for (int bl = 0; bl < BG; ++bl) dd[bl] = 0.0;
#pragma unroll
    for (int k = 0; k < EP_RPT; ++k) {
      if (wid[k] >= 0) {
        const int i = (warp * EP_RPT + k) * 32 + lane;
#pragma unroll
        for (int ql = 0; ql < NPG; ++ql) {
          const int q = g * NPG + ql;
          const double2 pv = sp[(size_t)q * RH + i], wv = swo[(size_t)q * R + i], zo = szo[(size_t)q * R + i];
          x[k][2 * q] = fma(al[2 * ql], pv.x, x[k][2 * q]);
          x[k][2 * q + 1] = fma(al[2 * ql + 1], pv.y, x[k][2 * q + 1]);
          const double z0 = fma(-al[2 * ql], wv.x, zo.x);
          const double z1 = fma(-al[2 * ql + 1], wv.y, zo.y);
          szo[(size_t)q * R + i] = make_double2(z0, z1);
          if (check) {
            const double2 dv = *reinterpret_cast<const double2*>(P.dg + (size_t)(lo + i) * B + 2 * q);
            dd[2 * ql] = fma(z0 * dv.x, z0, dd[2 * ql]);
            dd[2 * ql + 1] = fma(z1 * dv.y, z1, dd[2 * ql + 1]);
          } else {
            sp[(size_t)q * RH + i] = make_double2(fma(be[2 * ql], pv.x, z0), fma(be[2 * ql + 1], pv.y, z1));
          }
        }
      }
    }
    for (int idx = tid; idx < nh * NPG; idx += EP_T) {
      const int ql = idx / nh, h = idx - ql * nh;
      const int q = g * NPG + ql;
      const double a0 = s_alpha[2 * q], a1 = s_alpha[2 * q + 1];
      const double2 wv = swh[(size_t)q * HC + h], zo = szh[(size_t)q * HC + h];
      const double2 zv = make_double2(fma(-a0, wv.x, zo.x), fma(-a1, wv.y, zo.y));
      szh[(size_t)q * HC + h] = zv;
      if (!check) {
        const double2 pv = sp[(size_t)q * RH + R + h];
        sp[(size_t)q * RH + R + h] = make_double2(fma(s_beta[2 * q], pv.x, zv.x), fma(s_beta[2 * q + 1], pv.y, zv.y));
      }
    }
    ++it[g];
    if (check) {
      ++gen[g];
      auto eb1 = [&](int bl, int) {
        const int b = g * BG + bl;
        return hf_clamp_exp(max(hf_exp2(s_rz[b]), 2 * hf_exp2(fabs(s_alpha[b])) + 8 + hf_exp2(s_pp[b])) + 2 + HF_FX_MARGIN);
      };
      ep_arrive<BG, 1>(dd, eb1, P.acc, g, gen[g], sred + g * (2 * EP_W * BG * NVM), P.fail);
      double t1[1];
      ep_wait<BG, 1, NVM>(t1, eb1, P.acc, g, gen[g], G, s_prev + (size_t)g * (2 * 32 * 2 * NVM), P.fail, HF_POLL_DELAY);
      if (warp == 0) {
        const int b = g * BG + (lane < BG ? lane : 0);
        bool act = s_act[b] != 0;
        if (lane < BG) {
          double b2 = 0.0;
          if (act) {
            const double rz_new = t1[0];
            b2 = rz_new / s_rz[b];
            s_pp[b] = fma(b2 * b2, s_pp[b], fabs(rz_new));
            s_rz[b] = rz_new;
            s_ref[b] = rz_new;
            act = rz_new > s_thr[b];                              // false for NaN as well: the variant stops
            if (!act) b2 = 0.0;                                   // frozen: p = z
            s_act[b] = act ? 1 : 0;
          }
          s_beta[b] = b2;
        }
        const bool any = __any_sync(0xffffffffu, lane < BG && act);
        if (lane == 0) s_ctl[g * 4 + 1] = any ? 0 : 1;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < EP_RPT; ++k) {
        if (wid[k] >= 0) {
          const int i = (warp * EP_RPT + k) * 32 + lane;
#pragma unroll
          for (int ql = 0; ql < NPG; ++ql) {
            const int q = g * NPG + ql;
            const double2 pv = sp[(size_t)q * RH + i], zv = szo[(size_t)q * R + i];
            sp[(size_t)q * RH + i] = make_double2(fma(s_beta[2 * q], pv.x, zv.x), fma(s_beta[2 * q + 1], pv.y, zv.y));
          }
        }
      }
      for (int idx = tid; idx < nh * NPG; idx += EP_T) {
        const int ql = idx / nh, h = idx - ql * nh;
        const int q = g * NPG + ql;
        const double2 pv = sp[(size_t)q * RH + R + h], zv = szh[(size_t)q * HC + h];
        sp[(size_t)q * RH + R + h] = make_double2(fma(s_beta[2 * q], pv.x, zv.x), fma(s_beta[2 * q + 1], pv.y, zv.y));
      }
      done[g] = s_ctl[g * 4 + 1] != 0;
    }
    EP_PT_MARK(4);
    __syncthreads();
    EP_PT_MARK(5);
  };

  bool capped = false;
#pragma unroll
  for (int g = 0; g < NH; ++g)
    if (!done[g] && P.max_it > 0) phase1(g);
  for (;;) {
    bool all = true;
#pragma unroll
    for (int g = 0; g < NH; ++g) {
      if (!done[g]) {
        if (it[g] >= P.max_it) {
          capped = true;
          done[g] = true;
        } else {
          phase2(g);
          if (!done[g] && it[g] < P.max_it) phase1(g);
        }
      }
      all = all && done[g];
    }
    if (all) break;
  }
  EP_PT_STORE
#pragma unroll
  for (int k = 0; k < EP_RPT; ++k)
    if (wid[k] >= 0 && lo + (warp * EP_RPT + k) * 32 + lane < P.nrows) {      // padding rows keep their zeros
      const size_t gi = ((size_t)lo + (warp * EP_RPT + k) * 32 + lane) * B;
#pragma unroll
      for (int q = 0; q < 2; ++q) *reinterpret_cast<double2*>(P.x + gi + 2 * q) = make_double2(x[k][2 * q], x[k][2 * q + 1]);
    }
  if (blockIdx.x == 0 && tid == 0) {
    int itmax = 0;
    bool bad = capped;
#pragma unroll
    for (int g = 0; g < NH; ++g) itmax = max(itmax, it[g]);
    for (int b = 0; b < B; ++b) {
      P.c->rz[b] = s_rz[b];
      P.c->active[b] = s_act[b];
      bad = bad || s_act[b] || !isfinite(s_rz[b]);
    }
#pragma unroll
    for (int g = 0; g < NH; ++g) P.gen[g] = gen[g];
    P.c->it = itmax;
    P.c->done = bad ? 0 : 1;
    if (P.iters_out) *P.iters_out = itmax;
    if (bad) atomicAdd(P.fail, 1);
  }
}

// ---------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------
// CSR values -> sliced-ELL order of opA; kuni is the common conductivity of a uniform tile (values = base0 + kuni S0,
// no slice flagged) or NaN (values = base0, S0 kept, slices with a non-zero S0 entry flagged)
__global__ void k_ens_sell_fill(int N, int Npad, const int* __restrict__ rowptr, const int* __restrict__ slice_ptr,
                                const double* __restrict__ base0, const double* __restrict__ s0, double kuni,
                                double* __restrict__ out_v, double* __restrict__ out_s, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const int s = i / HF_SLICE, lane = i % HF_SLICE;
  const int b = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - b) / HF_SLICE;
  int len = 0, r0 = 0;
  if (i < N) {
    r0 = rowptr[i];
    len = rowptr[i + 1] - r0;
  }
  const bool uni = kuni == kuni;
  bool any = false;
  for (int k = 0; k < w; ++k) {
    double v = 0.0, sv = 0.0;
    if (k < len) {
      v = base0[r0 + k];
      sv = s0[r0 + k];
      if (uni) {
        v = fma(kuni, sv, v);
        sv = 0.0;
      }
      any = any || sv != 0.0;
    }
    out_v[b + k * HF_SLICE + lane] = v;
    out_s[b + k * HF_SLICE + lane] = sv;
  }
  if (any) flag[s] = 1;
}

// rz_b = sum z^2 d of the initial residual (after the recycled-basis projection): per-CTA partials, the last CTA adds
// them in CTA order
template <int LB>
__global__ void __launch_bounds__(HF_ET)
k_ens_rz(size_t total, const double* __restrict__ z, const double* __restrict__ dg, double* __restrict__ part, EnsCtrl* __restrict__ c) {
  constexpr int B = 1 << LB;
  __shared__ double sh[HF_ET];
  __shared__ int s_last;
  double l = 0.0;
  for (size_t idx = (size_t)blockIdx.x * HF_ET + threadIdx.x; idx < total; idx += (size_t)gridDim.x * HF_ET) {
    const double zv = z[idx];
    l = fma(zv * dg[idx], zv, l);
  }
  sh[threadIdx.x] = l;
  __syncthreads();
  if (threadIdx.x < B) {
    double t = 0.0;
    for (int i = threadIdx.x; i < HF_ET; i += B) t += sh[i];     // HF_ET is a multiple of B: thread i holds variant i & (B-1)
    __stcg(part + (size_t)blockIdx.x * B + threadIdx.x, t);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&c->counter[2], 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) c->counter[2] = 0u;
    __threadfence();
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= B) return;
  double rz = 0.0;
  for (unsigned cta = 0; cta < gridDim.x; ++cta) rz += __ldcg(part + (size_t)cta * B + threadIdx.x);
  c->rz[threadIdx.x] = rz;
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int NH>
static size_t ep_smem_bytes(int mat_cap, int s0_cap, int halo_cap) {
  constexpr int B = 4, BG = B / NH;
  size_t dbl = (size_t)mat_cap + s0_cap + (size_t)(EP_R + halo_cap) * B + 2 * (size_t)halo_cap * B + 3 * (size_t)EP_R * B +
               NH * 2 * EP_W * BG * 3 + 7 * B;
  size_t u64 = (size_t)NH * 2 * 32 * 2 * 3;
  size_t i32 = (size_t)B + NH * 4 + halo_cap + ((EP_R / 32 + 4) & ~3) + EP_R / 32;
  return dbl * 8 + u64 * 8 + i32 * 4 + (size_t)mat_cap * 2 + 16;
}

static const void* ep_kernel(int nh) { return nh == 2 ? (const void*)k_ens_patch<2> : (const void*)k_ens_patch<1>; }
static size_t ep_bytes(int nh, int mat_cap, int s0_cap, int halo_cap) {
  return nh == 2 ? ep_smem_bytes<2>(mat_cap, s0_cap, halo_cap) : ep_smem_bytes<1>(mat_cap, s0_cap, halo_cap);
}

// Plans the on-chip kernel for the tile in `e` (called by hf_ens_create): needs the patch plan of the single-simulation
// kernel at 1024 rows per CTA (same sparsity pattern, same halo lists) and the tile's operator within shared memory.
int hf_ens_oc_plan(hf_ctx* c, EnsState* e) {
  e->oc_ok = false;
  if (getenv("HF_ENS_STREAM")) return HF_OK;                     // tuning / test knob: keep the streaming kernels
  if (e->LB != 2) return HF_OK;                                  // tiles of 4 variants (smaller tiles are padded to 4)
  SellOp& op = c->opA;
  if (!op.pp_rpt) HF_TRY(hf_patch_plan(c, op));
  if (op.pp_rpt * HF_PT != EP_R || op.pp_grid > c->sm_count) return HF_OK;   // same 1024-row patches as the single-simulation kernel
  const int nsl = op.nslices, B = e->B;
  e->oc_rows = (size_t)e->nchunks * (HF_EPAIRS / B);
  if ((size_t)op.pp_grid * EP_R > e->oc_rows + EP_R) return HF_OK;
  // tile with one conductivity (sweeps sorted by k): fold k S0 into the values
  std::vector<double> ks(B);
  HF_TRY(e->ks.download(ks.data(), B, c->stream));
  e->oc_uniform = true;
  for (int b = 1; b < B; ++b) e->oc_uniform = e->oc_uniform && ks[b] == ks[0];
  HF_TRY(e->oc_val.alloc(op.padded_nnz, c->stream));
  HF_TRY(e->oc_s0.alloc(op.padded_nnz, c->stream));
  HF_TRY(e->oc_flag.alloc(nsl, c->stream));
  k_ens_sell_fill<<<(c->Npad + 255) / 256, 256, 0, c->stream>>>(c->N, c->Npad, c->rowptr.p, op.slice_ptr.p, e->base0.p, e->s0.p,
                                                               e->oc_uniform ? ks[0] : NAN, e->oc_val.p, e->oc_s0.p, e->oc_flag.p);
  HF_CUDA(cudaGetLastError());
  std::vector<int> sp(nsl + 1), flag(nsl);
  HF_CUDA(cudaMemcpyAsync(sp.data(), op.slice_ptr.p, sizeof(int) * (nsl + 1), cudaMemcpyDeviceToHost, c->stream));
  HF_TRY(e->oc_flag.download(flag.data(), nsl, c->stream));
  int mat_cap = 0, s0_cap = 0;
  const int G = op.pp_grid;
  for (int b = 0; b < G; ++b) {
    const int f = std::min(b * (EP_R / 32), nsl), fe = std::min(f + EP_R / 32, nsl);
    int n = 0, n0 = 0;
    for (int sl = f; sl < fe; ++sl) {
      const int w = (sp[sl + 1] - sp[sl]) >> 5;
      n += std::max(w - EP_K, 0) * 32;
      if (flag[sl]) n0 += w * 32;
    }
    mat_cap = std::max(mat_cap, n);
    s0_cap = std::max(s0_cap, n0);
  }
  mat_cap = (mat_cap + 7) & ~7;
  s0_cap = (s0_cap + 7) & ~7;
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
  int nh = 1;                                                    // measured: 7.2 us per iteration of the tile against 10.2 with two half-tiles
  if (const char* env = getenv("HF_ENS_NH")) nh = std::max(1, std::min(2, atoi(env)));
  const size_t bytes = ep_bytes(nh, mat_cap, s0_cap, op.pp_halo_cap);
  if (bytes > (size_t)max_smem) return HF_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
  if (!coop) return HF_OK;
  HF_CUDA(cudaFuncSetAttribute(ep_kernel(nh), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  e->oc_nh = nh;
  e->oc_grid = G;
  e->oc_mat_cap = mat_cap;
  e->oc_s0_cap = s0_cap;
  e->oc_halo_cap = op.pp_halo_cap;
  e->oc_smem = bytes;
  const size_t acc_words = (size_t)nh * 2 * EP_NREP * (B / nh) * EP_LINE;
  HF_TRY(e->oc_acc.alloc(acc_words, c->stream));
  HF_TRY(e->oc_qpk.alloc((size_t)2 * (e->oc_rows + EP_R) * B, c->stream));
  HF_TRY(e->oc_fail.alloc(1, c->stream));
  HF_TRY(e->oc_gen.alloc(2, c->stream));
  e->oc_ok = true;
  return HF_OK;
}

int hf_ens_oc_solve_async(hf_ctx* c, EnsState* e, int step_slot) {
  if (!e->oc_ok) return hf_fail(HF_ERR_STATE, "on-chip ensemble kernel is not planned for this tile");
  const SellOp& op = c->opA;
  const size_t total = e->oc_rows * e->B;
  k_ens_rz<2><<<e->grid, HF_ET, 0, c->stream>>>(total, e->z.p, e->dg.p, e->part.p, e->ctrl.p);
  HF_CUDA(cudaGetLastError());
  HF_CUDA(cudaMemsetAsync(e->oc_acc.p, 0, e->oc_acc.n * sizeof(unsigned long long), c->stream));
  EnsOcArgs a;
  a.nslices = op.nslices;
  a.nrows = c->N;
  a.max_it = c->max_iters;
  a.mat_cap = e->oc_mat_cap;
  a.s0_cap = e->oc_s0_cap;
  a.halo_cap = e->oc_halo_cap;
  a.eb_shift = c->debug_fx_shift;
  a.slice_ptr = op.slice_ptr.p;
  a.val = e->oc_val.p;
  a.s0 = e->oc_s0.p;
  a.sflag = e->oc_flag.p;
  a.lcol = op.pp_lcol.p;
  a.halo_ptr = op.pp_halo_ptr.p;
  a.halo_idx = op.pp_halo_idx.p;
  a.pub = op.pp_pub.p;
  a.ks = e->ks.p;
  a.dg = e->dg.p;
  a.x = e->x.p;
  a.z = e->z.p;
  a.qpk = e->oc_qpk.p;
  a.pk_stride = (e->oc_rows + EP_R) * e->B;
  a.vec_len = total;
  a.c = e->ctrl.p;
  a.acc = e->oc_acc.p;
  a.iters_out = (step_slot >= 0 && (size_t)step_slot < e->oc_iters.n) ? e->oc_iters.p + step_slot : nullptr;
  a.fail = e->oc_fail.p;
  a.gen = e->oc_gen.p;
  a.phase = c->debug_phase.n ? c->debug_phase.p : nullptr;
  void* args[] = {&a};
  HF_CUDA(cudaFuncSetAttribute(ep_kernel(e->oc_nh), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->oc_smem));
  HF_CUDA(cudaLaunchCooperativeKernel(ep_kernel(e->oc_nh), dim3(e->oc_grid), dim3(EP_T), args, e->oc_smem, c->stream));
  c->stat_launches += 2;
  return HF_OK;
}
