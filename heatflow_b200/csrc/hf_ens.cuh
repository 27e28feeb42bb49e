// heatflow_b200 - state of the batched multi-RHS ensemble (hf_ensemble.cu: streaming kernels, hf_enspatch.cu: on-chip kernel).
#pragma once
#include "hf_ctx.cuh"

#define HF_EB 32            // max variants per tile
#define HF_ET 256           // threads per CTA of the set-up kernels
#ifndef HF_ENT
#define HF_ENT 256          // threads per CTA of the iteration kernel
#endif
#ifndef HF_EMINB
#define HF_EMINB 2           // its CTAs per SM (measured at 1.15 M dofs, B = 16: 394 us per iteration with 512 x 1, 346 with 256 x 2, 355 with 256 x 3)
#endif
#define HF_EPAIRS 1024      // (row, variant) pairs per chunk: R = HF_EPAIRS / B rows
#define HF_ERPT (HF_EPAIRS / HF_ENT)
#define HF_EHPT 4           // halo pairs per thread carried in registers

struct EnsCtrl {
  int done, it, n_active, pad;
  unsigned counter[4];
  int active[HF_EB];
  double thr[HF_EB], rz[HF_EB], alpha[HF_EB], beta[HF_EB], bn[HF_EB];
};

struct EnsState {
  int B = 0, LB = 0;        // tile width (power of two) and its log2
  int nb = 0;               // real variants in the tile (<= B)
  int grid = 0;             // CTAs of the set-up kernels
  int last_iters = 0;
  DevBuf<double> base0, s0;           // [nnz]
  DevBuf<double> ks, coeff;           // [B]
  DevBuf<double> dg, g, u, uprev, x, z, z1, p0, p1, w, w1;   // [Nalloc*B]; dg = diagonal of A_b (0 on Dirichlet rows)
  DevBuf<double> drow;                // [2][rows]: base0_ii and S0_ii per row (0 on Dirichlet rows), d_ib = base0_ii + k_b S0_ii
  bool have_prev = false;
  DevBuf<double> part;                // [4][CTAs][B]
  // patch decomposition for the iteration kernel
  int R = 0, nchunks = 0, halo_max = 0, halo_cap = 0, mcap = 0, nstages = 0, igrid = 0;
  size_t stage_bytes = 0, iter_smem = 0;
  DevBuf<int> halo_ptr, halo_idx;
  DevBuf<int> rowptr_pad;             // CSR row pointers padded to whole chunks + 4
  DevBuf<int> lc_off;                 // [nchunks+1] offsets of the chunks' column blocks (multiples of 8)
  DevBuf<unsigned short> lcol;        // local columns, chunk blocks padded to 16 bytes
  DevBuf<double2> bs;                 // {base0, S0} per CSR slot (16 bytes: any row range is TMA-aligned)
  // recycled initial guess, one basis per variant (same scheme as hf_recycle.cu): slots of [nb] doubles in the
  // [row, variant] layout; W = corrections, AW = D^-1 A_b W, inv[slot, b] = 1 / (w . A_b w)
  int rc_cap = 0, rc_count = 0, rc_nseg = 0;
  DevBuf<double> rc_W, rc_AW, rc_inv, rc_coef, rc_parts, rc_part_nn, rc_d, rc_ad;
  DevBuf<EnsCtrl> ctrl;
  EnsCtrl* h_ctrl = nullptr;          // pinned mirror
  DevBuf<double> hist, stage;
  DevBuf<int> watch;
  cudaGraphExec_t chunk_exec[3] = {nullptr, nullptr, nullptr};
  // on-chip batched kernel (hf_enspatch.cu); oc_ok: planned for this tile
  bool oc_ok = false, oc_uniform = false;
  int oc_grid = 0, oc_mat_cap = 0, oc_s0_cap = 0, oc_halo_cap = 0, oc_nh = 1;
  size_t oc_smem = 0, oc_rows = 0;    // oc_rows: rows of the padded [row, variant] vectors
  DevBuf<double> oc_val, oc_s0;       // base0 (+ k S0 when all variants share k) and S0 in the sliced-ELL order of opA
  DevBuf<int> oc_flag;                // [nslices] 1 = the slice has sample-stiffness entries
  DevBuf<unsigned long long> oc_acc;  // fixed-point accumulators of the grid reductions (zeroed before every solve)
  DevBuf<uint4> oc_qpk;               // [2][rows * B] w = D^-1 A p exchange packets
  DevBuf<int> oc_iters, oc_fail;      // per-step iteration counts, failed solves
  DevBuf<unsigned> oc_gen;            // [2] packet generation per half-tile, carried from launch to launch
  DevBuf<double> oc_u0;               // [2][rows * B] u, uprev at the start of hf_ens_run (repeat on the streaming kernels)
  ~EnsState() {
    for (auto& g : chunk_exec)
      if (g) cudaGraphExecDestroy(g);
    if (h_ctrl) cudaFreeHost(h_ctrl);
  }
};


int hf_ens_oc_plan(hf_ctx* c, EnsState* e);
int hf_ens_oc_solve_async(hf_ctx* c, EnsState* e, int step_slot);
