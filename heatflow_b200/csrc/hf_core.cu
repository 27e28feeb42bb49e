// heatflow_b200 - context, mesh/dof-map/sparsity, P1 assembly, time step, C-ABI (sm_100a).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "hf_ctx.cuh"

int hf_pcg_prepare(hf_ctx* c);
int hf_pcg_prepare_from_r(hf_ctx* c);

thread_local std::string hf_err_msg;
int hf_fail(int code, const std::string& msg) {
  hf_err_msg = msg;
  return code;
}

extern "C" int hf_version(void) { return 100; }
extern "C" const char* hf_last_error(void) { return hf_err_msg.c_str(); }
extern "C" int hf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" hf_ctx* hf_create(int device) {
  int n = hf_device_count();
  if (n <= 0) {
    hf_fail(HF_ERR_CUDA, "hf_create: no CUDA device available (there is no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    hf_fail(HF_ERR_ARG, "hf_create: device index out of range");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    hf_fail(HF_ERR_CUDA, "hf_create: cudaSetDevice failed");
    return nullptr;
  }
  hf_ctx* c = new hf_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    hf_fail(HF_ERR_CUDA, "hf_create: cudaStreamCreate/cudaEventCreate failed");
    delete c;
    return nullptr;
  }
  return c;
}

extern "C" void hf_destroy(hf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hf_ens_free(c);
  if (c->ws.h_ctrl) cudaFreeHost(c->ws.h_ctrl);
  c->opA.drop_graphs();
  c->opMr.drop_graphs();
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  cudaStream_t s = c->stream;
  delete c;
  cudaStreamDestroy(s);
}

// Position of (x, y) in [0, 65535]^2 along the Hilbert curve.
static uint64_t hilbert_key(uint32_t x, uint32_t y) {
  const uint32_t n = 65536u;
  uint64_t d = 0;
  for (uint32_t s = n / 2; s > 0; s /= 2) {
    const uint32_t rx = (x & s) ? 1u : 0u, ry = (y & s) ? 1u : 0u;
    d += (uint64_t)s * s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) {
        x = n - 1 - x;
        y = n - 1 - y;
      }
      std::swap(x, y);
    }
  }
  return d;
}

extern "C" int hf_set_ordering(hf_ctx* c, int32_t ordering) {
  if (!c || ordering < 0 || ordering > 2) return hf_fail(HF_ERR_ARG, "hf_set_ordering: ordering must be 0 (auto), 1 (as given) or 2 (Hilbert)");
  c->ordering_req = ordering;
  return HF_OK;
}

__global__ void k_gather_nodal(int n, int ncomp, const int* __restrict__ idx, const double* __restrict__ src,
                               double* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n * ncomp) {
    const int i = t / ncomp, a = t - i * ncomp;
    dst[t] = src[(size_t)idx[i] * ncomp + a];
  }
}

// host array in the caller's node numbering -> device array in internal numbering
int hf_upload_nodal(hf_ctx* c, const double* h_user, double* d_internal) {
  if (!c->permuted) {
    HF_CUDA(cudaMemcpyAsync(d_internal, h_user, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
  } else {
    HF_CUDA(cudaMemcpyAsync(c->stage.p, h_user, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
    k_gather_nodal<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, 1, c->order_d.p, c->stage.p, d_internal);
    HF_CUDA(cudaGetLastError());
  }
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

// device array [N, ncomp] in internal numbering -> host array in the caller's numbering (ncomp <= 2)
int hf_download_nodal(hf_ctx* c, const double* d_internal, double* h_user, int ncomp) {
  const double* src = d_internal;
  if (c->permuted) {
    k_gather_nodal<<<(c->N * ncomp + 255) / 256, 256, 0, c->stream>>>(c->N, ncomp, c->rank_d.p, d_internal, c->stage.p);
    HF_CUDA(cudaGetLastError());
    src = c->stage.p;
  }
  HF_CUDA(cudaMemcpyAsync(h_user, src, sizeof(double) * c->N * ncomp, cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

int hf_internal_nodes(hf_ctx* c, int n, const int32_t* user_nodes, std::vector<int>& out) {
  out.resize(n);
  for (int i = 0; i < n; ++i) {
    if (user_nodes[i] < 0 || user_nodes[i] >= c->N) return hf_fail(HF_ERR_ARG, "node index out of range");
    out[i] = c->permuted ? c->h_rank[user_nodes[i]] : user_nodes[i];
  }
  return HF_OK;
}

// ---------------------------------------------------------------------------------------
// mesh -> dof map, node->cell adjacency, CSR sparsity pattern (host, integer work, one-off)
// ---------------------------------------------------------------------------------------
extern "C" int hf_set_mesh(hf_ctx* c, int32_t N, int32_t E, int32_t nv, const double* xy_user,
                           const int32_t* cells_user, const int32_t* cell_tag) {
  if (!c || !xy_user || !cells_user || !cell_tag) return hf_fail(HF_ERR_ARG, "hf_set_mesh: null argument");
  if (N <= 0 || E <= 0 || (nv != 2 && nv != 3)) return hf_fail(HF_ERR_ARG, "hf_set_mesh: need N > 0, E > 0, nv in {2,3}");
  for (int64_t k = 0; k < (int64_t)E * nv; ++k)
    if (cells_user[k] < 0 || cells_user[k] >= N) return hf_fail(HF_ERR_ARG, "hf_set_mesh: cell references a node outside [0, N)");
  cudaSetDevice(c->device);
  hf_ens_free(c);
  c->N = N;
  c->E = E;
  c->nv = nv;
  c->Npad = (N + HF_SLICE - 1) / HF_SLICE * HF_SLICE;
  c->op_built = c->proj_built = false;
  c->opA.struct_valid = c->opMr.struct_valid = false;
  c->opA.pp_rpt = c->opMr.pp_rpt = 0;
  c->opMr.plan_from = nullptr;
  // ---- internal numbering.  auto: triangle meshes are sorted along a Hilbert curve so that consecutive
  // rows form compact 2-D patches with short halo lists (patch kernel, streaming kernel, ensembles); the
  // caller's (banded) order is kept on request (hf_set_ordering(ctx, 1): contiguous-range kernel) and for
  // 1-D meshes.
  int ordering = c->ordering_req;
  if (ordering == 0) ordering = (nv == 3) ? 2 : 1;
  if (nv != 3) ordering = 1;
  c->permuted = (ordering == 2);
  c->h_rank.clear();
  c->h_order.clear();
  std::vector<double> xy_store;
  std::vector<int> cells_store;
  const double* xy = xy_user;
  const int32_t* cells = cells_user;
  if (c->permuted) {
    double lo[2] = {xy_user[0], xy_user[1]}, hi[2] = {xy_user[0], xy_user[1]};
    for (int i = 0; i < N; ++i)
      for (int a = 0; a < 2; ++a) {
        lo[a] = std::min(lo[a], xy_user[2 * i + a]);
        hi[a] = std::max(hi[a], xy_user[2 * i + a]);
      }
    const double span = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), 1e-300);
    std::vector<std::pair<uint64_t, int>> keyed(N);
    for (int i = 0; i < N; ++i) {
      const uint32_t qx = (uint32_t)std::min(65535.0, std::max(0.0, (xy_user[2 * i] - lo[0]) / span * 65535.0));
      const uint32_t qy = (uint32_t)std::min(65535.0, std::max(0.0, (xy_user[2 * i + 1] - lo[1]) / span * 65535.0));
      keyed[i] = std::make_pair(hilbert_key(qx, qy), i);
    }
    std::sort(keyed.begin(), keyed.end());
    c->h_order.resize(N);
    c->h_rank.resize(N);
    for (int n = 0; n < N; ++n) {
      c->h_order[n] = keyed[n].second;
      c->h_rank[keyed[n].second] = n;
    }
    // Meshes of at most one 1024-row patch per SM run in the on-chip kernels with four 32-row slices per warp
    // (hf_patch.cu, rpt = 4).  There the rows whose q value a neighbouring patch needs ("boundary" rows: a cell joins
    // them to a node of another patch, ~250 of 1024) are moved into the FIRST slice of every warp, the interior rows
    // into the other three: a warp publishes its boundary packets after a quarter of its SpMV and computes the interior
    // rows while they travel through L2, instead of publishing at the end and waiting a full trip for the neighbours'.
    // The patches themselves (sets of rows, halo lists) do not change; inside a class the Hilbert order is kept.
    // (the same for 1536- and 2048-row patches, six / eight slices per warp, on meshes of up to one such patch per SM)
    const int spw_bf = ((int64_t)N + 1023) / 1024 <= c->sm_count ? 4 : ((int64_t)N + 1535) / 1536 <= c->sm_count ? 6
                       : ((int64_t)N + 2047) / 2048 <= c->sm_count ? 8 : 0;
    if (nv == 3 && !getenv("HF_NO_BOUNDARY_FIRST") && spw_bf) {
      const int SPW = spw_bf, PB = 256 * spw_bf;               // slices per warp, rows per patch
      std::vector<unsigned char> bnd(N, 0);
      for (int e = 0; e < E; ++e) {
        const int a = c->h_rank[cells_user[3 * e]] / PB, b = c->h_rank[cells_user[3 * e + 1]] / PB,
                  d = c->h_rank[cells_user[3 * e + 2]] / PB;
        if (a != b || a != d)
          for (int v = 0; v < 3; ++v) bnd[cells_user[3 * e + v]] = 1;   // conservative: every node of a straddling cell
      }
      std::vector<int> fresh(N);
      for (int lo = 0; lo < N; lo += PB) {
        const int cnt = std::min(PB, N - lo), nsl = (cnt + 31) / 32;
        std::vector<int> slots;                                // positions of the block in filling order: slice class 0, 1, 2, 3
        slots.reserve(cnt);
        for (int cls = 0; cls < SPW; ++cls)
          for (int sl = cls; sl < nsl; sl += SPW)
            for (int l = 0; l < 32 && sl * 32 + l < cnt; ++l) slots.push_back(lo + sl * 32 + l);
        int at = 0;
        for (int pass = 1; pass >= 0; --pass)                  // boundary rows first, then interior rows, each in Hilbert order
          for (int pos = lo; pos < lo + cnt; ++pos)
            if ((int)bnd[c->h_order[pos]] == pass) fresh[slots[at++]] = c->h_order[pos];
      }
      c->h_order.swap(fresh);
      for (int n = 0; n < N; ++n) c->h_rank[c->h_order[n]] = n;
    }
    xy_store.resize((size_t)N * 2);
    for (int n = 0; n < N; ++n) {
      xy_store[2 * n] = xy_user[2 * c->h_order[n]];
      xy_store[2 * n + 1] = xy_user[2 * c->h_order[n] + 1];
    }
    cells_store.resize((size_t)E * nv);
    for (int64_t k = 0; k < (int64_t)E * nv; ++k) cells_store[k] = c->h_rank[cells_user[k]];
    xy = xy_store.data();
    cells = cells_store.data();
    HF_TRY(c->rank_d.upload(c->h_rank.data(), N, c->stream));
    HF_TRY(c->order_d.upload(c->h_order.data(), N, c->stream));
  }
  HF_TRY(c->stage.alloc((size_t)2 * N, c->stream));
  // node -> cells, ascending cell index (fixes the summation order of the gather assembly)
  std::vector<int> ptr(N + 1, 0);
  for (int64_t k = 0; k < (int64_t)E * nv; ++k) ptr[cells[k] + 1]++;
  for (int i = 0; i < N; ++i) ptr[i + 1] += ptr[i];
  std::vector<int> idx((size_t)E * nv), fill(ptr.begin(), ptr.end() - 1);
  for (int e = 0; e < E; ++e)
    for (int a = 0; a < nv; ++a) idx[fill[cells[(size_t)e * nv + a]]++] = e;
  // pattern: neighbours through shared cells, plus the diagonal, sorted
  c->h_rowptr.assign(N + 1, 0);
  c->h_col.clear();
  c->h_col.reserve((size_t)N * 8);
  std::vector<int> tmp;
  for (int i = 0; i < N; ++i) {
    tmp.clear();
    for (int a = ptr[i]; a < ptr[i + 1]; ++a)
      for (int b = 0; b < nv; ++b) tmp.push_back(cells[(size_t)idx[a] * nv + b]);
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    c->h_col.insert(c->h_col.end(), tmp.begin(), tmp.end());
    if (c->h_col.size() > (size_t)INT32_MAX) return hf_fail(HF_ERR_ARG, "hf_set_mesh: nnz exceeds int32");
    c->h_rowptr[i + 1] = (int)c->h_col.size();
  }
  c->nnz = (int64_t)c->h_col.size();
  HF_TRY(c->xy.upload(xy, (size_t)N * 2, c->stream));
  HF_TRY(c->cells.upload(cells, (size_t)E * nv, c->stream));
  HF_TRY(c->cell_tag.upload(cell_tag, E, c->stream));
  HF_TRY(c->n2c_ptr.upload(ptr.data(), N + 1, c->stream));
  HF_TRY(c->n2c_idx.upload(idx.data(), idx.size(), c->stream));
  HF_TRY(c->rowptr.upload(c->h_rowptr.data(), N + 1, c->stream));
  HF_TRY(c->col.upload(c->h_col.data(), c->h_col.size(), c->stream));
  HF_TRY(c->cm.alloc(E, c->stream));
  HF_TRY(c->ck.alloc(E, c->stream));
  HF_TRY(c->bcflag.alloc(c->Npad, c->stream));
  HF_TRY(c->gfull.alloc(c->Npad, c->stream));
  HF_TRY(c->u.alloc(c->Npad, c->stream));
  HF_TRY(c->uprev.alloc(c->Npad, c->stream));
  HF_TRY(c->b.alloc(c->Npad, c->stream));
  HF_TRY(c->source.alloc(c->Npad, c->stream));
  c->have_prev = c->have_source = false;
  c->n_bc = c->n_gauss = 0;
  HF_TRY(hf_pcg_alloc(c));
  HF_CUDA(cudaStreamSynchronize(c->stream));   // the staging vectors above go out of scope here
  return HF_OK;
}

extern "C" int hf_set_materials(hf_ctx* c, int32_t n, const int32_t* tags, const double* kappa, const double* rho_c) {
  if (!c || n <= 0 || !tags || !kappa || !rho_c) return hf_fail(HF_ERR_ARG, "hf_set_materials: bad arguments");
  c->mat_tags.assign(tags, tags + n);
  c->mat_kappa.assign(kappa, kappa + n);
  c->mat_rhoc.assign(rho_c, rho_c + n);
  c->op_built = false;
  return HF_OK;
}

__global__ void k_scatter_bc(int n, const int* __restrict__ dofs, const double* __restrict__ val,
                             unsigned char* __restrict__ flag, double* __restrict__ g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (flag) flag[dofs[i]] = 1;
    g[dofs[i]] = val[i];
  }
}

extern "C" int hf_set_bcs(hf_ctx* c, int32_t n_bc, const int32_t* bc_dofs, const double* bc_value,
                          int32_t n_gauss, const int32_t* gauss_slot, const double* gauss_r) {
  if (!c || c->N == 0) return hf_fail(HF_ERR_STATE, "hf_set_bcs: call hf_set_mesh first");
  if (n_bc < 0 || n_gauss < 0 || (n_bc && (!bc_dofs || !bc_value)) || (n_gauss && (!gauss_slot || !gauss_r)))
    return hf_fail(HF_ERR_ARG, "hf_set_bcs: bad arguments");
  for (int i = 0; i < n_bc; ++i) {
    if (bc_dofs[i] < 0 || bc_dofs[i] >= c->N) return hf_fail(HF_ERR_ARG, "hf_set_bcs: dof out of range");
    if (i && bc_dofs[i] <= bc_dofs[i - 1]) return hf_fail(HF_ERR_ARG, "hf_set_bcs: bc_dofs must be sorted and unique");
  }
  std::vector<int> bd, gd(n_gauss);   // internal numbering (slot order unchanged)
  HF_TRY(hf_internal_nodes(c, n_bc, bc_dofs, bd));
  for (int i = 0; i < n_gauss; ++i) {
    if (gauss_slot[i] < 0 || gauss_slot[i] >= n_bc) return hf_fail(HF_ERR_ARG, "hf_set_bcs: gauss_slot out of range");
    gd[i] = bd[gauss_slot[i]];
  }
  cudaSetDevice(c->device);
  hf_ens_free(c);
  c->n_bc = n_bc;
  c->n_gauss = n_gauss;
  HF_CUDA(cudaMemsetAsync(c->bcflag.p, 0, c->Npad, c->stream));
  HF_CUDA(cudaMemsetAsync(c->gfull.p, 0, sizeof(double) * c->Npad, c->stream));
  HF_TRY(c->bc_dofs.upload(bd.data(), n_bc, c->stream));
  HF_TRY(c->gauss_dof.upload(gd.data(), n_gauss, c->stream));
  HF_TRY(c->gauss_r.upload(gauss_r, n_gauss, c->stream));
  if (n_bc) {
    DevBuf<double> v;
    HF_TRY(v.upload(bc_value, n_bc, c->stream));
    k_scatter_bc<<<(n_bc + 255) / 256, 256, 0, c->stream>>>(n_bc, c->bc_dofs.p, v.p, c->bcflag.p, c->gfull.p);
    HF_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->op_built = false;
  return HF_OK;
}

extern "C" int hf_set_bc_values(hf_ctx* c, const double* bc_value) {
  if (!c || !bc_value) return hf_fail(HF_ERR_ARG, "hf_set_bc_values: null argument");
  if (c->n_bc == 0) return HF_OK;
  cudaSetDevice(c->device);
  DevBuf<double> v;
  HF_TRY(v.upload(bc_value, c->n_bc, c->stream));
  k_scatter_bc<<<(c->n_bc + 255) / 256, 256, 0, c->stream>>>(c->n_bc, c->bc_dofs.p, v.p, nullptr, c->gfull.p);
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

// ---------------------------------------------------------------------------------------
// P1 element matrices, gather assembly (one thread per matrix row, ascending cell order)
// ---------------------------------------------------------------------------------------
// Closed forms (exact for affine r; reference forms: run_with_diamond.py:328-335):
//   axisymmetric  M_ii = |T|(r_i/10 + (r_j+r_k)/30)   M_ij = |T|((r_i+r_j)/30 + r_k/60)
//                 K_ij = |T| rbar grad phi_i . grad phi_j
//   planar        M = |T|/12 (1 + delta_ij)           K_ij = |T| grad phi_i . grad phi_j
//   interval      M = h/6 (1 + delta_ij)              K = +-1/h          (run_no_diamond_1d.py:537-540)
__global__ void __launch_bounds__(256)
k_assemble(int N, int nv, int axisym, const double* __restrict__ xy, const int* __restrict__ cells,
           const double* __restrict__ cm, const double* __restrict__ ck,
           const int* __restrict__ n2c_ptr, const int* __restrict__ n2c_idx,
           const int* __restrict__ rowptr, const int* __restrict__ col, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int r0 = rowptr[i], r1 = rowptr[i + 1];
  for (int k = r0; k < r1; ++k) out[k] = 0.0;
  for (int a = n2c_ptr[i]; a < n2c_ptr[i + 1]; ++a) {
    const int e = n2c_idx[a];
    const double m_e = cm[e], k_e = ck[e];
    int v[3];
    double em[3];
    if (nv == 3) {
      v[0] = cells[3 * e];
      v[1] = cells[3 * e + 1];
      v[2] = cells[3 * e + 2];
      const int li = (v[0] == i) ? 0 : (v[1] == i) ? 1 : 2;
      double z[3], r[3];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        z[t] = xy[2 * v[t]];
        r[t] = xy[2 * v[t] + 1];
      }
      const double bb[3] = {r[1] - r[2], r[2] - r[0], r[0] - r[1]};
      const double cc[3] = {z[2] - z[1], z[0] - z[2], z[1] - z[0]};
      const double det = bb[0] * cc[1] - bb[1] * cc[0];
      const double area = 0.5 * fabs(det);
      const double rsum = r[0] + r[1] + r[2];
      const double kw = axisym ? area * rsum / 3.0 : area;
#pragma unroll
      for (int lj = 0; lj < 3; ++lj) {
        double m;
        if (axisym) {
          if (lj == li) m = area * (r[li] / 10.0 + (rsum - r[li]) / 30.0);
          else m = area * ((r[li] + r[lj]) / 30.0 + r[3 - li - lj] / 60.0);
        } else {
          m = (lj == li) ? area / 6.0 : area / 12.0;
        }
        const double g = (bb[li] / det) * (bb[lj] / det) + (cc[li] / det) * (cc[lj] / det);
        em[lj] = m_e * m + k_e * (kw * g);
      }
    } else {
      v[0] = cells[2 * e];
      v[1] = cells[2 * e + 1];
      v[2] = -1;
      const int li = (v[0] == i) ? 0 : 1;
      const double h = fabs(xy[2 * v[1]] - xy[2 * v[0]]);
#pragma unroll
      for (int lj = 0; lj < 2; ++lj)
        em[lj] = m_e * (h / 6.0) * ((lj == li) ? 2.0 : 1.0) + k_e * (1.0 / h) * ((lj == li) ? 1.0 : -1.0);
      em[2] = 0.0;
    }
    for (int lj = 0; lj < nv; ++lj) {
      const int j = v[lj];
      int k = r0;
      while (col[k] != j) ++k;   // j is in the pattern by construction
      out[k] += em[lj];
    }
  }
}

int hf_assemble_values(hf_ctx* c, const double* cm, const double* ck, int axisym, double* out) {
  k_assemble<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, c->nv, axisym, c->xy.p, c->cells.p, cm, ck,
                                                        c->n2c_ptr.p, c->n2c_idx.p, c->rowptr.p, c->col.p, out);
  HF_CUDA(cudaGetLastError());
  return HF_OK;
}

// per-cell coefficients from the tag tables: cm = fm * rho_c(tag), ck = fk * kappa(tag)
__global__ void k_cell_coef(int E, const int* __restrict__ tag, int nt, const int* __restrict__ tags,
                            const double* __restrict__ kappa, const double* __restrict__ rhoc,
                            double fm, double fk, double* __restrict__ cm, double* __restrict__ ck,
                            int* __restrict__ missing) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int t = tag[e];
  int f = -1;
  for (int k = 0; k < nt; ++k)
    if (tags[k] == t) f = k;
  if (f < 0) {
    atomicExch(missing, t + 1);
    cm[e] = 0.0;
    ck[e] = 0.0;
  } else {
    cm[e] = fm * rhoc[f];
    ck[e] = fk * kappa[f];
  }
}

__global__ void k_fill(int n, double v, double* __restrict__ p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------
// Dirichlet treatment + Jacobi scaling + sliced-ELL conversion
// ---------------------------------------------------------------------------------------
__global__ void k_diag_scale(int N, int Npad, const int* __restrict__ rowptr, const int* __restrict__ col,
                             const double* __restrict__ val, const unsigned char* __restrict__ bcflag,
                             int apply_bc, double* __restrict__ scale, int* __restrict__ bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  double s = 1.0;
  if (i < N && !(apply_bc && bcflag[i])) {
    double d = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) d = val[k];
    if (!(d > 0.0)) {
      atomicExch(bad, i + 1);
      d = 1.0;
    }
    s = sqrt(d);
  }
  scale[i] = s;
}

__global__ void k_fill_sell(int N, int Npad, int R, const int* __restrict__ rowptr, const int* __restrict__ col,
                            const double* __restrict__ val, const unsigned char* __restrict__ bcflag,
                            int apply_bc, const double* __restrict__ scale, const int* __restrict__ slice_ptr,
                            const unsigned short* __restrict__ lcol_csr, int* __restrict__ scol,
                            unsigned short* __restrict__ slcol, double* __restrict__ sval, double* __restrict__ val_bc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const int s = i / HF_SLICE, lane = i % HF_SLICE;
  const int base = slice_ptr[s];
  const int w = (slice_ptr[s + 1] - base) / HF_SLICE;
  int len = 0, r0 = 0;
  if (i < N) {
    r0 = rowptr[i];
    len = rowptr[i + 1] - r0;
  }
  const bool bci = apply_bc && i < N && bcflag[i];
  const double si = scale[i];
  for (int k = 0; k < w; ++k) {
    int cj = i;
    unsigned short lc = (unsigned short)(i % R);   // padding entries point at the row itself, value 0
    double v = 0.0;
    if (k < len) {
      cj = col[r0 + k];
      lc = lcol_csr[r0 + k];
      double a = val[r0 + k];
      if (apply_bc && (bci || bcflag[cj])) a = (cj == i) ? 1.0 : 0.0;
      if (val_bc) val_bc[r0 + k] = a;
      v = a / (si * scale[cj]);
    }
    scol[base + k * HF_SLICE + lane] = cj;
    slcol[base + k * HF_SLICE + lane] = lc;
    sval[base + k * HF_SLICE + lane] = v;
  }
}

// Chunks of R consecutive rows -> halo lists (columns outside the chunk, ascending) and 16-bit
// local column indices in CSR slot order.  Host, integer work, once per mesh and R.
int hf_build_patches(const hf_ctx* c, int R, std::vector<int>& halo_ptr, std::vector<int>& halo_idx,
                     std::vector<unsigned short>& lcol, int* halo_max) {
  const int N = c->N;
  const int nchunks = (c->Npad + R - 1) / R;
  halo_ptr.assign(nchunks + 1, 0);
  halo_idx.clear();
  lcol.resize(c->nnz);
  std::vector<int> slot(N, -1), touched;
  *halo_max = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int lo = ch * R, hi = std::min(lo + R, N);
    touched.clear();
    for (int i = lo; i < hi; ++i)
      for (int k = c->h_rowptr[i]; k < c->h_rowptr[i + 1]; ++k) {
        const int j = c->h_col[k];
        if ((j < lo || j >= lo + R) && slot[j] < 0) {
          slot[j] = 0;
          touched.push_back(j);
        }
      }
    std::sort(touched.begin(), touched.end());
    if (R + (int)touched.size() > 65535) return hf_fail(HF_ERR_ARG, "patch halo exceeds the 16-bit local column range");
    for (size_t h = 0; h < touched.size(); ++h) slot[touched[h]] = (int)h;
    for (int i = lo; i < hi; ++i)
      for (int k = c->h_rowptr[i]; k < c->h_rowptr[i + 1]; ++k) {
        const int j = c->h_col[k];
        lcol[k] = (unsigned short)((j >= lo && j < lo + R) ? (j - lo) : (R + slot[j]));
      }
    for (int j : touched) slot[j] = -1;
    halo_idx.insert(halo_idx.end(), touched.begin(), touched.end());
    halo_ptr[ch + 1] = (int)halo_idx.size();
    *halo_max = std::max(*halo_max, (int)touched.size());
  }
  return HF_OK;
}

// values of an operator whose structure (slices, patches, local columns, plans) is already on the device
static int sell_fill_values(hf_ctx* c, const DevBuf<double>& csr_val, bool apply_bc, SellOp& op, DevBuf<double>* val_bc_out) {
  if (val_bc_out && val_bc_out->n != (size_t)c->nnz) HF_TRY(val_bc_out->alloc(c->nnz, c->stream));
  // scratch kept in the context: a temporary released with cudaFree would synchronise the whole device and stall on
  // the other context's queued solves when two simulations share the GPU (hf_set_sharing)
  DevBuf<int>& bad = c->sc_flag;
  if (bad.n != 1) HF_TRY(bad.alloc(1, c->stream));
  HF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c->stream));
  const int g = (c->Npad + 255) / 256;
  k_diag_scale<<<g, 256, 0, c->stream>>>(c->N, c->Npad, c->rowptr.p, c->col.p, csr_val.p, c->bcflag.p, apply_bc ? 1 : 0,
                                         op.scale.p, bad.p);
  k_fill_sell<<<g, 256, 0, c->stream>>>(c->N, c->Npad, op.R, c->rowptr.p, c->col.p, csr_val.p, c->bcflag.p, apply_bc ? 1 : 0,
                                        op.scale.p, op.slice_ptr.p, op.lcol_csr.p, op.col.p, op.lcol.p, op.val.p,
                                        val_bc_out ? val_bc_out->p : nullptr);
  HF_CUDA(cudaGetLastError());
  int hbad = 0;
  HF_TRY(bad.download(&hbad, 1, c->stream));
  if (hbad) return hf_fail(HF_ERR_STATE, "operator has a non-positive diagonal at row " + std::to_string(hbad - 1) +
                                             " (unset material or degenerate cell?)");
  return HF_OK;
}

int hf_build_sell(hf_ctx* c, const DevBuf<double>& csr_val, bool apply_bc, SellOp& op, DevBuf<double>* val_bc_out) {
  // sweeps re-assemble with other material values on the same mesh: the structure is kept, only values move
  if (op.struct_valid) return sell_fill_values(c, csr_val, apply_bc, op, val_bc_out);
  const int nsl = c->Npad / HF_SLICE;
  std::vector<int> sp(nsl + 1, 0);
  for (int s = 0; s < nsl; ++s) {
    int w = 0;
    for (int i = s * HF_SLICE; i < std::min((s + 1) * HF_SLICE, c->N); ++i) w = std::max(w, c->h_rowptr[i + 1] - c->h_rowptr[i]);
    int64_t next = (int64_t)sp[s] + (int64_t)w * HF_SLICE;
    if (next > INT32_MAX) return hf_fail(HF_ERR_ARG, "sliced-ELL storage exceeds int32 offsets");
    sp[s + 1] = (int)next;
  }
  op.drop_graphs();
  op.nslices = nsl;
  op.padded_nnz = (size_t)sp[nsl];
  // chunk size of the streaming kernel (one persistent CTA per SM walks the chunks): 512 rows when every
  // CTA still gets >= 8 chunks, else 256
  op.R = (c->Npad >= 8 * c->sm_count * 512) ? 512 : 256;
  if (HF_IT < 512) op.R = 256;                     // one row per thread in phase 1
  if (const char* env = getenv("HF_CHUNK_R")) {   // tuning knob
    const int r = atoi(env);
    if (r == 256 || (r == 512 && HF_IT >= 512)) op.R = r;
  }
  op.nchunks = (c->Npad + op.R - 1) / op.R;
  op.mat_cap = 0;
  for (int ch = 0; ch < op.nchunks; ++ch) {
    const int s0 = ch * (op.R / HF_SLICE), s1 = std::min(s0 + op.R / HF_SLICE, nsl);
    op.mat_cap = std::max(op.mat_cap, sp[s1] - sp[s0]);
  }
  std::vector<int> hptr, hidx;
  std::vector<unsigned short> lcol;
  HF_TRY(hf_build_patches(c, op.R, hptr, hidx, lcol, &op.halo_max));
  // shared-memory stage: operator block (values + 16-bit columns) + x r q p (own rows) + halo p
  op.halo_cap = (op.halo_max + 1) & ~1;
  op.stage_bytes = (((size_t)op.mat_cap * 10 + sizeof(double) * ((size_t)4 * op.R + op.halo_cap) + 4 * (op.R / HF_SLICE + 4)) + 127) &
                   ~(size_t)127;
  {
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device);
    max_smem = max_smem / HF_IT_MINB - 1024;         // HF_IT_MINB co-resident CTAs share the SM's shared memory
    op.nstages = (int)std::min<size_t>(4, ((size_t)max_smem - 2048) / op.stage_bytes);
    if (const char* env = getenv("HF_STAGES")) op.nstages = std::max(1, std::min(op.nstages, atoi(env)));
    if (op.nstages < 1)
      return hf_fail(HF_ERR_STATE, "streaming PCG chunk does not fit in shared memory (halo of " + std::to_string(op.halo_max) +
                                       " rows); use hf_set_ordering(ctx, 2)");
    op.iter_smem = op.stage_bytes * op.nstages;
  }
  HF_TRY(op.lcol_csr.upload(lcol.data(), lcol.size(), c->stream));
  HF_TRY(op.halo_ptr.upload(hptr.data(), hptr.size(), c->stream));
  if (hidx.empty()) hidx.push_back(0);
  HF_TRY(op.halo_idx.upload(hidx.data(), hidx.size(), c->stream));
  const int sp_end = sp[nsl];
  sp.resize((size_t)op.nchunks * (op.R / HF_SLICE) + 5, sp_end);   // whole chunks + 4: zero-width padding slices
  HF_TRY(op.slice_ptr.upload(sp.data(), sp.size(), c->stream));
  HF_TRY(op.col.alloc(op.padded_nnz, c->stream));
  HF_TRY(op.lcol.alloc(op.padded_nnz, c->stream));
  HF_TRY(op.val.alloc(op.padded_nnz, c->stream));
  HF_TRY(op.scale.alloc(c->Npad, c->stream));
  if (c->ws.parts.n < (size_t)16 * c->sm_count) HF_TRY(c->ws.parts.alloc((size_t)16 * c->sm_count, c->stream));
  HF_TRY(sell_fill_values(c, csr_val, apply_bc, op, val_bc_out));
  op.struct_valid = true;
  return HF_OK;
}

static int cell_coefs(hf_ctx* c, double fm, double fk) {
  const int nt = (int)c->mat_tags.size();
  if (nt == 0) return hf_fail(HF_ERR_STATE, "call hf_set_materials first");
  DevBuf<int>&tags = c->sc_tags, &miss = c->sc_flag;      // context scratch: no cudaFree (device-wide sync) per call
  DevBuf<double>&kap = c->sc_kap, &rc = c->sc_rc;
  HF_TRY(tags.upload(c->mat_tags.data(), nt, c->stream));
  HF_TRY(kap.upload(c->mat_kappa.data(), nt, c->stream));
  HF_TRY(rc.upload(c->mat_rhoc.data(), nt, c->stream));
  if (miss.n != 1) HF_TRY(miss.alloc(1, c->stream));
  HF_CUDA(cudaMemsetAsync(miss.p, 0, sizeof(int), c->stream));
  k_cell_coef<<<(c->E + 255) / 256, 256, 0, c->stream>>>(c->E, c->cell_tag.p, nt, tags.p, kap.p, rc.p, fm, fk, c->cm.p,
                                                         c->ck.p, miss.p);
  int hm = 0;
  HF_TRY(miss.download(&hm, 1, c->stream));
  if (hm) return hf_fail(HF_ERR_ARG, "cell tag " + std::to_string(hm - 1) + " has no material");
  return HF_OK;
}

extern "C" int hf_build_operator(hf_ctx* c, double dt, int32_t axisymmetric) {
  if (!c || c->N == 0) return hf_fail(HF_ERR_STATE, "hf_build_operator: call hf_set_mesh first");
  if (!(dt > 0.0)) return hf_fail(HF_ERR_ARG, "hf_build_operator: dt must be positive");
  cudaSetDevice(c->device);
  c->dt = dt;
  c->axisym = (c->nv == 3) ? (axisymmetric ? 1 : 0) : 0;
  if (c->valM.n != (size_t)c->nnz) HF_TRY(c->valM.alloc(c->nnz, c->stream));
  if (c->valA0.n != (size_t)c->nnz) HF_TRY(c->valA0.alloc(c->nnz, c->stream));
  HF_TRY(cell_coefs(c, 1.0, 0.0));
  HF_TRY(hf_assemble_values(c, c->cm.p, c->ck.p, c->axisym, c->valM.p));
  HF_TRY(cell_coefs(c, 1.0, dt));
  HF_TRY(hf_assemble_values(c, c->cm.p, c->ck.p, c->axisym, c->valA0.p));
  const bool new_structure = !c->opA.struct_valid;
  HF_TRY(hf_build_sell(c, c->valA0, true, c->opA, &c->valA));
  if (new_structure) {
    HF_TRY(hf_patch_plan(c, c->opA));
  }
  c->valM1.release();
  // the pipelined on-chip kernel (k_pcg_pipe) trades one reduction trip for a longer recurrence chain: fine for the
  // mass-dominated transient operator (measured 1e-12 .. 1.5e-11 against the LU oracle), not for a pure stiffness
  // operator (steady state, cond ~ 1e8: its recurrence residual drifts to ~1e-5 of the true one) - those keep classic CG
  c->op_transient = !c->mat_rhoc.empty();
  for (double rc : c->mat_rhoc) c->op_transient = c->op_transient && rc > 0.0;
  c->op_built = true;
  c->proj_built = false;
  c->last_iters = 0;
  hf_rc_reset(c);
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

extern "C" int hf_get_sizes(hf_ctx* c, int32_t* n_nodes, int64_t* nnz) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  if (n_nodes) *n_nodes = c->N;
  if (nnz) *nnz = c->nnz;
  return HF_OK;
}

// CSR arrays in the caller's node numbering.  With an internal permutation the pattern is mapped
// back row by row (columns re-sorted) and `slot` records where each entry lives internally.
static void user_pattern(const hf_ctx* c, std::vector<int>& rowptr, std::vector<int>& col, std::vector<int>& slot) {
  const int N = c->N;
  rowptr.assign(N + 1, 0);
  col.resize(c->nnz);
  slot.resize(c->nnz);
  for (int i = 0; i < N; ++i) {
    const int ri = c->h_rank[i];
    rowptr[i + 1] = rowptr[i] + (c->h_rowptr[ri + 1] - c->h_rowptr[ri]);
  }
  std::vector<std::pair<int, int>> tmp;
  for (int i = 0; i < N; ++i) {
    const int ri = c->h_rank[i];
    tmp.clear();
    for (int k = c->h_rowptr[ri]; k < c->h_rowptr[ri + 1]; ++k) tmp.emplace_back(c->h_order[c->h_col[k]], k);
    std::sort(tmp.begin(), tmp.end());
    for (size_t t = 0; t < tmp.size(); ++t) {
      col[rowptr[i] + t] = tmp[t].first;
      slot[rowptr[i] + t] = tmp[t].second;
    }
  }
}

extern "C" int hf_get_csr(hf_ctx* c, int32_t* rowptr, int32_t* col, double* val_A, double* val_M, double* val_A0) {
  if (!c || c->N == 0) return hf_fail(HF_ERR_STATE, "hf_get_csr: no mesh");
  cudaSetDevice(c->device);
  if ((val_A || val_M || val_A0) && !c->op_built) return hf_fail(HF_ERR_STATE, "hf_get_csr: operator not built");
  if (!c->permuted) {
    if (rowptr) std::copy(c->h_rowptr.begin(), c->h_rowptr.end(), rowptr);
    if (col) std::copy(c->h_col.begin(), c->h_col.end(), col);
    if (val_A) HF_TRY(c->valA.download(val_A, c->nnz, c->stream));
    if (val_M) HF_TRY(c->valM.download(val_M, c->nnz, c->stream));
    if (val_A0) HF_TRY(c->valA0.download(val_A0, c->nnz, c->stream));
    return HF_OK;
  }
  std::vector<int> rp, cl, slot;
  user_pattern(c, rp, cl, slot);
  if (rowptr) std::copy(rp.begin(), rp.end(), rowptr);
  if (col) std::copy(cl.begin(), cl.end(), col);
  std::vector<double> tmp(c->nnz);
  const DevBuf<double>* src[3] = {&c->valA, &c->valM, &c->valA0};
  double* dst[3] = {val_A, val_M, val_A0};
  for (int a = 0; a < 3; ++a) {
    if (!dst[a]) continue;
    HF_TRY(src[a]->download(tmp.data(), c->nnz, c->stream));
    for (int64_t k = 0; k < c->nnz; ++k) dst[a][k] = tmp[slot[k]];
  }
  return HF_OK;
}

// ---------------------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------------------
extern "C" int hf_set_state(hf_ctx* c, const double* u) {
  if (!c || !u || c->N == 0) return hf_fail(HF_ERR_ARG, "hf_set_state: bad arguments");
  cudaSetDevice(c->device);
  HF_TRY(hf_upload_nodal(c, u, c->u.p));
  c->have_prev = false;
  hf_rc_reset(c);
  return HF_OK;
}

extern "C" int hf_get_state(hf_ctx* c, double* u) {
  if (!c || !u || c->N == 0) return hf_fail(HF_ERR_ARG, "hf_get_state: bad arguments");
  cudaSetDevice(c->device);
  return hf_download_nodal(c, c->u.p, u, 1);
}

extern "C" int hf_get_rhs(hf_ctx* c, double* b) {
  if (!c || !b || c->N == 0) return hf_fail(HF_ERR_ARG, "hf_get_rhs: bad arguments");
  cudaSetDevice(c->device);
  return hf_download_nodal(c, c->b.p, b, 1);
}

extern "C" int hf_set_source(hf_ctx* c, const double* s) {
  if (!c || c->N == 0) return hf_fail(HF_ERR_ARG, "hf_set_source: bad arguments");
  cudaSetDevice(c->device);
  if (!s) {
    c->have_source = false;
    return HF_OK;
  }
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_set_source: operator not built");
  if (c->valM1.n != (size_t)c->nnz) {   // unit-coefficient mass of the linear form dt * s * v
    HF_TRY(c->valM1.alloc(c->nnz, c->stream));
    k_fill<<<(c->E + 255) / 256, 256, 0, c->stream>>>(c->E, 1.0, c->cm.p);
    k_fill<<<(c->E + 255) / 256, 256, 0, c->stream>>>(c->E, 0.0, c->ck.p);
    HF_TRY(hf_assemble_values(c, c->cm.p, c->ck.p, c->axisym, c->valM1.p));
  }
  HF_TRY(hf_upload_nodal(c, s, c->source.p));
  c->have_source = true;
  return HF_OK;
}

extern "C" int hf_set_solver(hf_ctx* c, double rtol, int32_t max_iters, double warm, int32_t mode) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  if (!(rtol > 0.0) || max_iters <= 0 || mode < 0 || mode > 4) return hf_fail(HF_ERR_ARG, "hf_set_solver: bad arguments");
  c->rtol = rtol;
  c->max_iters = max_iters;
  c->warm = warm;
  c->mode = mode;
  return HF_OK;
}

// ---------------------------------------------------------------------------------------
// one backward-Euler step
// ---------------------------------------------------------------------------------------
// g_j = (amp - t_ic) exp(coeff r_j^2) + t_ic   (reference: run_with_diamond.py:354-359, bc.py:128-137)
__global__ void k_bc_gauss(int n, const int* __restrict__ dof, const double* __restrict__ r, double amp, double t_ic,
                           double coeff, double* __restrict__ g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[dof[i]] = (amp - t_ic) * exp(coeff * (r[i] * r[i])) + t_ic;
}

// assemble_vector + apply_lifting + set_bc (reference: run_with_diamond.py:474-479) fused with the
// start of the solve: b = M u_n [+ dt M1 s] - A0[:,bc] g, b[bc] = g; x0 = warm-started u_n with
// x0[bc] = g; scaled residual rhat = (b - A x0)/s = ((M u_n)_i - (A0 x0)_i)/s_i on free rows.
__global__ void __launch_bounds__(HF_BLOCK)
k_step_init(int N, int Npad, const int* __restrict__ rowptr, const int* __restrict__ col,
            const double* __restrict__ valM, const double* __restrict__ valA0, const double* __restrict__ valM1,
            const unsigned char* __restrict__ bcflag, const double* __restrict__ gfull,
            const double* __restrict__ u, const double* __restrict__ uprev, double warm,
            const double* __restrict__ src, double dt, const double* __restrict__ scale,
            double* __restrict__ b, double* __restrict__ xh, double* __restrict__ rh, HfCtrl* __restrict__ c) {
  __shared__ double sh[HF_BLOCK / 32];
  double l_rr = 0.0, l_bn = 0.0;
  for (int i = blockIdx.x * HF_BLOCK + threadIdx.x; i < Npad; i += gridDim.x * HF_BLOCK) {
    double xv = 0.0, rv = 0.0;
    if (i < N) {
      if (bcflag[i]) {
        const double g = gfull[i];
        b[i] = g;
        xv = g;
      } else {
        double t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
          const int j = col[k];
          const double m = valM[k], a = valA0[k];
          const double uj = u[j];
          double x0j, gj = 0.0;
          if (bcflag[j]) {
            gj = gfull[j];
            x0j = gj;
          } else {
            x0j = (warm != 0.0) ? fma(warm, uj - uprev[j], uj) : uj;
          }
          t1 = fma(m, uj, t1);
          if (valM1) t1 = fma(dt * valM1[k], src[j], t1);
          t2 = fma(a, x0j, t2);
          t3 = fma(a, gj, t3);
        }
        const double s = scale[i];
        const double bi = t1 - t3;
        b[i] = bi;
        const double ui = u[i];
        xv = s * ((warm != 0.0) ? fma(warm, ui - uprev[i], ui) : ui);
        rv = (t1 - t2) / s;
        const double bh = bi / s;
        l_bn = fma(bh, bh, l_bn);
        l_rr = fma(rv, rv, l_rr);
      }
    }
    xh[i] = xv;
    rh[i] = rv;
  }
  const double trr = hf_block_sum(l_rr, sh);
  const double tbn = hf_block_sum(l_bn, sh);
  if (threadIdx.x == 0) {
    c->part_rr[0][blockIdx.x] = trr;
    c->part_bn[blockIdx.x] = tbn;
  }
}

// End of a time step: u_{n+1} = xhat / s, u_n -> uprev.  hf_run folds two more per-step kernels into the same launch
// (each a 3 us launch for a few hundred threads of work): the watcher samples of this step (computed from xhat with
// the same division, so they are the bits of u) and the Gaussian boundary values of the NEXT step.
__global__ void k_step_finalize(int N, const double* __restrict__ xh, const double* __restrict__ scale,
                                double* __restrict__ u, double* __restrict__ uprev, int n_watch, const int* __restrict__ nodes,
                                double* __restrict__ hist, int n_gauss, const int* __restrict__ gdof, const double* __restrict__ gr,
                                double amp_next, double t_ic, double coeff, double* __restrict__ g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const double un = xh[i] / scale[i];
    uprev[i] = u[i];
    u[i] = un;
  }
  if (i < n_watch) hist[i] = xh[nodes[i]] / scale[nodes[i]];
  if (i < n_gauss) g[gdof[i]] = (amp_next - t_ic) * exp(coeff * (gr[i] * gr[i])) + t_ic;
}

__global__ void k_sample(int n, const int* __restrict__ nodes, const double* __restrict__ u, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = u[nodes[i]];
}

// Which PCG kernel solves with `op`: 1 = streaming kernel, one launch per iteration, the host polls (k_pcg_iter);
// 2 = persistent streaming kernel, one cooperative launch per solve (k_pcg_stream); 3 = on-chip patch kernel, one
// cooperative launch per solve (k_pcg_patch, or its pipelined variant k_pcg_pipe, which hf_get_solver_path reports
// as 4).  Paths 2 and 3 are fully asynchronous.
static bool has_patch_plan(const SellOp& op) { return op.pp_rpt != 0 || (op.plan_from && op.plan_from->pp_rpt != 0); }

static int pick_path(hf_ctx* c, const SellOp& op, int* path) {
  const int mode = c->force_mode >= 0 ? c->force_mode : c->mode;
  // a specific single-launch kernel was requested: plan it on demand
  if (mode >= 3 && !has_patch_plan(op)) HF_TRY(hf_patch_plan(c, const_cast<SellOp&>(op)));
  if (mode >= 3 && !has_patch_plan(op))
    return hf_fail(HF_ERR_STATE, "the on-chip PCG kernel was requested (solver mode 3 / 4) but the mesh does not fit on chip");
  if ((mode == 2 || (mode == 0 && !has_patch_plan(op))) && !op.st_grid) HF_TRY(hf_stream_plan(c, const_cast<SellOp&>(op)));
  if (mode == 2 && !op.st_grid)
    return hf_fail(HF_ERR_STATE, "the persistent streaming PCG kernel was requested (solver mode 2) but cannot be launched "
                                 "cooperatively on this device");
  if (mode == 0) *path = has_patch_plan(op) ? 3 : (op.st_grid ? 2 : 1);
  else *path = mode >= 3 ? 3 : mode;
  return HF_OK;
}

// one cooperative launch per solve
static int solve_single_launch(hf_ctx* c, const SellOp& op, int path, int step_slot, bool sum_parts) {
  if (path == 3) return hf_patch_solve_async(c, op, step_slot, sum_parts);
  return hf_stream_solve_async(c, op, step_slot, sum_parts);
}

// Synchronous completion of a persistent solve: read the control-block header.
static int finish_sync(hf_ctx* c, int* iters, double* relres) {
  PcgWork& w = c->ws;
  HF_CUDA(cudaMemcpyAsync(w.h_ctrl, w.ctrl.p, sizeof(double) * 3 + sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  c->last_iters = w.h_ctrl->itA;
  c->stat_iters += w.h_ctrl->itA;
  c->stat_relres = (w.h_ctrl->bn2 > 0.0) ? std::sqrt(w.h_ctrl->rr / w.h_ctrl->bn2) : 0.0;
  if (iters) *iters = w.h_ctrl->itA;
  if (relres) *relres = c->stat_relres;
  if (!w.h_ctrl->done) {
    char msg[160];
    snprintf(msg, sizeof msg, "PCG did not converge in %d iterations (relres %.3e)", w.h_ctrl->itA, c->stat_relres);
    return hf_fail(HF_ERR_NOCONV, msg);
  }
  return HF_OK;
}

// step_slot >= 0: fully asynchronous (persistent kernel only), iteration count goes to ws.step_iters[step_slot]
// `fuse` (hf_run): the Gaussian values of this step were written by the previous step's finalize kernel
// (fuse->gauss_done), which also takes this step's watcher samples and the next step's Gaussian values.
struct StepFuse {
  bool gauss_done = false, gauss_next = false;
  double amp_next = 0.0;
  int n_watch = 0;
  double* hist = nullptr;
};
static int step_device(hf_ctx* c, int use_gauss, double amp, double t_ic, double coeff, int* iters, double* relres,
                       int step_slot, int prof_slot = -1, const StepFuse* fuse = nullptr) {
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_step: operator not built");
  int path = 1;
  HF_TRY(pick_path(c, c->opA, &path));
  const bool persist = path >= 2;
  const bool gauss_now = use_gauss && c->n_gauss && !(fuse && fuse->gauss_done);
  c->stat_launches += 2 + (gauss_now ? 1 : 0);
  if (gauss_now)
    k_bc_gauss<<<(c->n_gauss + 255) / 256, 256, 0, c->stream>>>(c->n_gauss, c->gauss_dof.p, c->gauss_r.p, amp, t_ic, coeff,
                                                                c->gfull.p);
  PcgWork& w = c->ws;
  const double warm = c->have_prev ? c->warm : 0.0;
  k_step_init<<<w.grid, HF_BLOCK, 0, c->stream>>>(c->N, c->Npad, c->rowptr.p, c->col.p, c->valM.p, c->valA0.p,
                                                  c->have_source ? c->valM1.p : nullptr, c->bcflag.p, c->gfull.p, c->u.p,
                                                  c->uprev.p, warm, c->source.p, c->dt, c->opA.scale.p, c->b.p, w.x.p,
                                                  w.r.p, w.ctrl.p);
  HF_CUDA(cudaGetLastError());
  HF_TRY(hf_rc_project(c));
  if (!persist) HF_TRY(hf_pcg_prepare(c));          // the persistent kernel sums the partials itself
  const bool prof = c->profile && prof_slot >= 0 && (size_t)(2 * prof_slot + 1) < c->prof_ev.size();
  const unsigned long long l0 = c->stat_launches;
  if (prof) HF_CUDA(cudaEventRecord(c->prof_ev[2 * prof_slot], c->stream));
  if (persist) {
    HF_TRY(solve_single_launch(c, c->opA, path, step_slot, true));
    if (step_slot < 0) HF_TRY(finish_sync(c, iters, relres));
  } else {
    HF_TRY(hf_pcg_solve(c, c->opA, iters, relres));
  }
  if (prof) {
    HF_CUDA(cudaEventRecord(c->prof_ev[2 * prof_slot + 1], c->stream));
    c->stat_solve_launches += c->stat_launches - l0;
  }
  HF_TRY(hf_rc_store(c, c->opA));
  const int nw = fuse ? fuse->n_watch : 0, ng = (fuse && fuse->gauss_next && use_gauss) ? c->n_gauss : 0;
  const int nthr = std::max(c->N, std::max(nw, ng));
  k_step_finalize<<<(nthr + 255) / 256, 256, 0, c->stream>>>(c->N, w.x.p, c->opA.scale.p, c->u.p, c->uprev.p, nw, c->watch.p,
                                                          fuse ? fuse->hist : nullptr, ng, c->gauss_dof.p, c->gauss_r.p,
                                                          fuse ? fuse->amp_next : 0.0, t_ic, coeff, c->gfull.p);
  HF_CUDA(cudaGetLastError());
  c->have_prev = true;
  return HF_OK;
}

extern "C" int hf_step(hf_ctx* c, int32_t use_gauss, double amp, double t_ic, double coeff, int32_t* iters_out,
                       double* relres_out) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  cudaSetDevice(c->device);
  HF_TRY(step_device(c, use_gauss, amp, t_ic, coeff, iters_out, relres_out, -1));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return HF_OK;
}

// The time loop of hf_run on the device.  `path` >= 2: every step is one asynchronous cooperative solve, iteration
// counts land in ws.step_iters, failures are counted in ws.fail; path 1: the host polls every solve.
static int run_steps(hf_ctx* c, int path, int32_t n_steps, const double* amp, double t_ic, double coeff, int32_t n_watch,
                     double* fields, int32_t* iters) {
  const bool persist = path >= 2;
  for (int s = 0; s < n_steps; ++s) {
    int it = 0;
    StepFuse fuse;
    fuse.gauss_done = s > 0;                         // written by the finalize kernel of step s - 1
    fuse.gauss_next = s + 1 < n_steps;
    fuse.amp_next = fuse.gauss_next ? amp[s + 1] : 0.0;
    fuse.n_watch = n_watch;
    fuse.hist = n_watch ? c->hist.p + (size_t)s * n_watch : nullptr;
    HF_TRY(step_device(c, 1, amp[s], t_ic, coeff, &it, nullptr, persist ? s : -1, s, &fuse));
    if (iters && !persist) iters[s] = it;
    if (fields) {
      // XDMF field output needs the state on the host after every step; the copy is stream ordered
      const double* src = c->u.p;
      if (c->permuted) {   // the staging buffer is reused only after the copy below has run
        k_gather_nodal<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, 1, c->rank_d.p, c->u.p, c->stage.p);
        src = c->stage.p;
      }
      HF_CUDA(cudaMemcpyAsync(fields + (size_t)s * c->N, src, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  return HF_OK;
}

extern "C" int hf_run(hf_ctx* c, int32_t n_steps, const double* amp, double t_ic, double coeff, int32_t n_watch,
                      const int32_t* watch_nodes, double* hist, double* fields, int32_t* iters) {
  if (!c || n_steps < 0 || (n_steps && !amp) || n_watch < 0 || (n_watch && (!watch_nodes || !hist)))
    return hf_fail(HF_ERR_ARG, "hf_run: bad arguments");
  cudaSetDevice(c->device);
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_run: operator not built");
  std::vector<int> wn;
  if (hf_internal_nodes(c, n_watch, watch_nodes, wn) != HF_OK) return hf_fail(HF_ERR_ARG, "hf_run: watch node out of range");
  if (n_watch) {
    HF_TRY(c->watch.upload(wn.data(), n_watch, c->stream));
    if (c->hist.n < (size_t)n_steps * n_watch) HF_TRY(c->hist.alloc((size_t)n_steps * n_watch, c->stream));
  }
  int path = 1;
  HF_TRY(pick_path(c, c->opA, &path));
  const bool persist = path >= 2;
  PcgWork& w = c->ws;
  if (persist) {
    if (w.step_iters.n < (size_t)n_steps) HF_TRY(w.step_iters.alloc(n_steps, c->stream));
    HF_CUDA(cudaMemsetAsync(w.fail.p, 0, sizeof(int), c->stream));
    // state at the start of the run: a failed asynchronous solve is only seen at the end, and the run is then
    // repeated from here with the host-polled streaming kernel (see below)
    if (c->u0_keep.n != (size_t)2 * c->N) HF_TRY(c->u0_keep.alloc((size_t)2 * c->N, c->stream));
    HF_CUDA(cudaMemcpyAsync(c->u0_keep.p, c->u.p, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
    HF_CUDA(cudaMemcpyAsync(c->u0_keep.p + c->N, c->uprev.p, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
  }
  const bool had_prev = c->have_prev;
  // per-step field output (XDMF): pin the caller's buffer for the duration of the run so that the copies
  // are asynchronous DMA transfers that overlap the next step instead of staged synchronous ones
  struct HostPin {                                        // unpins on every exit path, after the stream has drained
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~HostPin() {
      if (p) {
        cudaStreamSynchronize(s);
        cudaHostUnregister(p);
      }
    }
  } pin;
  // (page-locking costs ~10 ms per call: only worth it when several steps' worth of fields come back)
  if (fields && n_steps >= 4) {
    if (cudaHostRegister(fields, sizeof(double) * (size_t)n_steps * c->N, cudaHostRegisterDefault) == cudaSuccess) {
      pin.p = fields;
      pin.s = c->stream;
    } else {
      cudaGetLastError();                                 // not fatal: pageable copies still work
    }
  }
  c->stat_solve_ms = 0.0;
  c->stat_solve_launches = 0;
  if (c->profile)
    while (c->prof_ev.size() < (size_t)2 * n_steps) {
      cudaEvent_t e;
      HF_CUDA(cudaEventCreate(&e));
      c->prof_ev.push_back(e);
    }
  HF_CUDA(cudaEventRecord(c->ev0, c->stream));
  HF_TRY(run_steps(c, path, n_steps, amp, t_ic, coeff, n_watch, fields, iters));
  if (persist && n_steps) {
    int nfail = 0;
    HF_TRY(w.fail.download(&nfail, 1, c->stream));
    if (nfail) {
      // A single-launch solve hit its iteration cap or produced a non-finite residual (e.g. a partial sum outside the
      // fixed-point range of the on-chip reduction).  The state it left behind is not trustworthy, so the whole run is
      // repeated from its initial state with the host-polled streaming kernel, which reports failures per step -
      // still the GPU path; if that fails as well the error is raised.
      c->stat_retries += 1;
      HF_CUDA(cudaMemcpyAsync(c->u.p, c->u0_keep.p, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
      HF_CUDA(cudaMemcpyAsync(c->uprev.p, c->u0_keep.p + c->N, sizeof(double) * c->N, cudaMemcpyDeviceToDevice, c->stream));
      c->have_prev = had_prev;
      hf_rc_reset(c);
      c->force_mode = 1;
      const int rc = run_steps(c, 1, n_steps, amp, t_ic, coeff, n_watch, fields, iters);
      c->force_mode = -1;
      if (rc != HF_OK) return rc;
      path = 1;
    }
  }
  HF_CUDA(cudaEventRecord(c->ev1, c->stream));
  if (n_watch && n_steps)
    HF_CUDA(cudaMemcpyAsync(hist, c->hist.p, sizeof(double) * (size_t)n_steps * n_watch, cudaMemcpyDeviceToHost, c->stream));
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
  if (path >= 2 && n_steps) {
    std::vector<int> hit(n_steps);
    HF_TRY(w.step_iters.download(hit.data(), n_steps, c->stream));
    for (int s = 0; s < n_steps; ++s) {
      if (iters) iters[s] = hit[s];
      c->stat_iters += hit[s];
    }
    c->last_iters = hit[n_steps - 1];
  }
  return HF_OK;
}

extern "C" int hf_set_sharing(hf_ctx* c, int32_t n_concurrent) {
  if (!c || n_concurrent < 1 || n_concurrent > 2) return hf_fail(HF_ERR_ARG, "hf_set_sharing: n_concurrent must be 1 or 2");
  if (c->share != n_concurrent) {      // the kernel plans depend on it
    c->share = n_concurrent;
    c->opA.struct_valid = c->opMr.struct_valid = false;
    c->opA.pp_rpt = c->opMr.pp_rpt = 0;
  c->opMr.plan_from = nullptr;
    c->op_built = c->proj_built = false;
  }
  return HF_OK;
}

extern "C" int hf_set_profile(hf_ctx* c, int32_t on) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  c->profile = on != 0;
  return HF_OK;
}

extern "C" int hf_get_solve_profile(hf_ctx* c, double* solve_ms, int64_t* solve_launches) {
  if (!c || !solve_ms || !solve_launches) return hf_fail(HF_ERR_ARG, "hf_get_solve_profile: null argument");
  *solve_ms = c->stat_solve_ms;
  *solve_launches = (int64_t)c->stat_solve_launches;
  return HF_OK;
}

extern "C" int hf_get_solver_path(hf_ctx* c) {
  if (!c) return hf_fail(HF_ERR_ARG, "null context");
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_get_solver_path: operator not built");
  int path = 1;
  HF_TRY(pick_path(c, c->opA, &path));
  return (path == 3 && hf_patch_pipelined(c, c->opA)) ? 4 : path;
}

extern "C" int hf_get_stats(hf_ctx* c, double* st) {
  if (!c || !st) return hf_fail(HF_ERR_ARG, "hf_get_stats: null argument");
  st[0] = c->stat_run_ms;
  st[1] = (double)c->stat_launches;
  st[2] = (double)c->stat_iters;
  st[3] = c->stat_relres;
  st[4] = (double)c->stat_retries;
  return HF_OK;
}

// Diagnostics: per-CTA clock64 cycles spent in the phases of the pipelined on-chip kernel during the last solve
// ([grid][2 warps][8 phases]; all zero unless the library was built with -DHF_PHASE_TIMING).  n = 0 switches it off.
extern "C" int hf_debug_phase_times(hf_ctx* c, int64_t* out, int32_t n) {
  if (!c || n < 0) return hf_fail(HF_ERR_ARG, "hf_debug_phase_times: bad arguments");
  cudaSetDevice(c->device);
  if (n == 0) {
    c->debug_phase.release();
    return HF_OK;
  }
  if (c->debug_phase.n != (size_t)n) return c->debug_phase.alloc(n, c->stream);     // first call: arm
  if (!out) return hf_fail(HF_ERR_ARG, "hf_debug_phase_times: null output");
  return c->debug_phase.download(reinterpret_cast<long long*>(out), n, c->stream);
}

extern "C" int hf_debug_fx_shift(hf_ctx* c, int32_t bits) {
  if (!c || bits < 0 || bits > 400) return hf_fail(HF_ERR_ARG, "hf_debug_fx_shift: bits must be in [0, 400]");
  c->debug_fx_shift = bits;
  return HF_OK;
}

extern "C" int hf_sample(hf_ctx* c, int32_t n, const int32_t* nodes, double* out) {
  if (!c || n < 0 || (n && (!nodes || !out))) return hf_fail(HF_ERR_ARG, "hf_sample: bad arguments");
  if (n == 0) return HF_OK;
  cudaSetDevice(c->device);
  std::vector<int> in;
  if (hf_internal_nodes(c, n, nodes, in) != HF_OK) return hf_fail(HF_ERR_ARG, "hf_sample: node out of range");
  DevBuf<int> d;
  DevBuf<double> o;
  HF_TRY(d.upload(in.data(), n, c->stream));
  HF_TRY(o.alloc(n, c->stream));
  k_sample<<<(n + 255) / 256, 256, 0, c->stream>>>(n, d.p, c->u.p, o.p);
  return o.download(out, n, c->stream);
}

// ---------------------------------------------------------------------------------------
// SpMV through the production kernel (parity + roofline tests)
// ---------------------------------------------------------------------------------------
__global__ void k_scale_in(int N, int Npad, const double* __restrict__ x, const double* __restrict__ s, double* __restrict__ o) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Npad) o[i] = (i < N) ? x[i] * s[i] : 0.0;
}
__global__ void k_scale_out(int N, const double* __restrict__ q, const double* __restrict__ s, double* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) y[i] = q[i] * s[i];
}
__global__ void k_ctrl_force(HfCtrl* c, int nparts) {
  for (int i = threadIdx.x; i < HF_MAX_PART; i += blockDim.x) c->part_rr[0][i] = 1.0;
  if (threadIdx.x == 0) {
    c->thr = -1.0;
    c->done = 0;
    c->itA = 0;
    c->itB = 0;
    c->nparts = nparts;
    c->alpha = 0.0;
    c->beta = 0.0;
    c->counter = 0u;
  }
}
int hf_spmv_device(hf_ctx* c, const SellOp& op);   // hf_pcg.cu

extern "C" int hf_spmv(hf_ctx* c, const double* x, double* y) {
  if (!c || !x || !y) return hf_fail(HF_ERR_ARG, "hf_spmv: null argument");
  if (!c->op_built) return hf_fail(HF_ERR_STATE, "hf_spmv: operator not built");
  cudaSetDevice(c->device);
  PcgWork& w = c->ws;
  DevBuf<double> dx;
  HF_TRY(dx.alloc(c->N, c->stream));
  HF_TRY(hf_upload_nodal(c, x, dx.p));
  const int g = (c->Npad + 255) / 256;
  k_scale_in<<<g, 256, 0, c->stream>>>(c->N, c->Npad, dx.p, c->opA.scale.p, w.r.p);
  k_ctrl_force<<<1, 256, 0, c->stream>>>(w.ctrl.p, w.grid);
  HF_TRY(hf_spmv_device(c, c->opA));
  k_scale_out<<<g, 256, 0, c->stream>>>(c->N, w.q1.p, c->opA.scale.p, dx.p);
  HF_CUDA(cudaGetLastError());
  return hf_download_nodal(c, dx.p, y, 1);
}

// ---------------------------------------------------------------------------------------
// r-weighted L2 projection of grad(u) onto vector P1 (reference: run_no_diamond.py:471-491, :544-550)
// ---------------------------------------------------------------------------------------
// load b_c[i] = sum_T (d_c u)_T |T| (2 r_i + r_j + r_k)/12, written Jacobi-scaled into rz / rr.
__global__ void __launch_bounds__(256)
k_grad_load(int N, int Npad, const double* __restrict__ xy, const int* __restrict__ cells,
            const int* __restrict__ n2c_ptr, const int* __restrict__ n2c_idx, const double* __restrict__ u,
            const double* __restrict__ scale, double* __restrict__ bz, double* __restrict__ br) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  double az = 0.0, ar = 0.0;
  if (i < N) {
    for (int a = n2c_ptr[i]; a < n2c_ptr[i + 1]; ++a) {
      const int e = n2c_idx[a];
      const int v0 = cells[3 * e], v1 = cells[3 * e + 1], v2 = cells[3 * e + 2];
      const double z0 = xy[2 * v0], r0 = xy[2 * v0 + 1], z1 = xy[2 * v1], r1 = xy[2 * v1 + 1], z2 = xy[2 * v2],
                   r2 = xy[2 * v2 + 1];
      const double b0 = r1 - r2, b1 = r2 - r0, b2 = r0 - r1;
      const double c0 = z2 - z1, c1 = z0 - z2, c2 = z1 - z0;
      const double det = b0 * c1 - b1 * c0;
      const double area = 0.5 * fabs(det);
      const double u0 = u[v0], u1 = u[v1], u2 = u[v2];
      const double dz = (b0 / det) * u0 + (b1 / det) * u1 + (b2 / det) * u2;
      const double dr = (c0 / det) * u0 + (c1 / det) * u1 + (c2 / det) * u2;
      const double ri = (v0 == i) ? r0 : (v1 == i) ? r1 : r2;
      const double wgt = area * (ri + (r0 + r1 + r2)) / 12.0;
      az = fma(wgt, dz, az);
      ar = fma(wgt, dr, ar);
    }
    const double s = scale[i];
    az /= s;
    ar /= s;
  }
  bz[i] = az;
  br[i] = ar;
}

__global__ void k_store_comp(int N, const double* __restrict__ xh, const double* __restrict__ scale, int comp,
                             double* __restrict__ g2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) g2[2 * i + comp] = xh[i] / scale[i];
}

extern "C" int hf_project_gradient(hf_ctx* c, double* grad, int32_t* iters_out) {
  if (!c || !grad) return hf_fail(HF_ERR_ARG, "hf_project_gradient: null argument");
  if (c->N == 0 || c->nv != 3) return hf_fail(HF_ERR_STATE, "hf_project_gradient: needs a triangle mesh");
  cudaSetDevice(c->device);
  if (!c->proj_built) {
    DevBuf<double> vals;
    HF_TRY(vals.alloc(c->nnz, c->stream));
    k_fill<<<(c->E + 255) / 256, 256, 0, c->stream>>>(c->E, 1.0, c->cm.p);
    k_fill<<<(c->E + 255) / 256, 256, 0, c->stream>>>(c->E, 0.0, c->ck.p);
    HF_TRY(hf_assemble_values(c, c->cm.p, c->ck.p, 1, vals.p));
    HF_TRY(hf_build_sell(c, vals, false, c->opMr, nullptr));
    if (c->opA.pp_rpt) {
      c->opMr.plan_from = &c->opA;               // same sparsity pattern: same patches, halo lists, local columns
    } else if (!c->opMr.pp_rpt) {
      HF_TRY(hf_patch_plan(c, c->opMr));
    }
    HF_TRY(c->proj_b.alloc((size_t)2 * c->Npad, c->stream));
    HF_TRY(c->proj_g.alloc((size_t)2 * c->N, c->stream));
    c->proj_built = true;
  }
  PcgWork& w = c->ws;
  const int g = (c->Npad + 255) / 256;
  k_grad_load<<<g, 256, 0, c->stream>>>(c->N, c->Npad, c->xy.p, c->cells.p, c->n2c_ptr.p, c->n2c_idx.p, c->u.p,
                                        c->opMr.scale.p, c->proj_b.p, c->proj_b.p + c->Npad);
  HF_CUDA(cudaGetLastError());
  const int keep = c->last_iters;
  int total = 0;
  for (int comp = 0; comp < 2; ++comp) {
    HF_CUDA(cudaMemcpyAsync(w.r.p, c->proj_b.p + (size_t)comp * c->Npad, sizeof(double) * c->Npad, cudaMemcpyDeviceToDevice,
                            c->stream));
    HF_CUDA(cudaMemsetAsync(w.x.p, 0, sizeof(double) * c->Npad, c->stream));
    HF_TRY(hf_pcg_prepare_from_r(c));
    int it = 0;
    c->last_iters = 40;
    int path = 1;
    HF_TRY(pick_path(c, c->opMr, &path));
    if (path >= 2) {
      HF_TRY(solve_single_launch(c, c->opMr, path, -1, false));
      HF_TRY(finish_sync(c, &it, nullptr));
    } else {
      HF_TRY(hf_pcg_solve(c, c->opMr, &it, nullptr));
    }
    if (getenv("HF_DEBUG")) fprintf(stderr, "[hf] projection comp %d: %d iterations, relres %.3e\n", comp, it, c->stat_relres);
    total += it;
    k_store_comp<<<(c->N + 255) / 256, 256, 0, c->stream>>>(c->N, w.x.p, c->opMr.scale.p, comp, c->proj_g.p);
  }
  c->last_iters = keep;
  if (iters_out) *iters_out = total;
  return hf_download_nodal(c, c->proj_g.p, grad, 2);
}
