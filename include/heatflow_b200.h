/* heatflow_b200 - C-ABI of the B200 (sm_100a) FEM heat-conduction hot path.
 *
 * Drop-in boundary for heatflow's per-timestep solve.  The reference has no FFI of its own:
 * its hot path is a sequence of dolfinx / petsc4py calls made inline by the Python runners.
 * Each entry point below replaces one of those third-party call sites (reference file:line
 * given per function).  Conventions:
 *   - Python (or any host) owns all host arrays; the library owns device memory behind an
 *     opaque handle.  One handle <-> one CUDA device + one stream.  Not thread-safe per handle.
 *   - Every call returns 0 on success, <0 on error; hf_last_error() gives the message of the
 *     last failing call of the calling thread.  No C++ exception crosses the boundary.
 *   - All floating point is IEEE double; all indices are int32 (dolfinx default).
 *   - There is no CPU fallback: without a usable CUDA device hf_create fails.
 */
#ifndef HEATFLOW_B200_H
#define HEATFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hf_ctx hf_ctx;

/* error codes */
#define HF_OK 0
#define HF_ERR_ARG (-1)
#define HF_ERR_CUDA (-2)
#define HF_ERR_STATE (-3)
#define HF_ERR_NOCONV (-4)

int hf_version(void);
const char* hf_last_error(void);
int hf_device_count(void);

/* create a context on CUDA device `device` (one per process/GPU; reference workers are
 * separate spawned processes, parameter_sweep.py:322-327).  NULL on failure. */
hf_ctx* hf_create(int device);
void hf_destroy(hf_ctx* ctx);

/* Internal node numbering used for the device data structures: 0 = auto (default), 1 = as given,
 * 2 = Hilbert-curve order (compact 2-D patches per CTA; what the streaming and ensemble kernels
 * want).  Call before hf_set_mesh.  Purely internal: every array crossing this ABI stays in the
 * caller's numbering, which is also the dof numbering (dolfinx P1: dof i = node i). */
int hf_set_ordering(hf_ctx* ctx, int32_t ordering);

/* Mesh arrays as gmshio.model_to_mesh would hand them to dolfinx
 * (reference: run_with_diamond.py:240-245): node coordinates xy[N,2] = (z, r), cell
 * connectivity cells[E,nv] (nv = 3 triangles, nv = 2 intervals for the 1-D path,
 * run_no_diamond_1d.py:94), cell_tag[E] = physical tag.  Builds the P1 dof map
 * (identity: dof i = node i), the node->cell adjacency and the CSR sparsity pattern
 * (fem.functionspace + create_sparsity_pattern inside assemble_matrix, :279, :381). */
int hf_set_mesh(hf_ctx* ctx, int32_t n_nodes, int32_t n_cells, int32_t nv,
                const double* xy, const int32_t* cells, const int32_t* cell_tag);

/* DG0 coefficient tables by tag (reference: run_with_diamond.py:286-301). */
int hf_set_materials(hf_ctx* ctx, int32_t n_tags, const int32_t* tags,
                     const double* kappa, const double* rho_c);

/* Dirichlet dofs with list-order "last wins" already resolved by the host
 * (reference: dirichlet_bc/bc.py:104-112, run_with_diamond.py:361-374).
 * bc_dofs: sorted unique; bc_value[n_bc]: current value g of every dof;
 * gauss_slot[n_gauss]: indices into bc_dofs whose value follows the Gaussian-in-r profile,
 * gauss_r[n_gauss]: their r coordinate (bc.py:128-137, run_with_diamond.py:354-359). */
int hf_set_bcs(hf_ctx* ctx, int32_t n_bc, const int32_t* bc_dofs, const double* bc_value,
               int32_t n_gauss, const int32_t* gauss_slot, const double* gauss_r);

/* Overwrite the Dirichlet values g (general RowDirichletBC.update with a host callable). */
int hf_set_bc_values(hf_ctx* ctx, const double* bc_value);

/* fem.form + assemble_matrix(lhs_form, bcs) + solver set-up
 * (reference: run_with_diamond.py:321-337, :381-394):
 *   M  = sum_T rho_c_T int phi_i phi_j w        A0 = M + dt sum_T kappa_T int grad phi_i . grad phi_j w
 *   A  = A0 with Dirichlet rows/cols zeroed and unit diagonal,  w = r (axisymmetric) or 1.
 * Also builds the Jacobi-scaled operator D^-1/2 A D^-1/2 in sliced-ELL form for the PCG. */
int hf_build_operator(hf_ctx* ctx, double dt, int32_t axisymmetric);

int hf_get_sizes(hf_ctx* ctx, int32_t* n_nodes, int64_t* nnz);
/* CSR arrays for the bit-exact pattern check; any pointer may be NULL.
 * val_A = operator with BCs applied, val_M = mass, val_A0 = operator without BCs. */
int hf_get_csr(hf_ctx* ctx, int32_t* rowptr, int32_t* col, double* val_A, double* val_M, double* val_A0);

/* u_n (reference: run_with_diamond.py:317-319) */
int hf_set_state(hf_ctx* ctx, const double* u);
int hf_get_state(hf_ctx* ctx, double* u);
/* nodal source s of the 1-D radial correction, b += dt * M_1 s
 * (reference: run_no_diamond_1d.py:543-544, :743-747); NULL clears it. */
int hf_set_source(hf_ctx* ctx, const double* s);

/* solver options: rtol on ||r||_{D^-1} / ||b_free||_{D^-1}, iteration cap, warm start
 * (x0 = u_n + warm*(u_n - u_{n-1})), PCG kernels: 0 = auto (on-chip patch kernel when the mesh
 * fits, else the persistent streaming kernel), 1 = streaming kernel with one launch per PCG iteration
 * (the host polls for convergence), 2 = persistent streaming kernel (one cooperative launch per solve,
 * the iteration loop runs on the device), 3 = on-chip patch kernel (pipelined CG variant when the mesh
 * is small enough for its six register-resident vectors per row: <= 148 x 1024 dofs), 4 = on-chip patch kernel, classic CG only. */
int hf_set_solver(hf_ctx* ctx, double rtol, int32_t max_iters, double warm, int32_t mode);

/* Initial guess from the previous time steps: keep the corrections of up to max_vectors
 * solves (A-orthogonalised; the basis is frozen once full) and start every solve from the Galerkin projection of its
 * right-hand side onto them.  The solver, its tolerance and therefore the converged answer are
 * unchanged (the reference re-uses its LU factors across steps in the same spirit,
 * run_with_diamond.py:389-394); 0 disables.  Memory: 2 * max_vectors * N doubles.  The basis
 * belongs to one simulation: hf_set_state and hf_build_operator drop it (measured: a basis
 * carried over to a simulation with other boundary data fills up with directions the new
 * run does not use and costs more than it saves). */
int hf_set_recycle(hf_ctx* ctx, int32_t max_vectors);

/* One backward-Euler step (reference loop body, run_with_diamond.py:471-481):
 * Gaussian BC update with amplitude `amp` (g = (amp - t_ic) exp(coeff r^2) + t_ic), RHS
 * assemble_vector + apply_lifting + set_bc, then the solve, result overwrites u_n.
 * use_gauss = 0 keeps the current g (set by hf_set_bcs / hf_set_bc_values). */
int hf_step(hf_ctx* ctx, int32_t use_gauss, double amp, double t_ic, double coeff,
            int32_t* iters_out, double* relres_out);

/* The RHS vector b of the last step (after apply_lifting and set_bc) - parity tests only. */
int hf_get_rhs(hf_ctx* ctx, double* b);

/* Whole time loop on the device: n_steps steps with amplitudes amp[n_steps]; after every
 * step the temperatures at watch_nodes[n_watch] are stored in hist[n_steps, n_watch]
 * (reference: run_with_diamond.py:485-493) and, when fields != NULL, the full state in
 * fields[n_steps, N] (xdmf.write_function, :483-484).  iters[n_steps] may be NULL. */
int hf_run(hf_ctx* ctx, int32_t n_steps, const double* amp, double t_ic, double coeff,
           int32_t n_watch, const int32_t* watch_nodes, double* hist, double* fields,
           int32_t* iters);

int hf_sample(hf_ctx* ctx, int32_t n, const int32_t* nodes, double* out);

/* Parameter sweeps on meshes that fit on chip: the on-chip PCG kernel is bound by the latency of its
 * grid reduction, so two independent simulations (two contexts driven from two host threads, each on
 * its own stream) sharing the SMs finish 1.44 x sooner than back to back.  n_concurrent = 2 plans the
 * kernel for two co-resident CTAs per SM (half the registers / shared memory each; a single solve is
 * ~15 % slower); 1 (default) gives one solve the whole SM.  Call before hf_build_operator. */
int hf_set_sharing(hf_ctx* ctx, int32_t n_concurrent);

/* Kernel timing for the roofline numbers: with profiling on, hf_run brackets the PCG solve of every
 * time step (the persistent kernel launch, or the streaming kernel launches of the step) with CUDA
 * events on the context stream; hf_get_solve_profile returns their summed device time in ms and the
 * number of solver kernels launched inside the brackets during the last hf_run.  Off by default (two
 * event records per step). */
int hf_set_profile(hf_ctx* ctx, int32_t on);
int hf_get_solve_profile(hf_ctx* ctx, double* solve_ms, int64_t* solve_launches);

/* Which PCG kernel hf_step / hf_run will use for the current operator and hf_set_solver mode:
 * 1 = streaming kernel, one launch per iteration; 2 = persistent streaming kernel; 3 = on-chip patch kernel
 * (the mesh fits in the SMs' shared memory + registers), classic CG; 4 = on-chip patch kernel, pipelined CG
 * (the grid reduction overlaps the SpMV); < 0 on error. */
int hf_get_solver_path(hf_ctx* ctx);

/* Counters for benchmarking: stats[0] = device time of the step loop of the last hf_run /
 * hf_ens_run in ms (CUDA events on the context stream), stats[1] = kernels launched since
 * hf_create (graph kernel nodes included), stats[2] = PCG iterations since hf_create,
 * stats[3] = relative residual of the last solve, stats[4] = hf_run calls that were repeated with the
 * host-polled streaming kernel because a single-launch solve failed (iteration cap, non-finite or
 * out-of-range partial sum; the repeat is still the GPU path - there is no CPU fallback). */
int hf_get_stats(hf_ctx* ctx, double* stats5);

/* Diagnostics for the tests: shrink the fixed-point range of the on-chip kernel's grid reduction by
 * 2^-bits so that its overflow path (sentinel + failure counter + repeat of the run) can be exercised. */
int hf_debug_fx_shift(hf_ctx* ctx, int32_t bits);

/* Diagnostics (profiling builds, -DHF_PHASE_TIMING): clock64 cycles per phase of the pipelined on-chip kernel in
 * the last solve, out[grid][2][8].  The first call with a new n arms the buffer, later calls read it; n = 0 disarms. */
int hf_debug_phase_times(hf_ctx* ctx, int64_t* out, int32_t n);

/* r-weighted L2 projection of grad(u_n) onto vector P1, grad[N,2] = (d/dz, d/dr)
 * (reference: run_no_diamond.py:471-491, :544-550). */
int hf_project_gradient(hf_ctx* ctx, double* grad, int32_t* iters_out);

/* y = A_bc x with the sliced-ELL scaled operator mapped back to physical units
 * (parity/roofline tests of the SpMV kernel). */
int hf_spmv(hf_ctx* ctx, const double* x, double* y);

/* Device-resident benchmarking helpers: time `reps` launches of the fused PCG kernels on the
 * current operator with CUDA events (ms per launch written to ms_out[2] = {spmv, update}). */
int hf_bench_kernels(hf_ctx* ctx, int32_t reps, int32_t flush_l2, float* ms_out);

/* ---- ensemble (parameter_sweep.py:123-192, :436-438): B simulations on one mesh --------
 * Variant s has conductivity k_sample[s] on cells tagged sample_tag and Gaussian width
 * fwhm via coeff[s] = -4 ln2 / fwhm_s^2; everything else is shared. */
int hf_ens_create(hf_ctx* ctx, int32_t batch, const double* k_sample, const double* coeff,
                  int32_t sample_tag);
int hf_ens_run(hf_ctx* ctx, int32_t n_steps, const double* amp, double t_ic,
               int32_t n_watch, const int32_t* watch_nodes, double* hist /*[B,n_steps,n_watch]*/,
               int32_t* iters /*[n_steps]*/);
int hf_ens_get_state(hf_ctx* ctx, double* u /*[B,N]*/);
/* Which kernels advance the current tile: 1 = streaming ensemble kernels (one launch per PCG iteration, operator read
 * from HBM once per tile and iteration), 5 = batched on-chip kernel (one cooperative launch per time step; tiles of
 * up to 4 variants on meshes whose 1024-row patches fit the SMs).  No reference counterpart (diagnostics). */
int hf_ens_get_path(hf_ctx* ctx);
int hf_ens_destroy(hf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* HEATFLOW_B200_H */
