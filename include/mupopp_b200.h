/* mupopp_b200 -- C-ABI of the B200-native MuPoPP per-timestep hot path.
 *
 * The reference (sashgeophysics/MuPoPP) has NO native/FFI interface: its hot path is the
 * un-vendored DOLFIN 2017.1 / FFC / PETSc stack reached through Python calls.  Each entry
 * point below therefore cites the reference CALL SITE whose work it replaces
 * (paths relative to /root/reference/src).  The Python host side (mupopp_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller (torch tensors); the
 *    library never frees caller memory.  fp64 values, int32 indices.
 *  - every call returns 0 on success, non-zero on failure; mp_last_error() gives the text.
 *  - one CUDA stream per mp_ctx; a handle is not thread-safe, distinct handles are.
 *  - no CPU fallback: a call without a usable CUDA device fails with an error.
 */
#ifndef MUPOPP_B200_H
#define MUPOPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mp_ctx mp_ctx;
typedef struct mp_plan mp_plan;
typedef struct mp_halo mp_halo;

/* ------------------------------------------------------------------ context */
int mp_version(void);
const char *mp_last_error(void);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = the CUDA default stream */
int mp_ctx_create(int device, void *stream, mp_ctx **out);
int mp_ctx_destroy(mp_ctx *ctx);
int mp_ctx_sync(mp_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t mp_ctx_launch_count(const mp_ctx *ctx);
/* optional CUDA-event timing of every SpMV launch (the dominant kernel) on the context's stream:
 * enable with a sample capacity, then read (count, summed milliseconds) -- used by bench.py's roofline */
int mp_ctx_timing(mp_ctx *ctx, int enable, int max_samples);
int mp_ctx_timing_read(mp_ctx *ctx, int *nsamples, double *total_ms);
/* tuning knobs for A/B measurements: "spmv_variant" = 0 SELL-32 when mirrored, else CSR CTA-stream (default), 1 CSR CTA-stream, 2 CSR sub-warp-per-row, 3 CSR warp-stream;
 * "spmv_c8" = 1 (default) use the compressed (slice-uniform) SELL columns when the operator has them, 0 always the int32 columns;
 * "spmv_lpr" = 0 (default) pick the lanes per row (1, 2 or 4) from the operator size, else force 1 / 2 / 4;
 * "pcg_variant" = -1 (default: single-reduction CG on several GPUs, classic on one), 0 classic PCG, 1 single-reduction (Chronopoulos-Gear) PCG;
 * "peer" = 0/1 NCCL or peer-memory collectives (after mp_comm_peer_attach) */
int mp_ctx_set_option(mp_ctx *ctx, const char *name, int value);

/* ------------------------------------------------------------------ mesh view (device pointers) */
typedef struct {
    int32_t dim;            /* 2 (triangles) or 3 (tetrahedra) */
    int64_t ncell;
    int64_t nvert;
    const int32_t *d_cells; /* [ncell][dim+1] vertex ids, ascending per cell (UFC ordering) */
    const double *d_coords; /* [nvert][dim] */
    /* optional [ncell][dim+1] P1 dof ids when they differ from the vertex ids (periodic `constrained_domain`,
     * run/CCS/CCS_3D_top_source.py:134-158,211): geometry comes from d_cells, nodal fields are gathered through d_cell_dofs */
    const int32_t *d_cell_dofs;
    /* optional [ncell][dim+1]: position of each local row in the ROW-ORDERED element storage of a P1xP1 plan
     * (mp_plan_row_positions); when set, the P1 element kernels store row i of cell c at elem[pos[c][i]] so that
     * mp_scatter_*_ordered stream their input (negative position = row not assembled on this rank) */
    const int32_t *d_row_pos;
} mp_mesh;

/* ------------------------------------------------------------------ scatter plan
 * Replaces DOLFIN's dofmap + SparsityPatternBuilder + MatSetValuesLocal inside
 * assemble_system / solve(a==L,..) (call sites: modules/mupopp.py:295-297,415-417,
 * run/CCS2D/CCS_3comp.py:261, run/darcy_advection_diffusion/darcy_advection_diffusion.py:81,114).
 * Built once per (row dofmap, column dofmap); holds the CSR pattern (columns ascending), the
 * dof->(cell,local) incidence lists and the per-entry slot map used by the deterministic
 * row-gather scatter.  d_col_dofs == NULL, lc == 0 gives a vector-only plan. */
int mp_plan_create(mp_ctx *ctx, const int32_t *d_row_dofs, int32_t lr, const int32_t *d_col_dofs, int32_t lc,
                   int64_t ncell, int64_t nrow, int64_t ncol, mp_plan **out);
int64_t mp_plan_nnz(const mp_plan *plan);
int32_t mp_plan_max_row_nnz(const mp_plan *plan);
int mp_plan_export_csr(mp_ctx *ctx, const mp_plan *plan, int32_t *d_rowptr, int32_t *d_colidx);
int mp_plan_destroy(mp_plan *plan);
/* dof -> (cell, local row) incidence lists of the plan (row r: d_inc[d_inc_ptr[r] .. d_inc_ptr[r+1]) = cell*lr + li, ascending) and the
 * CSR slot of every (incidence, local column) pair, uint16 [ninc][lc], 0xffff = not in the pattern.  Any out pointer may be NULL. */
int mp_plan_incidence(const mp_plan *plan, int64_t *nrow, int32_t *lr, int32_t *lc, int32_t *max_row_nnz, const int32_t **d_rowptr,
                      const int32_t **d_inc_ptr, const int32_t **d_inc, const uint16_t **d_slot);
/* SELL-32 layout of the plan's pattern (library-owned device arrays): number of slices, padded entries, pointers */
int mp_plan_sell_info(const mp_plan *plan, int64_t *nslice, int64_t *nentries, const int32_t **d_sliceptr, const int32_t **d_col);
/* compressed column indices of the SELL-32 layout (slice-uniform offsets, see mp_csr): number of (slice, j) positions that needed explicit
 * columns (-1: no compressed mirror) and the library-owned device arrays */
int mp_plan_sellu_info(const mp_plan *plan, int64_t *nexplicit, const int32_t **d_meta, const int32_t **d_xcol);
/* d_sell_vals[nentries] <- CSR values permuted into the SELL-32 layout (padding = 0); call after assembly / Dirichlet */
int mp_csr_to_sell(mp_ctx *ctx, const mp_plan *plan, const double *d_csr_vals, double *d_sell_vals);

/* vals[slot] = sum over incident cells (ascending cell id) of elem_mat[cell][i][j]   (no atomics) */
int mp_scatter_matrix(mp_ctx *ctx, const mp_plan *plan, const double *d_elem_mat, double *d_vals);
/* out[r] = beta*out[r] + sum over incident (cell,i) of elem_vec[cell][i] */
int mp_scatter_vector(mp_ctx *ctx, const mp_plan *plan, const double *d_elem_vec, double beta, double *d_out);

/* Row-ordered element storage: elem_mat[e][lc] / elem_vec[e] with e the incidence index (row-major, ascending cell id inside a row),
 * written directly by the P1 element kernels when mp_mesh.d_row_pos is set.  The scatter then reads contiguous memory
 * (the cell-ordered variant reads one random 32-byte sector per (cell,row) pair, 3.5x DRAM amplification: profiles/). */
int mp_plan_row_positions(const mp_plan *plan, const int32_t **d_pos);
int mp_scatter_matrix_ordered(mp_ctx *ctx, const mp_plan *plan, const double *d_elem_mat_ordered, double *d_vals);
int mp_scatter_vector_ordered(mp_ctx *ctx, const mp_plan *plan, const double *d_elem_vec_ordered, double beta, double *d_out);

/* DOLFIN assemble_system semantics (symmetric elimination): rhs -= A[:,D] g; zero rows/cols D;
 * A[D,D] = 1; rhs[D] = g.  d_isbc[ncol] flags, d_g[ncol] values (ignored where isbc == 0). */
int mp_apply_dirichlet_sym(mp_ctx *ctx, int64_t nrow, const int32_t *d_rowptr, const int32_t *d_colidx, double *d_vals,
                           double *d_rhs, const uint8_t *d_isbc, const double *d_g);

/* ------------------------------------------------------------------ P1 element kernels (mode F)
 * Replace FFC tabulate_tensor for the forms of modules/mupopp.py (line ranges per function). */
typedef struct {
    double Pe, Da, phi, beta, dt;
    double sigma, rho0, gamma;   /* buoyancy sigma*K*(rho0+gamma*c0) v.zhat: mupopp.py:804-807 (sigma=-1), :1201-1204 (+1) */
    double frac[3];              /* R1,R2,R3 prefactors: mupopp.py:818-821, :925-931, :1222-1229 */
    double zhat[3];
    double Kconst;               /* used when d_K == NULL */
    int32_t reaction;            /* 0: Da*c0*c1 (CCS)   1: Da*c0*c1^2 (DarcyAdvection)   2: Da*c0 (single component, mupopp.py:650) */
    int32_t ncomp;               /* 1 (quadrature kernel only), 2 or 3 concentration fields */
    int32_t supg;                /* 1 Sendur, 2 Codina, 3 Brooks-Hughes: mupopp.py:835-846; 4: h/(2|w|) (mupopp.py:663); 5: Sendur + Da*c1^n(x)/phi (mupopp.py:738, quadrature kernel only) */
    int32_t pad;
} mp_transport_params;

/* elem_mat[c][i][j] = scale * int phi_i phi_j   (P1 mass; c1/c2 rows mupopp.py:826,936-937,1234-1235) */
int mp_elem_p1_mass(mp_ctx *ctx, const mp_mesh *mesh, double scale, double *d_elem_mat);
/* elem_mat = scale * K_cell * int grad phi_i . grad phi_j   (pressure operator div(k grad p)); d_K nodal P1 or NULL */
int mp_elem_p1_stiffness(mp_ctx *ctx, const mp_mesh *mesh, const double *d_K, double Kconst, double scale, double *d_elem_mat);
/* c1/c2 rows: elem_b1 = int w (c1n - dt R2/(1-phi)), elem_b2 = int w (c2n + dt R3/(1-phi)) (NULL when ncomp==2) */
int mp_elem_p1_reaction_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_c0n,
                            const double *d_c1n, const double *d_c2n, double *d_elem_b1, double *d_elem_b2);
/* the reaction load alone: elem_T[c][i] = scale * int w_i R(c0n, c1n) (R = c0 c1 or c0 c1^2, prm->reaction).  The c1 and c2 rows share it up to
 * the factors -dt frac[1] Da/(1-phi) and +dt frac[2] Da/(1-phi), so one mass solve serves both increments. */
int mp_elem_p1_reaction_load(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_c0n,
                             const double *d_c1n, double scale, double *d_elem_T);
/* c0 row (Crank-Nicolson + SUPG, mupopp.py:823-848 / :934-958 / :1231-1258) with the lagged DG0 velocity d_ucell[c][dim]:
 * elem_mat (may be NULL to skip) and elem_vec; the new c1/c2 enter the rhs through the SUPG residual. */
int mp_elem_p1_transport(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_ucell,
                         const double *d_c0n, const double *d_c1n, const double *d_c2n, const double *d_c1new,
                         const double *d_c2new, const double *d_fsrc, double *d_elem_mat, double *d_elem_vec);

/* The same c0 row assembled ROW-WISE in one pass, without the element-tensor round trip through HBM: one thread per CSR row walks
 * the row's incident cells (plan incidence lists, ascending cell id => fixed summation order), recomputes each cell's contribution to
 * its row and writes the CSR values, the SELL mirror (d_sell_vals, may be NULL), the right-hand side with the symmetric Dirichlet
 * elimination (d_isbc / d_g over all local dofs; what assemble_system does in modules/mupopp.py:289-303), the Jacobi diagonal d_dinv
 * (may be NULL) and -- when line_stride > 0 -- the bands of the line preconditioner (d_lo, d_diag, d_up: inputs of mp_line_factor).
 * P1 plans only (lr == lc == dim + 1). */
int mp_assemble_p1_transport_rows(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const mp_plan *plan,
                                  const double *d_ucell, const double *d_c0n, const double *d_c1n, const double *d_c2n,
                                  const double *d_c1new, const double *d_c2new, const double *d_fsrc, const uint8_t *d_isbc,
                                  const double *d_g, double *d_vals, double *d_sell_vals, double *d_rhs, double *d_dinv,
                                  int64_t line_stride, double *d_lo, double *d_diag, double *d_up);

/* single-component form DarcyAdvection.advection_diffusion (mupopp.py:627-664), DG0 velocity variant */
int mp_elem_p1_adr_single(mp_ctx *ctx, const mp_mesh *mesh, double Pe, double Da, double dt, const double *d_ucell,
                          const double *d_cn, double *d_elem_mat, double *d_elem_vec);
/* pressure rhs: -sigma * int (K/phi) rho(c0bar) zhat.grad q */
int mp_elem_p1_pressure_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_K,
                            const double *d_c0, double *d_elem_vec);
/* fused pointwise velocity recovery u = -(K/phi)(grad p + sigma rho zhat) per cell (north_star item 3) */
int mp_recover_velocity(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const double *d_K,
                        const double *d_p, const double *d_c0, double *d_ucell);

/* elem_vec[c][i] = |cell| * (d_ucell ? d_ucell[c][comp] : 1): scattered and divided by the lumped volumes it gives the
 * volume-weighted vertex average of a cellwise field (the nodal velocity written by `velocity_out << u0`, run/CCS2D/CCS_3comp.py:265-266) */
int mp_elem_p1_cell_load(mp_ctx *ctx, const mp_mesh *mesh, const double *d_ucell, int comp, double *d_elem_vec);

/* Generic affine element kernel ("tensor representation"): elem_mat[c][i][j] = scale * sum_t geo_t(c) * T0[t][i][j]
 *   kind 0 MASS  (1 tensor,  geo = |detJ|)                      kind 1 GRAD  (dim tensors, geo_a = |detJ| Jinv[a][dir]: row value x column d/dx_dir)
 *   kind 2 GRADT (as GRAD with row derivative x column value)   kind 3 STIFF (dim^2 tensors, geo_ab = |detJ| (Jinv Jinv^T)_ab)
 *   kind 4 SYMGRAD: block (a,b) = (dir/dim, dir%dim) of 2 symgrad(u):symgrad(v) with the STIFF tensors (compaction, mupopp.py:361)
 * coef_kind 0 none | 1 nodal P1 d_coef[nvert], integrated exactly (T0 then has (dim+1) sub-tensors per t) | 2 cellwise d_coef[ncell].
 * T0 comes from mupopp_b200/reference_element.py and is staged in shared memory once per CTA; a CTA then produces tiles of 64 cells with
 * coalesced stores.  Mode R operators: P2 mass u.v, K div(v) p, div(u) q (mupopp.py:620,804-807,1201-1204). */
int mp_elem_reftensor(mp_ctx *ctx, const mp_mesh *mesh, int kind, int dir, int lr, int lc, const double *d_T0,
                      const double *d_coef, int coef_kind, double scale, double *d_elem_mat);
/* the same with ROW-ORDERED output for mp_scatter_matrix_ordered: local row i of cell c is stored at d_elem_mat[d_row_pos[c*lr + i] * lc ...]
 * (d_row_pos from mp_plan_row_positions of the block's plan; rows the plan does not assemble are skipped) */
int mp_elem_reftensor_rows(mp_ctx *ctx, const mp_mesh *mesh, int kind, int dir, int lr, int lc, const double *d_T0,
                           const double *d_coef, int coef_kind, double scale, const int32_t *d_row_pos, double *d_elem_mat);

/* quadrature tables (device): weights [nq] (sum = 1/dim!), barycentric coordinates [nq][dim+1], velocity basis values [nq][nbv] */
typedef struct {
    int32_t nq, nbv;
    const double *d_w, *d_lam, *d_vphi;
    const double *d_vdphi;   /* optional: reference gradients of the velocity basis [nq][nbv][dim] (porosity row: div u) */
} mp_quad;
/* c0 row by numerical quadrature with a Lagrange (P1/P2) lagged velocity given per component (mode R); d_vel_dofs [ncell][nbv].
 * prm->ncomp may be 1 with prm->reaction == 2 (first order Da*c) and prm->supg == 4 (tau = h/(2|w|)):
 * DarcyAdvection.advection_diffusion, mupopp.py:650-664. */
int mp_elem_p1_transport_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_transport_params *prm, const mp_quad *quad,
                              const int32_t *d_vel_dofs, const double *d_vel0, const double *d_vel1, const double *d_vel2,
                              const double *d_c0n, const double *d_c1n, const double *d_c2n, const double *d_c1new,
                              const double *d_c2new, const double *d_fsrc, double *d_elem_mat, double *d_elem_vec);

/* ------------------------------------------------------------------ compaction forms (modules/mupopp.py:125-471)
 * Coefficient functions of the P1 porosity evaluated at the points of the shared quadrature rule (degree 8: SURVEY.md appendix B):
 *   fn 0: dL*phi*phi^(1/phi^3)  (mupopp.py:365, reproduced literally)   fn 1: dL/phi  (mupopp.py:380)   fn 2: eta = 3 phi/(2(1-phi)(2-phi))  (:346-347) */
typedef struct {
    double da, R, B, theta, dL;
    double zhat[3];
} mp_compaction_params;
/* d_out[c] = cell average of fn(phi): the cellwise coefficient of the P1 stiffness blocks dL..grad(p).grad(q) */
int mp_cell_fn(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, int fn, double dL, const double *d_phi, double *d_out);
/* elem_mat[c][i][j] = scale * int fn(phi) lambda_i lambda_j   (-eta c omega, mupopp.py:366; 0.5 eta c omega, :380) */
int mp_elem_p1_mass_fn(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, int fn, double dL, double scale, const double *d_phi,
                       double *d_elem_mat);
/* L = -(h.v + da R buoy gam q/(1 - R buoy)) dx, h = dL(1-phi)(chi grad(phi)/B - R buoy zhat) + da grad(4 gam/(3 phi))  (mupopp.py:356-357,374-375);
 * evec_u<k> [ncell][nb2] per velocity component (P2 basis values in quad->d_vphi), evec_p [ncell][dim+1] */
int mp_elem_compaction_rhs(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, const mp_compaction_params *prm, const double *d_phi,
                           const double *d_gam, const double *d_buoy, double *d_evec_u0, double *d_evec_u1, double *d_evec_u2,
                           double *d_evec_p);
/* porosity row compaction.mass_conservation (mupopp.py:217-259): P1 porosity, P2 velocity (values + reference gradients in quad) */
int mp_elem_p1_porosity_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, double da, double dt, const int32_t *d_vel_dofs,
                             const double *d_vel0, const double *d_vel1, const double *d_vel2, const double *d_phi0, const double *d_gam,
                             double *d_elem_mat, double *d_elem_vec);

/* the same row with a P2 porosity space (X = CG2: benchmark/soliton/soliton.py:120, benchmark/sphere_3D_pure_shear/sphere.py:76): porosity, melting
 * rate and velocity all use the P2 dofmap d_p2_dofs[ncell][nb2]; elem_mat [ncell][nb2][nb2] (cell-ordered), elem_vec [ncell][nb2].  Includes the
 * -div(grad(phi_mid)) term of the SUPG residual (mupopp.py:250-251), which vanishes for P1 only.  One warp per cell. */
int mp_elem_p2_porosity_quad(mp_ctx *ctx, const mp_mesh *mesh, const mp_quad *quad, double da, double dt, const int32_t *d_p2_dofs,
                             const double *d_vel0, const double *d_vel1, const double *d_vel2, const double *d_phi0, const double *d_gam,
                             double *d_elem_mat, double *d_elem_vec);

/* ------------------------------------------------------------------ scalar functionals (dolfin.assemble(<scalar form>), errornorm)
 * J = sum_e weight[e] * sum_q w[q] * prod_{k<nfield} F_k(x_q) * (V(x_q) . g_e)      e = cells (dx) or boundary facets (ds(id))
 * Call sites replaced: assemble(c00*phi0*dot(u0, n)*ds(1)) run/CCS2D/CCS_3comp.py:277; assemble(c02*dx), assemble(f*Da*c00*c01*dx),
 * assemble(c00*dot(u, unitNormal)*ds(1)) run/stokes_ADR_3D/master_advective_stokes.py:232-265; errornorm(u, uh, 'L2')
 * benchmark/sphere_3D_pure_shear/sphere.py:157-159 (= sqrt of the functional of (u - uh)^2 per component).
 * F_k = scalar Lagrange field (P1/P2/P3) given by its dof vector, the cell dofmap and the basis values tabulated at the
 * quadrature points of every reference entity (table id per entity: 0 for cells, the UFC local facet index for facets).
 * V = optional vector factor: Lagrange components (vec_kind 1) or one constant vector per cell (vec_kind 2: mode F's DG0 velocity);
 * g_e = per-entity direction (facet normal) or one constant direction.  Deterministic reduction; allreduced over ranks. */
typedef struct {
    int32_t nfield;                /* 0..4 scalar factors */
    int32_t nb[4];                 /* local dofs of each factor */
    const double *d_field[4];
    const int32_t *d_dofs[4];      /* [ncell][nb[k]] */
    const double *d_tab[4];        /* [ntab][nq][nb[k]] */
    int32_t vec_kind, nbv;
    const double *d_vec[3];        /* vec_kind 1: one dof vector per component; vec_kind 2: d_vec[0] = [ncell][dim] */
    const int32_t *d_vdofs;        /* [ncell][nbv] */
    const double *d_vtab;          /* [ntab][nq][nbv] */
    int32_t nq, ntab;
    const double *d_w;             /* [nq], sum = 1 (entity measures are in d_ent_weight) */
} mp_functional;
int mp_integrate(mp_ctx *ctx, int32_t dim, const mp_functional *f, int64_t nent, const int32_t *d_ent_cell, const int32_t *d_ent_tab,
                 const double *d_ent_weight, const double *d_ent_g, const double *g_const, double *result);

/* max over cells of the cell average of |u| (core.u_max, modules/core.py:23-39; compute_dt = cfl*h_min/max(1e-6, u_max), :43-48).
 * vec_kind 1: Lagrange components d_u0..d_u2 with the cell dofmap d_vdofs[ncell][nbv] and the basis table d_vtab[nq][nbv] at the points of
 * a rule with weights d_w (sum 1); vec_kind 2: d_u0 = [ncell][dim] cellwise constant velocity (mode F).  Max over ranks. */
int mp_max_cell_speed(mp_ctx *ctx, int32_t dim, int64_t ncell, int32_t vec_kind, const double *d_u0, const double *d_u1, const double *d_u2,
                      const int32_t *d_vdofs, int32_t nbv, const double *d_vtab, int32_t nq, const double *d_w, double *result);

/* ------------------------------------------------------------------ linear algebra (PETSc Mat/Vec/KSP replacement)
 * Call sites replaced: KrylovSolver(...).solve modules/mupopp.py:289-303,410-425; the LU inside solve(a==L,..). */
typedef struct {
    int64_t nrow;          /* owned rows */
    int64_t ncol;          /* local columns (owned + ghost) */
    int64_t nnz;
    int32_t max_row_nnz;   /* longest row (0 = unknown): selects the CSR-stream SpMV for short regular rows */
    int32_t pad;
    const int32_t *d_rowptr;
    const int32_t *d_colidx;
    const double *d_vals;
    /* optional SELL-32 mirror used by SpMV (NULL -> CSR kernels): slice s = rows 32s..32s+31 stored [j][lane],
     * padded to the slice's longest row with (col = the row's first column, val = 0); see mp_plan_sell_info / mp_csr_to_sell */
    const int32_t *d_sell_sliceptr;
    const int32_t *d_sell_col;
    const double *d_sell_val;
    /* optional compressed column indices of the SELL mirror (mp_plan_sellu_info), "slice-uniform offsets": for column position j of
     * slice s, d_sell_meta[sliceptr[s]/32 + j] = (d << 1) | 1 when every row of the slice has col = row + d there, else (k << 1) with the
     * 32 explicit columns at d_sell_xcol[32 k ..].  One int32 per (slice, j) instead of 32: SpMV streams ~8.2-9 instead of 12 bytes per
     * nonzero on mesh-line numberings.  NULL -> d_sell_col is used */
    const int32_t *d_sell_meta;
    const int32_t *d_sell_xcol;
} mp_csr;

int mp_spmv(mp_ctx *ctx, const mp_csr *A, const double *d_x, double *d_y);
int mp_dot(mp_ctx *ctx, int64_t n, const double *d_x, const double *d_y, double *result);      /* host result; allreduced if a communicator is attached */
int mp_axpby(mp_ctx *ctx, int64_t n, double a, const double *d_x, double b, double *d_y);      /* y = a x + b y */
 int mp_vmul(mp_ctx *ctx, int64_t n, double a, const double *d_x, const double *d_y, double *d_z);   /* z = a * x .* y (Jacobi apply) */
int mp_jacobi_setup(mp_ctx *ctx, const mp_csr *A, double *d_dinv);                             /* dinv = 1/diag(A) */

typedef struct {
    double rtol, atol;     /* stop when ||r|| <= max(rtol*||b||, atol)  (true recurrence residual, 2-norm) */
    int32_t maxit;
    int32_t check_every;   /* host convergence poll interval (iterations); 0 -> default */
    int32_t niter;         /* out */
    int32_t converged;     /* out */
    double relres;         /* out: ||r||/||b|| */
} mp_solve_info;

/* Jacobi-preconditioned CG (SPD: mass, pressure).  d_x holds the initial guess on entry. */
int mp_solve_pcg(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                 mp_solve_info *info);
/* Chebyshev-Jacobi preconditioned CG: z = p_k(D^-1 A) D^-1 r on [lmin, lmax] (k = degree; k-1 extra SpMVs and no reductions per
 * application).  mp_spectral_bound returns the Gershgorin bound max_i dinv_i sum_j |a_ij| >= lambda_max(D^-1 A). */
int mp_spectral_bound(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, double *lmax);
int mp_solve_pcg_chebyshev(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                           int degree, double lmin, double lmax, mp_solve_info *info);

/* Line (block-tridiagonal) Jacobi preconditioned CG for operators whose rows are numbered plane by plane (structured BoxMesh /
 * RectangleMesh lattices, also z-slab partitions of them): M = diag(A) + the (i, i -+ stride) entries of A, i.e. one tridiagonal block
 * per grid line l = i mod stride.  Replaces what the reference gets from PETSc's LU / AMG on the thin boxes of
 * run/CCS/CCS_3D_top_source.py, run/LVL_3D/LVL_RII3D.py ([0,4]x[0,4]x[0,1]: the z-coupling is 16x the in-plane coupling).
 * mp_line_setup factors every block M_l = L D U (d_lo, d_invw, d_mb: nrow doubles each, caller-allocated; spd != 0: fails unless every
 * pivot is positive -- PCG needs that; spd == 0: general tridiagonal blocks for the non-symmetric transport row),
 * mp_line_apply computes z = M^-1 r (d_work: nrow doubles), mp_solve_pcg_line runs PCG with the two sweeps fused into the CG vector
 * updates, mp_solve_bicgstab_line runs right-preconditioned BiCGStab with the forward sweeps fused into the p / s updates.
 * A->nrow must be a multiple of stride. */
int mp_line_setup(mp_ctx *ctx, const mp_csr *A, int64_t stride, int spd, double *d_lo, double *d_invw, double *d_mb);
/* factor from bands already extracted (d_invw holds the diagonal, d_mb the (i, i + stride) entries on entry) */
int mp_line_factor(mp_ctx *ctx, int64_t nrow, int64_t stride, int spd, const double *d_lo, double *d_invw, double *d_mb);
int mp_line_apply(mp_ctx *ctx, int64_t nrow, int64_t stride, const double *d_lo, const double *d_invw, const double *d_mb,
                  const double *d_r, double *d_z, double *d_work);
int mp_solve_pcg_line(mp_ctx *ctx, const mp_csr *A, int64_t stride, const double *d_lo, const double *d_invw, const double *d_mb,
                      const double *d_b, double *d_x, mp_halo *halo, mp_solve_info *info);
int mp_solve_bicgstab_line(mp_ctx *ctx, const mp_csr *A, int64_t stride, const double *d_lo, const double *d_invw, const double *d_mb,
                           const double *d_b, double *d_x, mp_halo *halo, mp_solve_info *info);

/* Jacobi-preconditioned BiCGStab (non-symmetric transport rows). */
int mp_solve_bicgstab(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                      mp_solve_info *info);

/* Restarted GMRES(m), right Jacobi preconditioning, classical Gram-Schmidt applied twice; Hessenberg/Givens on the
 * device.  The robust choice for the SUPG transport rows (non-normal at large CFL).  restart_m <= 0 -> 50. */
int mp_solve_gmres(mp_ctx *ctx, const mp_csr *A, const double *d_dinv, const double *d_b, double *d_x, mp_halo *halo,
                   int restart_m, mp_solve_info *info);

/* ------------------------------------------------------------------ block operators and MINRES (saddle-point systems)
 * Replaces KrylovSolver("minres", "amg") on assemble_system(a, L, bcs) / assemble_system(b, L, bcs) of compaction.momentum_solver
 * (modules/mupopp.py:383-429), the MUMPS solve of the Stokes block (run/stokes_ADR_3D/master_advective_stokes.py:213) and the mixed Darcy
 * saddle system (modules/mupopp.py:804-807).  The operator is a sum of CSR blocks on a flat vector [block 0 | block 1 | ...]:
 *     y[rb] += scale * A x[cb]          and, with d_mask (1 free / 0 Dirichlet dof), y = M A M x + (I - M) x  (symmetric elimination). */
typedef struct {
    int32_t rb, cb;          /* row block, column block */
    const mp_csr *A;         /* nrow = size of block rb, ncol = size of block cb (CSR arrays are used; a SELL mirror is ignored) */
    double scale;
} mp_block_term;
typedef struct {
    int32_t nblk, nterm;     /* 1..8 blocks */
    int64_t off[9];          /* block offsets in the flat vector, off[nblk] = n */
    const mp_block_term *terms;
    const double *d_mask;    /* [n] or NULL */
} mp_block_op;
/* block-diagonal SPD preconditioner: per block z = p_k(D^-1 P) D^-1 v, the Chebyshev polynomial of degree k on [lmin, lmax]
 * (P == NULL or degree <= 1: Jacobi).  d_dinv[n] = inverse diagonal of the preconditioner blocks (1 on Dirichlet dofs). */
typedef struct {
    const mp_csr *P[8];
    int32_t degree[8];
    double lmin[8], lmax[8];
    const double *d_dinv;
} mp_block_prec;
int mp_block_apply(mp_ctx *ctx, const mp_block_op *op, const double *d_x, double *d_y);
/* preconditioned MINRES, all recurrence scalars on the device (no per-iteration host synchronisation); stops when the
 * preconditioned residual norm is <= max(rtol * ||b||_{P^-1}, atol), then re-evaluates the TRUE residual and restarts if it drifted.
 * d_x: initial guess in (Dirichlet dofs must already hold their values), solution out. */
int mp_solve_minres(mp_ctx *ctx, const mp_block_op *op, const mp_block_prec *prec, const double *d_b, double *d_x, mp_solve_info *info);
/* Mixed Darcy saddle system of DarcyAdvection.darcy_bilinear (modules/mupopp.py:595-625) and of the Darcy rows of the multi-component
 * forms (:804-807, :1201-1204) for a constant K, symmetrised by scaling the continuity row with -K:
 *     [ M_a      -K B_a^T ] [u_a]   [f_a]          M_a = phi * P2 mass (per velocity component), B_a = (q, d_a u)
 *     [ -K B_a      0     ] [ p ] = [ 0 ]
 * solved by mp_solve_minres with the block preconditioner diag(Chebyshev(M), Chebyshev(Sp)), Sp = the P1 operator (K^2/phi) grad.grad
 * that is spectrally equivalent to the pressure Schur complement B M^-1 K^2 B^T.
 * d_x = [u_0 .. u_{dim-1} | p]; d_mask / d_dinv as in mp_block_op / mp_block_prec (length dim*n2 + n1). */
int mp_schur_darcy_solve(mp_ctx *ctx, int32_t dim, const mp_csr *M, const mp_csr *const *BT, const mp_csr *const *B, const mp_csr *Sp,
                         double K, const double *d_mask, const double *d_dinv, int32_t deg_u, double lmin_u, double lmax_u, int32_t deg_p,
                         double lmin_p, double lmax_p, const double *d_b, double *d_x, mp_solve_info *info);

/* ------------------------------------------------------------------ multi-GPU (replaces PETSc VecScatter / MPI_Allreduce)
 * One process per GPU.  The NCCL unique id is created on rank 0 and broadcast by the host (torch.distributed). */
int mp_nccl_unique_id(uint8_t *out128);
int mp_comm_init(mp_ctx *ctx, int nranks, int rank, const uint8_t *id128);
int mp_comm_destroy(mp_ctx *ctx);
int mp_allreduce_sum(mp_ctx *ctx, double *d_buf, int count);
/* halo plan: for each neighbour k, send x[d_send_idx[send_off[k] .. send_off[k+1])] and receive into
 * x[recv_start[k] .. recv_start[k]+recv_count[k])  (ghost entries are contiguous per neighbour). */
int mp_halo_create(mp_ctx *ctx, int nneigh, const int32_t *neigh_rank, const int32_t *send_off, const int32_t *d_send_idx,
                   const int32_t *recv_start, const int32_t *recv_count, mp_halo **out);
int mp_halo_exchange(mp_ctx *ctx, mp_halo *halo, double *d_x);
int mp_halo_destroy(mp_halo *halo);

/* Peer-memory collectives over NVLink / NVSwitch (CUDA IPC), the fast path of the Krylov loops on N > 1 GPUs of one node:
 *  - the allreduce of the Krylov scalars happens INSIDE the reducing kernel (mailbox windows, sequence-number flags, rank-ordered
 *    sum => identical bits on every rank): no ncclAllReduce launch per inner product;
 *  - the halo exchange is fused into the SELL-32 SpMV kernel: interface values are stored straight into the neighbours' ghost
 *    windows, interior slices are multiplied while the stores are in flight, boundary slices after the neighbours' flags arrived.
 * Set-up (host): every rank exports a 64-byte IPC handle, the host all-gathers them (torch.distributed) and attaches.
 * NCCL stays the transport of everything outside the Krylov loops (coefficient halos before assembly) and the fallback when
 * the windows are not attached; option "peer" (mp_ctx_set_option) switches between the two for A/B measurements. */
int mp_comm_peer_export(mp_ctx *ctx, uint8_t *handle64);
int mp_comm_peer_attach(mp_ctx *ctx, const uint8_t *handles /* [nranks][64], indexed by rank */);
int mp_comm_peer_enabled(const mp_ctx *ctx);
/* ghost window of one halo plan: nown owned columns, nghost ghost columns (local columns nown .. nown+nghost-1) */
int mp_halo_peer_export(mp_ctx *ctx, mp_halo *halo, int64_t nown, int64_t nghost, uint8_t *handle64);
/* per neighbour k (order of mp_halo_create): its window handle, the offset of MY block inside its ghost window
 * (its recv_start for me minus its nown), its nghost, and my index in ITS neighbour list */
int mp_halo_peer_attach(mp_ctx *ctx, mp_halo *halo, const uint8_t *neigh_handles /* [nneigh][64] */, const int64_t *peer_goff,
                        const int64_t *peer_gcap, const int32_t *peer_slot);

#ifdef __cplusplus
}
#endif
#endif
