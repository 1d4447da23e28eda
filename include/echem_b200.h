/* echem_b200 -- C ABI of the B200 Newton-step library (libechem_b200.so).
 *
 * Drop-in boundary.  In the reference the hot path sits behind
 *     NonlinearVariationalProblem(Form, u, bcs) / NonlinearVariationalSolver(...).solve()
 * (echemfem/solver.py:714-718, 725) and, below that, behind the two PETSc SNES
 * callbacks  F(X) -> Vec  and  J(X) -> Mat(AIJ)  plus KSPSolve.  The entry
 * points below are what a binding for that path calls instead; each one cites
 * the reference-side operation it replaces.  Python binds them with ctypes
 * (echemfem_b200/capi.py; INTEGRATION.md shows the stub a maintainer would add
 * to the reference).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - The caller owns every buffer (state, matrix values, Krylov basis); the
 *    handle owns only plans.  No hot call allocates.
 *  - All work is enqueued on the caller's stream (a cudaStream_t passed as
 *    void*); only calls documented as synchronising read back a scalar.
 *  - Every function returns 0 on success, a negative ech_* code or a positive
 *    cudaError_t otherwise; ech_last_error() describes the last failure.
 *    Nothing throws across this boundary.  One handle per (GPU, stream); not
 *    re-entrant.
 *  - FP64 throughout; column indices int32, row pointers int64.
 */
#ifndef ECHEM_B200_H
#define ECHEM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECH_OK 0
#define ECH_ERR_ARG (-1)
#define ECH_ERR_MODULE (-2)
#define ECH_ERR_MISMATCH (-3)
#define ECH_ERR_BREAKDOWN (-4)
#define ECH_ERR_NOT_CONVERGED (-5)

typedef struct ech_handle ech_handle;

/* Mesh / discretisation plan: per-cell connectivity of a DG tensor-product space
 * (what Firedrake's cell_node_map + interior/exterior facet maps + coordinates
 * provide below solver.py:714). */
typedef struct ech_mesh_desc {
    int32_t dim, degree, nfields, naux;
    int64_t ncell;            /* owned cells: rows are produced for these            */
    int64_t ncell_total;      /* owned + ghost cells: state / geometry cover these   */
    const double *xcell;      /* [ncell_total][2^dim][dim]                            */
    const int32_t *nbr;       /* [ncell][2 dim]  neighbour or -(1 + boundary class)   */
    const uint8_t *code;      /* [ncell][2 dim]  neighbour facet | orientation << 3   */
    const uint8_t *slot;      /* [ncell][2 dim]                                       */
    const uint8_t *slot_self; /* [ncell]                                              */
    const uint8_t *ncol;      /* [ncell]                                              */
    const double *cell_volume;/* [ncell_total]  CellVolume  (solver.py:1491)          */
    const double *cell_diam;  /* [ncell_total]  CellDiameter (solver.py:1657)         */
    const double *facet_area; /* [ncell][2 dim] FacetArea   (solver.py:1491)          */
    /* continuous (CG) spaces: cell -> node map and its transpose (NULL / 0 for DG);
     * the CSR pattern of a CG space is passed to ech_jacobian by the caller */
    int64_t nnodes;
    const int32_t *cell_nodes;/* [ncell][n]                                            */
    const int64_t *inc_ptr;   /* [nnodes+1]                                            */
    const int32_t *inc_cell;  /* [ninc]                                                */
    const uint8_t *inc_loc;   /* [ninc]                                                */
    const uint8_t *colslot;   /* [ninc][n]                                             */
    const int32_t *adj_count; /* [nnodes]                                              */
} ech_mesh_desc;

/* -- lifetime ---------------------------------------------------------------- */
/* Binds a compiled physics module (the weak form: replaces Form + derivative(F,u) +
 * the TSFC kernels of NonlinearVariationalProblem, solver.py:714) to a mesh plan. */
int ech_create(const char *module_path, const ech_mesh_desc *mesh, ech_handle **out);
void ech_destroy(ech_handle *h);
const char *ech_last_error(const ech_handle *h);
/* d, p, nf, naux, nconst, nq, nclasses, family, own[nf*nf], nbr[nf*nf] */
int ech_describe(const ech_handle *h, int32_t *info, int capacity);

/* CellVolume / FacetArea / CellDiameter of every cell (the UFL geometric quantities of
 * solver.py:1491, 1657) from the per-cell vertex coordinates [ncell][2^dim][dim]. */
int ech_cell_geometry(int32_t dim, int64_t ncell, const double *xcell, double *vol, double *area, double *diam,
                      void *stream);

/* -- sparsity (PyOP2 Sparsity + MatSeqAIJ preallocation; "mat_type": "aij", solver.py:943) */
int ech_num_rows(const ech_handle *h, int64_t *nrows);
/* fills rowptr[nrows+1]; *nnz_host receives the total. Synchronises the stream. */
int ech_pattern_rowptr(ech_handle *h, int64_t *rowptr, int64_t *nnz_host, void *stream);
int ech_pattern_colidx(ech_handle *h, const int64_t *rowptr, int32_t *colidx, void *stream);

/* -- SNES callbacks --------------------------------------------------------------- */
/* form_function: F = assemble(Form) at state u (SNESFunctionEval; solver.py:725). */
int ech_residual(ech_handle *h, const double *u, const double *aux, const double *consts, double *F,
                 void *stream);
/* form_jacobian: CSR values of assemble(derivative(Form, u)) (SNESJacobianEval + MatSetValuesLocal). */
int ech_jacobian(ech_handle *h, const double *u, const double *aux, const double *consts,
                 const int64_t *rowptr, double *vals, void *stream);
/* Both callbacks at the same state in one pass (a Newton iteration needs F(u_k) and J(u_k)): geometry,
 * bases, state interpolation and staging are shared; for DG the two are one kernel. */
int ech_residual_jacobian(ech_handle *h, const double *u, const double *aux, const double *consts,
                          const int64_t *rowptr, double *vals, double *F, void *stream);
/* Rows flagged in row_mask[nrows] become identity rows and F is zeroed there: strong
 * DirichletBC rows (solver.py:1531, 2009) and the frozen concentration rows of the potential
 * pre-solve (solver.py:645-662). F and/or vals may be NULL. */
int ech_constrain_rows(int64_t nrows, const uint8_t *row_mask, const int64_t *rowptr, const int32_t *colidx,
                       double *vals, double *F, void *stream);

/* -- linear algebra (PETSc MatMult / Vec ops / PC bjacobi + sub ilu(0) / KSP gmres; solver.py:1179-1189) */
int ech_spmv(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *x,
             double *y, void *stream);
/* y = a x + b y */
int ech_axpby(int64_t n, double a, const double *x, double b, double *y, void *stream);
/* *result_host = x . y ; synchronises. `work` holds >= 1024 doubles. */
int ech_dot(int64_t n, const double *x, const double *y, double *work, double *result_host, void *stream);

/* Building blocks of the multi-GPU Krylov loop (results stay on the device so that the caller can
 * all-reduce them over NCCL before reading them; PETSc: VecMDot / VecMAXPY + MPI_Allreduce):
 * out[j] = V_j . w for j < k and out[k] = w . w, V column-major with leading dimension n;
 * work holds >= (k + 1) * 1024 doubles.   w += sign * sum_j h[j] V_j. */
int ech_multidot(int64_t n, int32_t k, const double *V, const double *w, double *work, double *out, void *stream);
int ech_multiaxpy(int64_t n, int32_t k, const double *V, const double *h, double sign, double *w, void *stream);
/* w += sign * sum_{j<k} h[j] V[j], and of the UPDATED w: out[j] = V[j].w (j < k), out[k] = w.w -- the update of one
 * Gram-Schmidt pass fused with the projection of the next (refinement) pass; 1 <= k <= 64.  h and out may be the same
 * buffer; work as for ech_multidot. */
int ech_multiaxpy_dot(int64_t n, int32_t k, const double *V, const double *h, double sign, double *w, double *work,
                      double *out, void *stream);

/* Block-Jacobi with ILU(0) inside each block.  Block b owns the cells
 * [block_cell_ptr[b], block_cell_ptr[b+1]) (all fields of those cells); couplings
 * between blocks are dropped.  nblocks = 1 is PETSc's default "one block per rank". */
int ech_bjacobi_ilu0_factor(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                            int64_t nblocks, const int64_t *block_cell_ptr, const int64_t *rowptr,
                            const int32_t *colidx, const double *vals, double *lu, int32_t *diag_pos,
                            void *stream);
/* The same without host synchronisation: `nnz` and a device flag word from the caller (0 ok, 1 zero pivot, 2 shifted
 * pivot; pass it to ech_bjacobi_ilu0_factor_status after joining the stream).  Lets the levels of a multigrid
 * hierarchy factor concurrently on separate streams. */
int ech_bjacobi_ilu0_factor_async(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc, int64_t nblocks,
                                  const int64_t *block_cell_ptr, const int64_t *rowptr, const int32_t *colidx,
                                  const double *vals, int64_t nnz, double *lu, int32_t *diag_pos, int32_t *flag,
                                  void *stream);
int ech_bjacobi_ilu0_factor_status(int32_t flag);
/* The factorisation on the plan's local column indices (`lcol` of ech_bjacobi_ilu0_plan_count): work row in shared
 * memory, no searches; same factors as ech_bjacobi_ilu0_factor.  Asynchronous like ..._factor_async. */
int ech_bjacobi_ilu0_factor_planned(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                                    int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                                    const int64_t *rowptr, const int32_t *colidx, const double *vals, int64_t nnz,
                                    const uint16_t *lcol, double *lu, int32_t *diag_pos, int32_t *flag, void *stream);
int ech_bjacobi_ilu0_apply(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                           int64_t nblocks, const int64_t *block_cell_ptr, const int64_t *rowptr,
                           const int32_t *colidx, const double *lu, const int32_t *diag_pos, const double *r,
                           double *z, void *stream);

/* Level-scheduled triangular solves (same factors, same result up to the summation order inside a row).  The PLAN
 * depends on the pattern and the blocks only and is built once: plan_count computes the dependency levels of
 * both sweeps, the block-local column table lcol[nnz] and the passes per block (passptr[2][nblocks+1], scanned);
 * *npass_host receives the total, the caller allocates slots[npass * (32 / lanes_per_row)] (8 bytes each) and
 * plan_fill sorts the rows into them.  levels: scratch of 2 * nrows uint16.  max_block_rows: rows (all fields)
 * of the largest block, at most 16384.  lanes_per_row in {2, 4, 8, 16}: lanes that share one row.
 * plan_count synchronises the stream. */
int ech_bjacobi_ilu0_plan_count(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                                int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                                const int64_t *rowptr, const int32_t *colidx, int32_t lanes_per_row, uint16_t *lcol,
                                uint16_t *levels, int64_t *passptr, int64_t *npass_host, void *stream);
int ech_bjacobi_ilu0_plan_fill(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                               int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                               const int64_t *rowptr, const int32_t *colidx, int32_t lanes_per_row,
                               const uint16_t *levels, const int64_t *passptr, uint64_t *slots, void *stream);
int ech_bjacobi_ilu0_apply_planned(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                                   int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                                   const int64_t *rowptr, const double *lu, const uint16_t *lcol,
                                   const int64_t *passptr, const uint64_t *slots, int32_t lanes_per_row,
                                   const double *r, double *z, void *stream);

/* Pass-packed variant of the level-scheduled solves: after every factorisation the factors are copied into the
 * order the solve consumes them -- one contiguous 16-byte aligned chunk per pass (header, values, block-local column
 * positions; entries whose column lies outside the block are dropped) -- and the solve fetches chunk t+4 with ONE cp.async.bulk (TMA bulk copy, mbarrier completion) into a
 * shared-memory ring while chunk t is reduced.  pack_count (after plan_count: it reads its `levels` and `lcol`)
 * returns totals_host = {bytes of `packed`, entries of `src`, passes, largest chunk}; pack_fill writes the static
 * parts (sizes: scratch of 6 * (nblocks + 1) int64 shared by both calls; bdesc: 2 * nblocks * 8 int64);
 * pack_values refreshes the values from `lu`; apply_packed solves.  pack_count synchronises the stream. */
int ech_bjacobi_ilu0_pack_count(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                                int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                                const int64_t *rowptr, const int32_t *colidx, int32_t lanes_per_row,
                                const uint16_t *levels, const uint16_t *lcol, int64_t *sizes, int64_t *totals_host,
                                void *stream);
int ech_bjacobi_ilu0_pack_fill(int64_t nrows, int32_t nfields, int64_t rows_per_field, int32_t nloc,
                               int64_t nblocks, const int64_t *block_cell_ptr, int32_t max_block_rows,
                               const int64_t *rowptr, const int32_t *colidx, int32_t lanes_per_row,
                               const uint16_t *levels, const uint16_t *lcol, int64_t *sizes, uint8_t *packed,
                               uint32_t *src, int64_t *bdesc, void *stream);
int ech_bjacobi_ilu0_pack_values(int64_t nblocks, int32_t lanes_per_row, const int64_t *bdesc, uint8_t *packed,
                                 const uint32_t *src, const double *lu, void *stream);
int ech_bjacobi_ilu0_apply_packed(int32_t nfields, int64_t rows_per_field, int32_t nloc, int64_t nblocks,
                                  const int64_t *block_cell_ptr, int32_t max_block_rows, int32_t max_chunk_bytes,
                                  int32_t lanes_per_row, const int64_t *bdesc, const uint8_t *packed, const double *r,
                                  double *z, void *stream);

/* Coarse space of the optional two-level preconditioner (not in the reference's "ilu" option set; it
 * plays the role of the coarse levels of its gmg/amg sets, solver.py:1190-1213, for meshes on which
 * block-Jacobi/ILU(0) alone stalls).  Piecewise-constant aggregation agg[dof] -> coarse dof:
 * Ac[nc*nc] = P^T A P (dense, row-major);  rc = P^T v;  z += P yc. */
int ech_coarse_galerkin(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals,
                        const int32_t *agg, int32_t nc, double *Ac, void *stream);
int ech_coarse_restrict(int64_t n, const int32_t *agg, int32_t nc, const double *v, double *rc, void *stream);
int ech_coarse_prolong_add(int64_t n, const int32_t *agg, const double *yc, double *z, void *stream);
/* y[nr] = M[nr x nc] x (dense, row-major; one warp per row) */
int ech_dense_gemv(int32_t nr, int32_t nc, const double *M, const double *x, double *y, void *stream);

/* Ghost-dof exchange, device side (VecScatter / PetscSF pack and unpack of the reference's halo updates): the send
 * lists of ALL peers are one index array, their messages contiguous slices of one buffer; the transport between
 * the slices is the caller's (NCCL send / recv through torch.distributed). */
int ech_halo_pack(int64_t n, const int64_t *idx, const double *u, double *buf, void *stream);
int ech_halo_unpack(int64_t n, const int64_t *idx, const double *buf, double *u, void *stream);

/* Transfers of the nested DG hierarchy from the tables of ech_galerkin_dg (PCMG's interpolation / restriction on the
 * reference's MeshHierarchy, solver.py:1190-1201): z_f[f][c*nloc + a] += sum_A pblk[c][a][A] zc[f][parent[c]*nloc + A],
 * and rc[f][K*nloc + A] = sum over the children c of K (children[K*nchild + k], -1: none) and a of
 * pblk[c][a][A] r[f][c*nloc + a]; fs_f / fs_c: entries per field of the fine / coarse vectors. */
int ech_dg_prolong_add(int32_t nloc, int32_t nf, int64_t ncell_f, int64_t fs_f, int64_t fs_c, const int32_t *parent,
                       const double *pblk, const double *zc, double *z, void *stream);
int ech_dg_restrict(int32_t nloc, int32_t nf, int32_t nchild, int64_t ncell_c, int64_t fs_f, int64_t fs_c,
                    const int32_t *children, const double *pblk, const double *r, double *rc, void *stream);

/* z[r] += sum_v W[a][v] zc[vtab[cell*nv + v]] for the DG rows r = cell*nv + a of one field (cells with vtab < 0 are
 * skipped): prolongation from the conforming auxiliary space, fused with the update. */
int ech_aux_prolong_add(int32_t nv, int64_t nrow, const int64_t *vtab, const double *W, const double *zc, double *z,
                        void *stream);

/* A_aux = Q^T A Q for the continuous-Q1 subspace of a DG-Q1 space (the conforming auxiliary space of the potential
 * fields in the multigrid preconditioner; the reference reaches conforming coarse spaces through its `gmg` / `amg`
 * option sets, solver.py:1190-1213).  Q on one cell is the dense nv x nv matrix W (row-major [dg node][vertex]) for
 * every cell.  Per cell block k of A: row0[k] = first of its nv rows, off[k] = offset of the block from each row
 * start, dst[k*nv*nv + i*nv + j] = position in A_aux of (vertex i of the row cell, vertex j of the column cell).
 * out[0..nout) is zeroed first. */
int ech_aux_galerkin(int32_t nv, int64_t nblk, const int64_t *row0, const int32_t *off, const int32_t *dst,
                     const int64_t *rowptr, const double *vals, const double *W, int64_t nout, double *out,
                     void *stream);

/* E^T A E for the conforming space on a coarser GLOBAL vertex grid shared by all ranks (the coarse space of the
 * multi-GPU preconditioner; complete owned rows, ghost columns included).  Wc: one nv x nv embedding block per local
 * cell ([dg node][coarse corner]); cI: coarse cell multi-index per local cell (int32[d]); vd: coarse vertices per
 * direction; fidx[nfields]: position of a field among the `na` auxiliary fields or -1; N: rows per field.  Output in
 * ELL layout, zeroed first: out[((ia * nvert + I) * na + ja) * 5^d + s], s = lexicographic offset J - I + 2 per
 * direction. */
int ech_aux_galerkin_global(int32_t nv, int32_t d, int64_t nblk, const int64_t *row0, const int32_t *off,
                            const int64_t *rowptr, const int32_t *colidx, const double *vals, int64_t N,
                            const int32_t *fidx, int32_t na, const double *Wc, const int32_t *cI, const int32_t *vd,
                            int64_t nout, double *out, void *stream);

/* Galerkin coarse operator A_c = P^T A P of the DG tensor hierarchy (multigrid set-up; stands in for the coarse-level
 * assembly of the reference's `gmg` option set on a MeshHierarchy, solver.py:1190-1201).  A and A_c share the DG
 * block-stencil layout; P is one dense nloc x nloc block per fine cell.  ip[8] = {nfields, nloc (2, 4 or 8), facets
 * per cell, children per coarse cell, fine cells, coarse cells, fine rows per field, coarse rows per field};
 * pp[17] = {rowptr_f, vals_f, nbr_f, slot_f, slot_self_f, ncol_f, parent_f, pblk, children, rowptr_c, vals_c, nbr_c,
 * slot_c, slot_self_c, ncol_c, own_mask, nbr_mask} (device pointers; masks uint8[nf*nf]).  Writes vals_c. */
int ech_galerkin_dg(const int64_t *ip, const void *const *pp, void *stream);

typedef struct ech_gmres_opts {
    double rtol, atol;
    int32_t restart, max_it;
    int32_t precond;          /* 0 none, 1 block-Jacobi/ILU(0) (factor must be current),
                                 2 = 1 + additive coarse correction, 3 = 1 + multiplicative coarse correction */
    int32_t monitor;          /* print residual norms */
    /* preconditioner plan (precond == 1) */
    int32_t nfields, nloc;
    int64_t rows_per_field, nblocks;
    const int64_t *block_cell_ptr;
    const double *lu;
    const int32_t *diag_pos;
    /* coarse space (precond >= 2): aggregation map, size, dense inverse of P^T A P (row-major), scratch */
    const int32_t *agg;
    int32_t nc, reserved;
    const double *coarse_inv;
    double *coarse_work;      /* 2 * nc (+ nrows for precond == 3) doubles */
    /* level-scheduled solves (ech_bjacobi_ilu0_plan_*): used when slots != NULL */
    const uint16_t *lcol;
    const int64_t *passptr;
    const uint64_t *slots;
    int32_t lanes_per_row, max_block_rows;
    /* Gram-Schmidt refinement (PETSc KSPGMRESCGSRefinementType): 0 never (PETSc's default), 1 if needed, 2 always */
    int32_t refine, reserved2;
} ech_gmres_opts;

/* Right-preconditioned restarted GMRES (classical Gram-Schmidt, unpreconditioned residual norm: PETSc fgmres +
 * fixed PC, solver.py:1181-1189).  basis: (restart+1) * nrows doubles, work: 3 * nrows + 4096 +
 * 2*(restart+2)*1024 doubles.  ONE device -> host read-back per iteration: the fused reduction returns the
 * Hessenberg column and w.w, from which the norm of the new basis vector follows; convergence claimed by the
 * recurrence is confirmed with the true residual.  its_host / rnorm_host receive iteration count and final norm. */
int ech_gmres(int64_t nrows, const int64_t *rowptr, const int32_t *colidx, const double *vals, const double *b,
              double *x, const ech_gmres_opts *opts, double *basis, double *work, int32_t *its_host,
              double *rnorm_host, void *stream);

/* -- host-buffer boundary: what a binding inside the reference's SNES callbacks would call ------- */
/* u_host -> device, residual (and Jacobian values into the device-resident matrix), F -> F_host.
 * vals stays on the device (the linear solve consumes it there). Synchronises.  u_host / F_host should be
 * page-locked (cudaHostAlloc / cudaHostRegister) for the copies to overlap the kernels. */
int ech_assemble_host(ech_handle *h, const double *u_host, double *u_dev, const double *aux,
                      const double *consts, const int64_t *rowptr, double *vals, double *F_dev,
                      double *F_host, void *stream);

/* Chunking of ech_assemble_host for discontinuous spaces (default 16): the owned cells are cut into nchunks
 * contiguous ranges; the state of range k+1 is copied in while range k is assembled and the residual of
 * range k-1 is copied out (three streams).  0 or 1: one copy in, one launch, one copy out.  Results are
 * bit-identical either way.  nchunks < 0: AUTO -- the first six calls time the (-nchunks)-chunk pipeline and the
 * single-copy path on the caller's buffers and the faster one is kept (several GPUs sharing one host bridge favour
 * the single copy). */
int ech_set_host_chunks(ech_handle *h, int nchunks);

/* number of kernels launched by this library (and the bound modules) in this PROCESS since it was loaded -- one
 * counter for all handles; callers take differences around the region they count */
int64_t ech_launch_count(const ech_handle *h);

#ifdef __cplusplus
}
#endif
#endif
