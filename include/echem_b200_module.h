/* ABI between libechem_b200.so and a compiled physics module (libech_phys_<key>.so).
 *
 * A physics module is the generated pointwise-integrand header (one per weak
 * form, see echemfem_b200/compiler.py) compiled together with the hand-written
 * kernel skeleton csrc/ech_dg.cuh.  It plays the role the TSFC-generated element
 * kernels + PyOP2 wrapper loops play below the reference's
 * NonlinearVariationalSolver (echemfem/solver.py:714-718).
 *
 * All pointers are DEVICE pointers.  Nothing here allocates.
 */
#ifndef ECHEM_B200_MODULE_H
#define ECHEM_B200_MODULE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ech_dg_args {
    int64_t ncell;            /* cells owned by this rank (rows are assembled for these) */
    int64_t ncell_total;      /* owned + ghost cells (state/geometry arrays cover these)  */
    const double *xcell;      /* [ncell_total][2^d][d] vertex coordinates per cell        */
    const int32_t *nbr;       /* [ncell][2d]  neighbour cell, or -(1+class) on the boundary */
    const uint8_t *code;      /* [ncell][2d]  neighbour local facet | orientation << 3     */
    const uint8_t *slot;      /* [ncell][2d]  column-block rank of the neighbour in a row   */
    const uint8_t *slot_self; /* [ncell]                                                    */
    const uint8_t *ncol;      /* [ncell]      number of cell blocks per coupled field       */
    const double *cell_volume;/* [ncell_total]                                              */
    const double *cell_diam;  /* [ncell_total]                                              */
    const double *facet_area; /* [ncell][2d]                                                */
    const double *u;          /* [nf][ncell_total * n] state, field-blocked                 */
    const double *aux;        /* [naux][ncell_total * n] data fields                        */
    const double *consts;     /* [nconst] runtime constants                                 */
    const int64_t *rowptr;    /* [nf * ncell * n + 1] (Jacobian only)                       */
    double *out;              /* residual [nf][ncell_total*n] (DG) / [nf][nnodes] (CG), or CSR values [nnz] */
    double *out2;             /* Jacobian call only: if non-NULL, the residual at the same state is written
                                 here by the same kernel (fused evaluation); NULL otherwise               */
    /* continuous spaces only (family CG): dofs are shared between cells, assembly gathers per node */
    int64_t nnodes;           /* nodes of the scalar space; field stride of u / aux / out    */
    const int32_t *cell_nodes;/* [ncell][n]  cell -> node map                                 */
    const int64_t *inc_ptr;   /* [nnodes+1]  node -> incident (cell, local node) pairs         */
    const int32_t *inc_cell;  /* [ninc]                                                        */
    const uint8_t *inc_loc;   /* [ninc]      local index of the node in that cell              */
    const uint8_t *colslot;   /* [ninc][n]   position of the cell's b-th node in the row's sorted column nodes */
    const int32_t *adj_count; /* [nnodes]    number of column nodes of a row                   */
    /* cell range of a launch: rows are assembled for cells [cell_begin, ncell).  0 for a whole-mesh launch;
     * non-zero only in the chunked host-buffer pipeline (ech_assemble_host), a multiple of
     * ech_mod_range_align() */
    int64_t cell_begin;
} ech_dg_args;

/* info[0..7] = d, p, nf, naux, nconst, nq, nclasses, family(1=DG); then nf*nf own-block flags,
 * nf*nf neighbour-block flags. Returns the number of ints written. */
#define ECH_MOD_EXPORT __attribute__((visibility("default")))
ECH_MOD_EXPORT int ech_mod_describe(int32_t *info, int capacity);
ECH_MOD_EXPORT int ech_mod_residual(const ech_dg_args *a, void *stream);
ECH_MOD_EXPORT int ech_mod_jacobian(const ech_dg_args *a, void *stream);
/* 0: cooperative kernels (default), 1: row-per-thread kernels */
ECH_MOD_EXPORT void ech_mod_set_variant(int variant);
/* granularity (in cells) of ech_dg_args.cell_begin for residual and Jacobian launches; 0 if this module's
 * kernels only do whole-mesh launches */
ECH_MOD_EXPORT int ech_mod_range_align(void);
/* number of kernel launches one residual / Jacobian call makes */
ECH_MOD_EXPORT int ech_mod_launches(int which);

#ifdef __cplusplus
}
#endif
#endif
