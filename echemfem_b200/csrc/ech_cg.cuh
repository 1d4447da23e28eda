// CG (continuous tensor-product Lagrange) residual / Jacobian assembly for sm_100a.
//
// Reference path: the same SNES function / Jacobian evaluation as ech_dg.cuh (echemfem/solver.py:714-725)
// for `family="CG"` (solver.py:1456-1459): cell integrals (incl. SUPG, solver.py:1655-1668) and
// exterior-facet integrals; no interior-facet terms; Dirichlet conditions are imposed strongly by the
// caller (ech_constrain_rows), as the reference does with DirichletBC (solver.py:1531, 2009).
//
// Same owner-computes idea as the DG kernels, transposed to shared dofs: the owner of a CSR row is the
// NODE.  A thread gathers, over the cells incident to its node, the row contributions of those cells
// (cell integral + exterior facets), so every CSR value is still written exactly once with a plain
// store -- no atomics and no element colouring (the north star's "colored or warp-aggregated
// scatter-add" becomes unnecessary).  The price is that a cell's quadrature work is repeated by each
// of its nodes' rows; sharing it per cell (as ech_dg_coop.cuh does) is the next optimisation.
//
// Row accumulators live in shared memory because the column slot of a trial node is data dependent.
#pragma once
#include "ech_dg.cuh"

namespace echcg {
using namespace echgen;
using namespace echdg;

constexpr int MAXADJ = ipow(2 * P + 1, D);      // column nodes of a row: the (2p+1)^d stencil
constexpr int CG_BLOCK = 64;

struct CellData {
    double X[NV][D];
    double U[NF][NLOC];
    double AX[ECH_NA1][NLOC];
};

__device__ __forceinline__ void gather_cell(const ech_dg_args &A, int64_t c, CellData &cd, int (&nodes)[NLOC]) {
    load_cell_coords(A.xcell, c, cd.X);
#pragma unroll
    for (int b = 0; b < NLOC; ++b) nodes[b] = A.cell_nodes[c * NLOC + b];
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int b = 0; b < NLOC; ++b) cd.U[j][b] = __ldg(A.u + j * A.nnodes + nodes[b]);
    if constexpr (NAUX > 0) {
#pragma unroll
        for (int j = 0; j < NAUX; ++j)
#pragma unroll
            for (int b = 0; b < NLOC; ++b) cd.AX[j][b] = __ldg(A.aux + j * A.nnodes + nodes[b]);
    }
}

// ---------------------------------------------------------------- residual: one thread per node
__global__ void __launch_bounds__(CG_BLOCK) cg_residual_kernel(const ech_dg_args A) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.nnodes) return;
    double res[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) res[i] = 0.0;
    Pt p;
    p.K = A.consts;
    for (int64_t e = A.inc_ptr[r]; e < A.inc_ptr[r + 1]; ++e) {
        const int64_t c = A.inc_cell[e];
        const int a = A.inc_loc[e];
        CellData cd;
        int nodes[NLOC];
        gather_cell(A, c, cd, nodes);
        p.cv = A.cell_volume[c];
        p.cd = A.cell_diam[c];
        double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
        Geo g;
        Res o;
        for (int q = 0; q < NQC; ++q) {
            int idx[D];
            double w;
            cell_point(q, idx, w);
            geometry(cd.X, idx, g);
            basis(idx, g.invJ, phi, gphi);
            interpolate<NF>(cd.U, phi, gphi, p.u, p.gu);
            if constexpr (NAUX > 0) interpolate<ECH_NA1>(cd.AX, phi, gphi, p.a, p.ga);
#pragma unroll
            for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
            cell_res(p, o);
            select_test(a, phi, gphi, phia, gphia);
            accum_res<MC>(o, w * fabs(g.detJ), phia, gphia, res);
        }
        for (int lf = 0; lf < NLF; ++lf) {
            const int nb = A.nbr[c * NLF + lf];
            if (nb >= 0) continue;
            const int cls = -nb - 1;
            p.fa = A.facet_area[c * NLF + lf];
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], tt[D > 1 ? D - 1 : 1];
                double w, ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(cd.X, idx, g);
                basis(idx, g.invJ, phi, gphi);
                interpolate<NF>(cd.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cd.AX, phi, gphi, p.a, p.ga);
                facet_normal(g, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                efacet_res(cls, p, o);
                select_test(a, phi, gphi, phia, gphia);
                accum_res<ME>(o, w * ds, phia, gphia, res);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) A.out[i * A.nnodes + r] = res[i];
}

// ---------------------------------------------------------------- Jacobian: one thread per row (I, node)
// shared accumulators: acc[(j * MAXADJ + slot) * CG_BLOCK + tid]
template <class M, int I>
__device__ __forceinline__ void accum_row_slots(const RowJ &rj, double w, double phia, const double (&gphia)[D],
                                                const double (&phi)[NLOC], const double (&gphi)[NLOC][D],
                                                const int (&slots)[NLOC], double *acc) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool h0 = M::G0[I][j], h1 = M::G1[I][j], h2 = M::G2[I][j];
        constexpr int h3 = M::G3[I][j];
        if constexpr (h0 || h1 || h2 || h3 != 0) {
            double Aj = 0.0, Bj[D];
#pragma unroll
            for (int l = 0; l < D; ++l) Bj[l] = 0.0;
            if constexpr (h0) Aj += rj.g0[j] * phia;
            if constexpr (h2) {
#pragma unroll
                for (int k = 0; k < D; ++k) Aj += rj.g2[j][k] * gphia[k];
            }
            if constexpr (h1) {
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += rj.g1[j][l] * phia;
            }
            if constexpr (h3 == 1) {
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += rj.g3[j][0][0] * gphia[l];
            }
            if constexpr (h3 == 2) {
#pragma unroll
                for (int l = 0; l < D; ++l)
#pragma unroll
                    for (int k = 0; k < D; ++k) Bj[l] += rj.g3[j][k][l] * gphia[k];
            }
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                double s = 0.0;
                if constexpr (h0 || h2) s += Aj * phi[b];
                if constexpr (h1 || h3 != 0) {
#pragma unroll
                    for (int l = 0; l < D; ++l) s += Bj[l] * gphi[b][l];
                }
                acc[(j * MAXADJ + slots[b]) * CG_BLOCK] += w * s;
            }
        }
    });
}

template <int I>
__global__ void __launch_bounds__(CG_BLOCK) cg_jacobian_kernel(const ech_dg_args A) {
    extern __shared__ double sacc[];
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double *acc = sacc + threadIdx.x;
    for (int k = 0; k < NF * MAXADJ; ++k) acc[k * CG_BLOCK] = 0.0;
    if (r >= A.nnodes) return;
    Pt p;
    p.K = A.consts;
    for (int64_t e = A.inc_ptr[r]; e < A.inc_ptr[r + 1]; ++e) {
        const int64_t c = A.inc_cell[e];
        const int a = A.inc_loc[e];
        CellData cd;
        int nodes[NLOC], slots[NLOC];
        gather_cell(A, c, cd, nodes);
#pragma unroll
        for (int b = 0; b < NLOC; ++b) slots[b] = A.colslot[e * NLOC + b];
        p.cv = A.cell_volume[c];
        p.cd = A.cell_diam[c];
        double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
        Geo g;
        RowJ rj;
        for (int q = 0; q < NQC; ++q) {
            int idx[D];
            double w;
            cell_point(q, idx, w);
            geometry(cd.X, idx, g);
            basis(idx, g.invJ, phi, gphi);
            interpolate<NF>(cd.U, phi, gphi, p.u, p.gu);
            if constexpr (NAUX > 0) interpolate<ECH_NA1>(cd.AX, phi, gphi, p.a, p.ga);
#pragma unroll
            for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
            cell_jac<I>(p, rj);
            select_test(a, phi, gphi, phia, gphia);
            accum_row_slots<MC, I>(rj, w * fabs(g.detJ), phia, gphia, phi, gphi, slots, acc);
        }
        for (int lf = 0; lf < NLF; ++lf) {
            const int nb = A.nbr[c * NLF + lf];
            if (nb >= 0) continue;
            const int cls = -nb - 1;
            p.fa = A.facet_area[c * NLF + lf];
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], tt[D > 1 ? D - 1 : 1];
                double w, ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(cd.X, idx, g);
                basis(idx, g.invJ, phi, gphi);
                interpolate<NF>(cd.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cd.AX, phi, gphi, p.a, p.ga);
                facet_normal(g, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                efacet_jac<I>(cls, p, rj);
                select_test(a, phi, gphi, phia, gphia);
                accum_row_slots<ME, I>(rj, w * ds, phia, gphia, phi, gphi, slots, acc);
            }
        }
    }
    // row (I, r): for every coupled field j, the adj_count[r] column nodes in ascending order
    const int nadj = A.adj_count[r];
    double *__restrict__ vals = A.out + A.rowptr[(int64_t)I * A.nnodes + r];
    int off = 0;
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool own = OWN_BLOCK[I][j];
        if constexpr (own) {
            for (int s = 0; s < nadj; ++s) vals[off + s] = acc[(j * MAXADJ + s) * CG_BLOCK];
            off += nadj;
        }
    });
}

static int launch_cg_residual(const ech_dg_args *a, cudaStream_t s) {
    if (a->nnodes == 0) return 0;
    const int64_t grid = (a->nnodes + CG_BLOCK - 1) / CG_BLOCK;
    cg_residual_kernel<<<(unsigned)grid, CG_BLOCK, 0, s>>>(*a);
    return (int)cudaGetLastError();
}

template <int I>
static int launch_cg_jacobian_rows(const ech_dg_args *a, cudaStream_t s) {
    if constexpr (I < NF) {
        constexpr int smem = NF * MAXADJ * CG_BLOCK * 8;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(cg_jacobian_kernel<I>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            configured = true;
        }
        const int64_t grid = (a->nnodes + CG_BLOCK - 1) / CG_BLOCK;
        cg_jacobian_kernel<I><<<(unsigned)grid, CG_BLOCK, smem, s>>>(*a);
        int e = (int)cudaGetLastError();
        if (e) return e;
        return launch_cg_jacobian_rows<I + 1>(a, s);
    } else {
        return 0;
    }
}

static int launch_cg_jacobian(const ech_dg_args *a, cudaStream_t s) {
    if (a->nnodes == 0) return 0;
    return launch_cg_jacobian_rows<0>(a, s);
}

}  // namespace echcg
