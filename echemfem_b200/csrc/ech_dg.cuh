// DG residual / Jacobian assembly kernels for sm_100a (hand-written skeleton).
//
// Replaces, for discontinuous tensor-product spaces, what runs below the
// reference's `NonlinearVariationalSolver.solve()` for SNES function and
// Jacobian evaluation (echemfem/solver.py:714-725; SURVEY.md 3.3 hot loops
// 1-3): the TSFC element kernels, the PyOP2 cell / exterior-facet /
// interior-facet loops and MatSetValuesLocal.
//
// Design (B200-first, not a translation of those loops):
//  * OWNER-COMPUTES BY ROW.  One thread owns one CSR row (equation I, cell c,
//    test node a) -- or, for the residual, the NF entries of (c, a).  It adds
//    the cell integral and the test-side half of every facet of c, so each
//    CSR value is written exactly once, with plain vector stores: no atomics,
//    no colouring, no read-modify-write of the 8*nnz bytes that bound the
//    kernel.  Interior facets are visited from both cells, but each side only
//    computes its own rows (no redundant entry flops).
//  * Rows of one equation are contiguous over (c, a), so consecutive threads
//    write consecutive rows of the value array.
//  * Reference-element data is ONE 1D table (basis, derivative, abscissae at
//    the Gauss points and at 0 / 1) in constant memory; cell, own-facet and
//    neighbour-facet points are index tuples into it (warp-uniform indices).
//  * Pointwise physics (f0, f1, g0..g3) comes from the generated header;
//    compile-time masks prune every structurally zero coupling.
//
// Included after the generated header (namespace echgen).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include <utility>
#include "echem_b200_module.h"

namespace echdg {
using namespace echgen;

// iterative on purpose: a recursive helper that is not inlined makes the stack size unknowable and
// caps the kernel at cudaLimitStackSize
__host__ __device__ constexpr int ipow(int b, int e) {
    int r = 1;
    for (int i = 0; i < e; ++i) r *= b;
    return r;
}
constexpr int NV = 1 << D;
constexpr int NLF = 2 * D;
constexpr int NLOC = ipow(N1, D);
constexpr int NQC = ipow(NQ, D);
constexpr int NQF = ipow(NQ, D - 1);
constexpr int NAUX = ECH_NAUX;

template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

__host__ __device__ constexpr int digit(int b, int k) { return (b / ipow(N1, D - 1 - k)) % N1; }

struct Geo {
    double x[D];
    double invJ[D][D];
    double detJ;
};

// multilinear map of the cell with vertices X at the table point idx
__device__ __forceinline__ void geometry(const double (&X)[NV][D], const int (&idx)[D], Geo &g) {
    double xi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = TAB_X[idx[k]];
    double J[D][D];
    // Edge form: x = X_0 + sum_v N_v (X_v - X_0), J = sum_v dN_v (X_v - X_0) (the shape functions sum to one, their
    // derivatives to zero).  The differences of neighbouring vertices are exact in floating point, so a cell of
    // size h at distance |X| from the origin keeps full relative accuracy in J instead of losing log10(|X| / h)
    // digits to cancellation (graded Bortels meshes: |X| / h ~ 1e4).
#pragma unroll
    for (int a = 0; a < D; ++a) {
        g.x[a] = X[0][a];
#pragma unroll
        for (int m = 0; m < D; ++m) J[a][m] = 0.0;
    }
#pragma unroll
    for (int v = 1; v < NV; ++v) {
        double N = 1.0, dN[D];
#pragma unroll
        for (int m = 0; m < D; ++m) dN[m] = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const bool bit = (v >> (D - 1 - k)) & 1;
            const double val = bit ? xi[k] : 1.0 - xi[k];
            const double sg = bit ? 1.0 : -1.0;
            N *= val;
#pragma unroll
            for (int m = 0; m < D; ++m) dN[m] *= (m == k) ? sg : val;
        }
#pragma unroll
        for (int a = 0; a < D; ++a) {
            const double e = X[v][a] - X[0][a];
            g.x[a] += N * e;
#pragma unroll
            for (int m = 0; m < D; ++m) J[a][m] += dN[m] * e;
        }
    }
    if constexpr (D == 1) {
        g.detJ = J[0][0];
        g.invJ[0][0] = 1.0 / J[0][0];
    } else if constexpr (D == 2) {
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double r = 1.0 / det;
        g.detJ = det;
        g.invJ[0][0] = J[1][1] * r;
        g.invJ[0][1] = -J[0][1] * r;
        g.invJ[1][0] = -J[1][0] * r;
        g.invJ[1][1] = J[0][0] * r;
    } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        const double r = 1.0 / det;
        g.detJ = det;
        g.invJ[0][0] = c00 * r;
        g.invJ[1][0] = c01 * r;
        g.invJ[2][0] = c02 * r;
        g.invJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
        g.invJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
        g.invJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
        g.invJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
        g.invJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
        g.invJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    }
}

// tensor basis and physical gradients at the table point idx
__device__ __forceinline__ void basis(const int (&idx)[D], const double (&invJ)[D][D], double (&phi)[NLOC],
                                      double (&gphi)[NLOC][D]) {
    double L[D][N1], dL[D][N1];
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
        for (int m = 0; m < N1; ++m) {
            L[k][m] = TAB_L[idx[k]][m];
            dL[k][m] = TAB_DL[idx[k]][m];
        }
#pragma unroll
    for (int b = 0; b < NLOC; ++b) {
        double v = 1.0, dref[D];
#pragma unroll
        for (int m = 0; m < D; ++m) dref[m] = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int bk = digit(b, k);
            v *= L[k][bk];
#pragma unroll
            for (int m = 0; m < D; ++m) dref[m] *= (m == k) ? dL[k][bk] : L[k][bk];
        }
        phi[b] = v;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) s += invJ[m][a] * dref[m];
            gphi[b][a] = s;
        }
    }
}

template <int NFLD>
__device__ __forceinline__ void interpolate(const double (&U)[NFLD][NLOC], const double (&phi)[NLOC],
                                            const double (&gphi)[NLOC][D], double (&u)[NFLD], double (&gu)[NFLD][D]) {
#pragma unroll
    for (int j = 0; j < NFLD; ++j) {
        double s = 0.0, g[D];
#pragma unroll
        for (int a = 0; a < D; ++a) g[a] = 0.0;
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            s += U[j][b] * phi[b];
#pragma unroll
            for (int a = 0; a < D; ++a) g[a] += U[j][b] * gphi[b][a];
        }
        u[j] = s;
#pragma unroll
        for (int a = 0; a < D; ++a) gu[j][a] = g[a];
    }
}

__device__ __forceinline__ void cell_point(int q, int (&idx)[D], double &w) {
    w = 1.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        idx[k] = (q / ipow(NQ, D - 1 - k)) % NQ;
        w *= TAB_W[idx[k]];
    }
}

// own-side facet point: direction kf pinned to side s, remaining directions take the digits of qf
__device__ __forceinline__ void facet_point(int lf, int qf, int (&idx)[D], int (&t)[D > 1 ? D - 1 : 1], double &w) {
    const int kf = lf >> 1, s = lf & 1;
    w = 1.0;
    if constexpr (D > 1) {
#pragma unroll
        for (int m = 0; m < D - 1; ++m) {
            t[m] = (qf / ipow(NQ, D - 2 - m)) % NQ;
            w *= TAB_W[t[m]];
        }
    } else {
        t[0] = 0;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        // remaining direction k carries facet digit (k < kf ? k : k - 1)
        const int lo = t[0];
        const int hi = t[D > 2 ? 1 : 0];
        const int m = (k < kf) ? k : k - 1;
        idx[k] = (k == kf) ? NQ + s : ((m == 0) ? lo : hi);
    }
}

// the same physical point seen from the neighbour (local facet lf2, orientation code o)
__device__ __forceinline__ void neighbour_point(int code, const int (&t)[D > 1 ? D - 1 : 1], int (&idx)[D]) {
    const int lf2 = code & 7, o = code >> 3;
    const int kf = lf2 >> 1, s = lf2 & 1;
    int a0 = (o & 2) ? NQ - 1 - t[0] : t[0];
    int a1 = a0;
    if constexpr (D == 3) {
        a1 = (o & 4) ? NQ - 1 - t[1] : t[1];
        if (o & 1) {
            const int tmp = a0;
            a0 = a1;
            a1 = tmp;
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const int m = (k < kf) ? k : k - 1;
        idx[k] = (k == kf) ? NQ + s : ((m == 0) ? a0 : a1);
    }
}

__device__ __forceinline__ void facet_normal(const Geo &g, int lf, double (&n)[D], double &ds) {
    const int kf = lf >> 1;
    const double sg = (lf & 1) ? 1.0 : -1.0;
    double nn = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) v = (k == kf) ? g.invJ[k][a] : v;
        n[a] = v * sg;
        nn += n[a] * n[a];
    }
    const double nrm = sqrt(nn);
    const double r = 1.0 / nrm;
#pragma unroll
    for (int a = 0; a < D; ++a) n[a] *= r;
    ds = fabs(g.detJ) * nrm;
}

template <int NFLD>
__device__ __forceinline__ void load_cell_fields(const double *__restrict__ base, int64_t stride, int64_t c,
                                                 double (&U)[NFLD][NLOC]) {
#pragma unroll
    for (int j = 0; j < NFLD; ++j) {
        const double *p = base + j * stride + c * NLOC;
        if constexpr (NLOC % 2 == 0) {
#pragma unroll
            for (int b = 0; b < NLOC; b += 2) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(p + b));
                U[j][b] = v.x;
                U[j][b + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int b = 0; b < NLOC; ++b) U[j][b] = __ldg(p + b);
        }
    }
}

__device__ __forceinline__ void load_cell_coords(const double *__restrict__ xcell, int64_t c, double (&X)[NV][D]) {
    const double2 *p = reinterpret_cast<const double2 *>(xcell + c * (NV * D));
#pragma unroll
    for (int i = 0; i < NV * D / 2; ++i) {
        const double2 v = __ldg(p + i);
        (&X[0][0])[2 * i] = v.x;
        (&X[0][0])[2 * i + 1] = v.y;
    }
}

template <class T>
__device__ __forceinline__ void select_test(int a, const double (&phi)[NLOC], const double (&gphi)[NLOC][D], T &phia,
                                            double (&gphia)[D]) {
    phia = phi[0];
#pragma unroll
    for (int k = 0; k < D; ++k) gphia[k] = gphi[0][k];
#pragma unroll
    for (int b = 1; b < NLOC; ++b) {
        const bool hit = (b == a);
        phia = hit ? phi[b] : phia;
#pragma unroll
        for (int k = 0; k < D; ++k) gphia[k] = hit ? gphi[b][k] : gphia[k];
    }
}

// ---------------------------------------------------------------- residual
template <class M>
__device__ __forceinline__ void accum_res(const Res &o, double w, double phia, const double (&gphia)[D],
                                          double (&r)[NF]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if constexpr (M::F0[i]) r[i] += w * o.f0[i] * phia;
        if constexpr (M::F1[i]) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) s += o.f1[i][k] * gphia[k];
            r[i] += w * s;
        }
    });
}

__global__ void __launch_bounds__(128) residual_kernel(const ech_dg_args A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.ncell * NLOC) return;
    const int64_t c = t / NLOC;
    const int a = (int)(t - c * NLOC);
    const int64_t fstride = A.ncell_total * NLOC;

    double X[NV][D], U[NF][NLOC], AX[ECH_NA1][NLOC];
    load_cell_coords(A.xcell, c, X);
    load_cell_fields<NF>(A.u, fstride, c, U);
    if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, c, AX);
    const double cv = A.cell_volume[c];
    const double cd = A.cell_diam[c];

    double r[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) r[i] = 0.0;

    Pt p;
    p.K = A.consts;
    p.cv = cv;
    p.cd = cd;
    double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
    Geo g;
    Res o;

    for (int q = 0; q < NQC; ++q) {
        int idx[D];
        double w;
        cell_point(q, idx, w);
        geometry(X, idx, g);
        basis(idx, g.invJ, phi, gphi);
        interpolate<NF>(U, phi, gphi, p.u, p.gu);
        if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
#pragma unroll
        for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
        cell_res(p, o);
        select_test(a, phi, gphi, phia, gphia);
        accum_res<MC>(o, w * fabs(g.detJ), phia, gphia, r);
    }

    for (int lf = 0; lf < NLF; ++lf) {
        const int nb = A.nbr[c * NLF + lf];
        p.fa = A.facet_area[c * NLF + lf];
        if (nb < 0) {
            const int cls = -nb - 1;
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], tt[D > 1 ? D - 1 : 1];
                double w, ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(X, idx, g);
                basis(idx, g.invJ, phi, gphi);
                interpolate<NF>(U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
                facet_normal(g, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                efacet_res(cls, p, o);
                select_test(a, phi, gphi, phia, gphia);
                accum_res<ME>(o, w * ds, phia, gphia, r);
            }
        } else {
            if constexpr (ECH_FAMILY_DG) {
                double X2[NV][D], U2[NF][NLOC], AX2[ECH_NA1][NLOC];
                load_cell_coords(A.xcell, nb, X2);
                load_cell_fields<NF>(A.u, fstride, nb, U2);
                if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, nb, AX2);
                p.cvm = A.cell_volume[nb];
                p.cdm = A.cell_diam[nb];
                const int code = A.code[c * NLF + lf];
                for (int qf = 0; qf < NQF; ++qf) {
                    int idx[D], idx2[D], tt[D > 1 ? D - 1 : 1];
                    double w, ds;
                    facet_point(lf, qf, idx, tt, w);
                    neighbour_point(code, tt, idx2);
                    Geo g2;
                    double phi2[NLOC], gphi2[NLOC][D];
                    geometry(X2, idx2, g2);
                    basis(idx2, g2.invJ, phi2, gphi2);
                    interpolate<NF>(U2, phi2, gphi2, p.um, p.gum);
                    if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX2, phi2, gphi2, p.am, p.gam);
                    geometry(X, idx, g);
                    basis(idx, g.invJ, phi, gphi);
                    interpolate<NF>(U, phi, gphi, p.u, p.gu);
                    if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
                    facet_normal(g, lf, p.n, ds);
#pragma unroll
                    for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                    ifacet_res(p, o);
                    select_test(a, phi, gphi, phia, gphia);
                    accum_res<MIP>(o, w * ds, phia, gphia, r);
                }
            }
        }
    }
    const int64_t ostride = A.ncell_total * NLOC;     // ghost rows exist (empty) so every vector shares one layout
#pragma unroll
    for (int i = 0; i < NF; ++i) A.out[i * ostride + t] = r[i];
}

// ---------------------------------------------------------------- Jacobian
template <class M, int I>
__device__ __forceinline__ void accum_row(const RowJ &r, double w, double phia, const double (&gphia)[D],
                                          const double (&phi)[NLOC], const double (&gphi)[NLOC][D],
                                          double (&acc)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool h0 = M::G0[I][j], h1 = M::G1[I][j], h2 = M::G2[I][j];
        constexpr int h3 = M::G3[I][j];
        if constexpr (h0 || h1 || h2 || h3 != 0) {
            double Aj = 0.0, Bj[D];
#pragma unroll
            for (int l = 0; l < D; ++l) Bj[l] = 0.0;
            if constexpr (h0) Aj += r.g0[j] * phia;
            if constexpr (h2) {
#pragma unroll
                for (int k = 0; k < D; ++k) Aj += r.g2[j][k] * gphia[k];
            }
            if constexpr (h1) {
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += r.g1[j][l] * phia;
            }
            if constexpr (h3 == 1) {
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += r.g3[j][0][0] * gphia[l];
            }
            if constexpr (h3 == 2) {
#pragma unroll
                for (int l = 0; l < D; ++l)
#pragma unroll
                    for (int k = 0; k < D; ++k) Bj[l] += r.g3[j][k][l] * gphia[k];
            }
            Aj *= w;
#pragma unroll
            for (int l = 0; l < D; ++l) Bj[l] *= w;
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                double s = acc[j][b];
                if constexpr (h0 || h2) s += Aj * phi[b];
                if constexpr (h1 || h3 != 0) {
#pragma unroll
                    for (int l = 0; l < D; ++l) s += Bj[l] * gphi[b][l];
                }
                acc[j][b] = s;
            }
        }
    });
}

__device__ __forceinline__ void store_block(double *__restrict__ dst, const double (&v)[NLOC]) {
    if constexpr (NLOC % 4 == 0) {
        // one 256-bit store per 4 entries (STG.E.ENL2.256): row blocks start on 32-byte boundaries
#pragma unroll
        for (int b = 0; b < NLOC; b += 4)
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + b), "d"(v[b]), "d"(v[b + 1]),
                         "d"(v[b + 2]), "d"(v[b + 3])
                         : "memory");
    } else if constexpr (NLOC % 2 == 0) {
#pragma unroll
        for (int b = 0; b < NLOC; b += 2) reinterpret_cast<double2 *>(dst)[b / 2] = make_double2(v[b], v[b + 1]);
    } else {
#pragma unroll
        for (int b = 0; b < NLOC; ++b) dst[b] = v[b];
    }
}

// column offset (in doubles, relative to the row start) of the cell block `pos` of field j in a row of equation I
template <int I, int J>
__device__ __forceinline__ int block_offset(int ncol, int pos) {
    int off = 0;
    static_for<0, J>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool own = OWN_BLOCK[I][j], nb = NBR_BLOCK[I][j];
        if constexpr (own) off += (nb ? ncol : 1) * NLOC;
    });
    constexpr bool nbJ = NBR_BLOCK[I][J];
    return off + (nbJ ? pos : 0) * NLOC;
}

template <int I>
__global__ void __launch_bounds__(128) jacobian_kernel(const ech_dg_args A) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.ncell * NLOC) return;
    const int64_t c = t / NLOC;
    const int a = (int)(t - c * NLOC);
    const int64_t fstride = A.ncell_total * NLOC;
    const int64_t row = (int64_t)I * A.ncell_total * NLOC + t;
    double *__restrict__ vals = A.out + A.rowptr[row];
    const int ncol = A.ncol[c];

    double X[NV][D], U[NF][NLOC], AX[ECH_NA1][NLOC];
    load_cell_coords(A.xcell, c, X);
    load_cell_fields<NF>(A.u, fstride, c, U);
    if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, c, AX);

    double acc[NF][NLOC];
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (OWN_BLOCK[I][j]) {
#pragma unroll
            for (int b = 0; b < NLOC; ++b) acc[j][b] = 0.0;
        }
    });

    Pt p;
    p.K = A.consts;
    p.cv = A.cell_volume[c];
    p.cd = A.cell_diam[c];
    double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
    Geo g;
    RowJ rj;

    for (int q = 0; q < NQC; ++q) {
        int idx[D];
        double w;
        cell_point(q, idx, w);
        geometry(X, idx, g);
        basis(idx, g.invJ, phi, gphi);
        interpolate<NF>(U, phi, gphi, p.u, p.gu);
        if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
#pragma unroll
        for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
        cell_jac<I>(p, rj);
        select_test(a, phi, gphi, phia, gphia);
        accum_row<MC, I>(rj, w * fabs(g.detJ), phia, gphia, phi, gphi, acc);
    }

    for (int lf = 0; lf < NLF; ++lf) {
        const int nb = A.nbr[c * NLF + lf];
        p.fa = A.facet_area[c * NLF + lf];
        if (nb < 0) {
            const int cls = -nb - 1;
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], tt[D > 1 ? D - 1 : 1];
                double w, ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(X, idx, g);
                basis(idx, g.invJ, phi, gphi);
                interpolate<NF>(U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
                facet_normal(g, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                efacet_jac<I>(cls, p, rj);
                select_test(a, phi, gphi, phia, gphia);
                accum_row<ME, I>(rj, w * ds, phia, gphia, phi, gphi, acc);
            }
        } else {
            if constexpr (ECH_FAMILY_DG) {
                double X2[NV][D], U2[NF][NLOC], AX2[ECH_NA1][NLOC];
                load_cell_coords(A.xcell, nb, X2);
                load_cell_fields<NF>(A.u, fstride, nb, U2);
                if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, nb, AX2);
                p.cvm = A.cell_volume[nb];
                p.cdm = A.cell_diam[nb];
                const int code = A.code[c * NLF + lf];
                double accn[NF][NLOC];
                static_for<0, NF>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if constexpr (NBR_BLOCK[I][j]) {
#pragma unroll
                        for (int b = 0; b < NLOC; ++b) accn[j][b] = 0.0;
                    }
                });
                RowJ rjm;
                for (int qf = 0; qf < NQF; ++qf) {
                    int idx[D], idx2[D], tt[D > 1 ? D - 1 : 1];
                    double w, ds;
                    facet_point(lf, qf, idx, tt, w);
                    neighbour_point(code, tt, idx2);
                    Geo g2;
                    double phi2[NLOC], gphi2[NLOC][D];
                    geometry(X2, idx2, g2);
                    basis(idx2, g2.invJ, phi2, gphi2);
                    interpolate<NF>(U2, phi2, gphi2, p.um, p.gum);
                    if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX2, phi2, gphi2, p.am, p.gam);
                    geometry(X, idx, g);
                    basis(idx, g.invJ, phi, gphi);
                    interpolate<NF>(U, phi, gphi, p.u, p.gu);
                    if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
                    facet_normal(g, lf, p.n, ds);
#pragma unroll
                    for (int k = 0; k < D; ++k) p.x[k] = g.x[k];
                    ifacet_jac<I>(p, rj, rjm);
                    select_test(a, phi, gphi, phia, gphia);
                    accum_row<MIP, I>(rj, w * ds, phia, gphia, phi, gphi, acc);
                    accum_row<MIM, I>(rjm, w * ds, phia, gphia, phi2, gphi2, accn);
                }
                const int pos = A.slot[c * NLF + lf];
                static_for<0, NF>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    if constexpr (NBR_BLOCK[I][j]) store_block(vals + block_offset<I, j>(ncol, pos), accn[j]);
                });
            }
        }
    }
    const int pos = A.slot_self[c];
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (OWN_BLOCK[I][j]) store_block(vals + block_offset<I, j>(ncol, pos), acc[j]);
    });
}

// ---------------------------------------------------------------- launches
static int launch_residual(const ech_dg_args *a, cudaStream_t s) {
    const int64_t nthreads = a->ncell * NLOC;
    if (nthreads == 0) return 0;
    const int block = 128;
    const int64_t grid = (nthreads + block - 1) / block;
    residual_kernel<<<(unsigned)grid, block, 0, s>>>(*a);
    return (int)cudaGetLastError();
}

template <int I>
static int launch_jacobian_rows(const ech_dg_args *a, cudaStream_t s) {
    if constexpr (I < NF) {
        const int64_t nthreads = a->ncell * NLOC;
        const int block = 128;
        const int64_t grid = (nthreads + block - 1) / block;
        jacobian_kernel<I><<<(unsigned)grid, block, 0, s>>>(*a);
        int e = (int)cudaGetLastError();
        if (e) return e;
        return launch_jacobian_rows<I + 1>(a, s);
    } else {
        return 0;
    }
}

static int launch_jacobian(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    return launch_jacobian_rows<0>(a, s);
}

}  // namespace echdg
