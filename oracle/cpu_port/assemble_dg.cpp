// CPU port of the DG residual + Jacobian assembly (TEST / BASELINE INFRASTRUCTURE -- oracle/, never shipped).
//
// What it restates: the work below the reference's NonlinearVariationalSolver.solve() for one SNES function +
// Jacobian evaluation (echemfem/solver.py:714-725; SURVEY.md 3.3 hot loops 1-3): TSFC-style element kernels
// (tabulate, interpolate, pointwise integrand, contraction with the test/trial bases), PyOP2-style loops over
// cells and their facets, insertion into a field-blocked AIJ matrix.  The pointwise integrands f0, f1, g0..g3
// are the generated header of the same weak form the GPU module is built from (compiled here by g++ through
// host_shim.h); everything else -- geometry, bases, loops, insertion -- is written independently of csrc/.
//
// Roles: (1) `bench.py`'s cpu_baseline / `--impl reference` arm ("kind": "port"): analytic Jacobian, compiled
// -O3 -march=native, OpenMP over cells -- an honest stand-in for Firedrake's generated C on the same cores;
// (2) a second oracle for parity at sizes the numpy oracle cannot reach (tests compare it with the numpy oracle
// on small meshes first).  Work decomposition: a cell owns its rows; it adds its cell integral and, for each of
// its facets, the half of the facet integral tested with its own basis (interior facets are visited from both
// sides, as under MPI with a one-facet overlap), so every CSR value is written once and threads never collide.
//
// Build (oracle/cpu_port/build.py):
//   g++ -O3 -march=native -fopenmp -std=c++17 -shared -fPIC -include host_shim.h -include <generated header>
//       assemble_dg.cpp -o oracle/_build/libcpu_port_<key>.so
#include <omp.h>
#include <string.h>
#include <type_traits>
#include <utility>

namespace port {
using namespace echgen;

constexpr int ipow(int b, int e) {
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
constexpr int NA1 = ECH_NA1;

template <int B, int E, class F>
inline void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

struct Mesh {
    int64_t ncell, ncell_total;
    const double *xcell;        // [ncell_total][2^d][d]
    const int32_t *nbr;         // [ncell][2d]
    const uint8_t *code, *slot; // [ncell][2d]
    const uint8_t *slot_self, *ncol;
    const double *cell_volume, *cell_diam, *facet_area;
};

// ---- one evaluation point of one cell: multilinear geometry + tensor Lagrange basis (table indices per direction;
// index NQ + s is the face xi_k = s; reference cell and local numbering of SURVEY.md 8c Q3)
struct Point {
    double x[D], invJ[D][D], detJ;
    double phi[NLOC], gphi[NLOC][D];
};

static inline void eval_point(const double *X /*[NV][D]*/, const int *idx, Point &P) {
    double xi[D], J[D][D];
    for (int k = 0; k < D; ++k) xi[k] = TAB_X[idx[k]];
    // edge form: differences of neighbouring vertices are exact (no cancellation for small cells far from the origin)
    for (int a = 0; a < D; ++a) {
        P.x[a] = X[a];
        for (int m = 0; m < D; ++m) J[a][m] = 0.0;
    }
    for (int v = 1; v < NV; ++v) {
        double N = 1.0, dN[D];
        for (int m = 0; m < D; ++m) dN[m] = 1.0;
        for (int k = 0; k < D; ++k) {
            const bool bit = (v >> (D - 1 - k)) & 1;
            const double val = bit ? xi[k] : 1.0 - xi[k];
            const double sg = bit ? 1.0 : -1.0;
            N *= val;
            for (int m = 0; m < D; ++m) dN[m] *= (m == k) ? sg : val;
        }
        for (int a = 0; a < D; ++a) {
            const double e = X[v * D + a] - X[a];
            P.x[a] += N * e;
            for (int m = 0; m < D; ++m) J[a][m] += dN[m] * e;
        }
    }
    if constexpr (D == 1) {
        P.detJ = J[0][0];
        P.invJ[0][0] = 1.0 / J[0][0];
    } else if constexpr (D == 2) {
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double r = 1.0 / det;
        P.detJ = det;
        P.invJ[0][0] = J[1][1] * r;
        P.invJ[0][1] = -J[0][1] * r;
        P.invJ[1][0] = -J[1][0] * r;
        P.invJ[1][1] = J[0][0] * r;
    } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        const double r = 1.0 / det;
        P.detJ = det;
        P.invJ[0][0] = c00 * r;
        P.invJ[1][0] = c01 * r;
        P.invJ[2][0] = c02 * r;
        P.invJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
        P.invJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
        P.invJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
        P.invJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
        P.invJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
        P.invJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    }
    for (int b = 0; b < NLOC; ++b) {
        double v = 1.0, dref[D];
        for (int m = 0; m < D; ++m) dref[m] = 1.0;
        int rem = b;
        for (int k = D - 1; k >= 0; --k) {
            const int bk = rem % N1;
            rem /= N1;
            v *= TAB_L[idx[k]][bk];
            for (int m = 0; m < D; ++m) dref[m] *= (m == k) ? TAB_DL[idx[k]][bk] : TAB_L[idx[k]][bk];
        }
        P.phi[b] = v;
        for (int a = 0; a < D; ++a) {
            double s = 0.0;
            for (int m = 0; m < D; ++m) s += P.invJ[m][a] * dref[m];
            P.gphi[b][a] = s;
        }
    }
}

template <int NFLD>
static inline void interpolate(const double *U /*[NFLD][NLOC]*/, const Point &P, double *u, double (*gu)[D]) {
    for (int j = 0; j < NFLD; ++j) {
        double s = 0.0, g[D];
        for (int a = 0; a < D; ++a) g[a] = 0.0;
        for (int b = 0; b < NLOC; ++b) {
            s += U[j * NLOC + b] * P.phi[b];
            for (int a = 0; a < D; ++a) g[a] += U[j * NLOC + b] * P.gphi[b][a];
        }
        u[j] = s;
        for (int a = 0; a < D; ++a) gu[j][a] = g[a];
    }
}

static inline double cell_point(int q, int *idx) {
    double w = 1.0;
    for (int k = 0; k < D; ++k) {
        idx[k] = (q / ipow(NQ, D - 1 - k)) % NQ;
        w *= TAB_W[idx[k]];
    }
    return w;
}

// own-side facet point: direction kf pinned to side s, the other directions carry the digits t of qf
static inline double facet_point(int lf, int qf, int *idx, int *t) {
    const int kf = lf >> 1, s = lf & 1;
    double w = 1.0;
    t[0] = 0;
    t[1] = 0;
    for (int m = 0; m < D - 1; ++m) {
        t[m] = (qf / ipow(NQ, D - 2 - m)) % NQ;
        w *= TAB_W[t[m]];
    }
    for (int k = 0; k < D; ++k) {
        const int m = (k < kf) ? k : k - 1;
        idx[k] = (k == kf) ? NQ + s : t[m == 0 ? 0 : 1];
    }
    return w;
}

// the same physical point in the neighbour: code = its local facet | orientation << 3
// (orientation bit0: swap of the two facet axes (3D), bit1 / bit2: flip of my facet axis 0 / 1)
static inline void neighbour_point(int code, const int *t, int *idx) {
    const int lf2 = code & 7, o = code >> 3;
    const int kf = lf2 >> 1, s = lf2 & 1;
    int a0 = (o & 2) ? NQ - 1 - t[0] : t[0];
    int a1 = a0;
    if (D == 3) {
        a1 = (o & 4) ? NQ - 1 - t[1] : t[1];
        if (o & 1) {
            const int tmp = a0;
            a0 = a1;
            a1 = tmp;
        }
    }
    for (int k = 0; k < D; ++k) {
        const int m = (k < kf) ? k : k - 1;
        idx[k] = (k == kf) ? NQ + s : (m == 0 ? a0 : a1);
    }
}

static inline double facet_normal(const Point &P, int lf, double *n) {
    const int kf = lf >> 1;
    const double sg = (lf & 1) ? 1.0 : -1.0;
    double nn = 0.0;
    for (int a = 0; a < D; ++a) {
        n[a] = P.invJ[kf][a] * sg;
        nn += n[a] * n[a];
    }
    const double nrm = sqrt(nn);
    for (int a = 0; a < D; ++a) n[a] /= nrm;
    return fabs(P.detJ) * nrm;
}

// ---- contraction of pointwise coefficients with the bases (what a TSFC kernel does per quadrature point)
template <class M>
static inline void add_res(const Res &o, double w, const Point &P, double (*R)[NLOC]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        for (int a = 0; a < NLOC; ++a) {
            double s = 0.0;
            if constexpr (M::F0[i]) s += o.f0[i] * P.phi[a];
            if constexpr (M::F1[i])
                for (int k = 0; k < D; ++k) s += o.f1[i][k] * P.gphi[a][k];
            R[i][a] += w * s;
        }
    });
}

// A[I][a][j][b] += w * (g0 phi_a phi_b + phi_a g1 . grad phi_b + grad phi_a . g2 phi_b + grad phi_a . g3 . grad phi_b);
// test functions from Pt (own cell), trial functions from Pu (own cell or neighbour)
template <class M, int I>
static inline void add_row(const RowJ &r, double w, const Point &Pt_, const Point &Pu, double (*A)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool h0 = M::G0[I][j], h1 = M::G1[I][j], h2 = M::G2[I][j];
        constexpr int h3 = M::G3[I][j];
        if constexpr (h0 || h1 || h2 || h3 != 0) {
            for (int a = 0; a < NLOC; ++a) {
                double Aj = 0.0, Bj[D];
                for (int l = 0; l < D; ++l) Bj[l] = 0.0;
                if constexpr (h0) Aj += r.g0[j] * Pt_.phi[a];
                if constexpr (h2)
                    for (int k = 0; k < D; ++k) Aj += r.g2[j][k] * Pt_.gphi[a][k];
                if constexpr (h1)
                    for (int l = 0; l < D; ++l) Bj[l] += r.g1[j][l] * Pt_.phi[a];
                if constexpr (h3 == 1)
                    for (int l = 0; l < D; ++l) Bj[l] += r.g3[j][0][0] * Pt_.gphi[a][l];
                if constexpr (h3 == 2)
                    for (int l = 0; l < D; ++l)
                        for (int k = 0; k < D; ++k) Bj[l] += r.g3[j][k][l] * Pt_.gphi[a][k];
                Aj *= w;
                for (int l = 0; l < D; ++l) Bj[l] *= w;
                double *row = A[a][j];
                for (int b = 0; b < NLOC; ++b) {
                    double s = row[b];
                    if constexpr (h0 || h2) s += Aj * Pu.phi[b];
                    if constexpr (h1 || h3 != 0)
                        for (int l = 0; l < D; ++l) s += Bj[l] * Pu.gphi[b][l];
                    row[b] = s;
                }
            }
        }
    });
}

// offset (doubles, from the row start) of the cell block `pos` of field j in a row of equation I
template <int I>
static inline int block_offset(int j, int ncol, int pos) {
    int off = 0;
    for (int jj = 0; jj < j; ++jj)
        if (OWN_BLOCK[I][jj]) off += (NBR_BLOCK[I][jj] ? ncol : 1) * NLOC;
    return off + (NBR_BLOCK[I][j] ? pos : 0) * NLOC;
}

static inline void load_fields(const double *base, int64_t stride, int64_t c, int nfld, double *U) {
    for (int j = 0; j < nfld; ++j)
        for (int b = 0; b < NLOC; ++b) U[j * NLOC + b] = base[j * stride + c * NLOC + b];
}

// fold one element-level piece (cell integral, one facet) into the cell's totals; the sums of absolute pieces are the
// element-level scales of the entries (tau_block of tests/parity.py)
template <int N>
static inline void fold(double *piece, double *total, double *total_abs) {
    for (int k = 0; k < N; ++k) {
        total[k] += piece[k];
        total_abs[k] += fabs(piece[k]);
        piece[k] = 0.0;
    }
}

// rows of cell c: residual (F != NULL) and / or Jacobian values (vals != NULL); Fs / vs (optional): sums of the
// absolute element-level contributions to the same entries
static void assemble_cell(const Mesh &m, int64_t c, const double *u, const double *aux, const double *consts,
                          const int64_t *rowptr, double *vals, double *F, double *vs, double *Fs) {
    const int64_t fstride = m.ncell_total * NLOC;
    const double *X = m.xcell + c * (NV * D);
    double U[NF * NLOC], AX[NA1 * NLOC];
    load_fields(u, fstride, c, NF, U);
    if (NAUX > 0) load_fields(aux, fstride, c, NAUX, AX);
    constexpr int NR = NF * NLOC, NA = NF * NLOC * NF * NLOC;
    double R[NF][NLOC], Rt[NF][NLOC], Rabs[NF][NLOC];                // piece, total, sum of |pieces|
    double Aown[NF][NLOC][NF][NLOC], Anb[NF][NLOC][NF][NLOC];       // [I][a][j][b]; Aown: current piece
    double At[NF][NLOC][NF][NLOC], Aabs[NF][NLOC][NF][NLOC];
    memset(R, 0, sizeof(R));
    memset(Rt, 0, sizeof(Rt));
    memset(Rabs, 0, sizeof(Rabs));
    memset(Aown, 0, sizeof(Aown));
    memset(At, 0, sizeof(At));
    memset(Aabs, 0, sizeof(Aabs));
    Pt p;
    p.K = consts;
    p.cv = m.cell_volume[c];
    p.cd = m.cell_diam[c];
    Point P, P2;
    Res o;
    const bool doJ = vals != nullptr, doF = F != nullptr;
    const int ncol = m.ncol[c];

    for (int q = 0; q < NQC; ++q) {
        int idx[D];
        double w = cell_point(q, idx);
        eval_point(X, idx, P);
        interpolate<NF>(U, P, p.u, p.gu);
        if (NAUX > 0) interpolate<NA1>(AX, P, p.a, p.ga);
        for (int k = 0; k < D; ++k) p.x[k] = P.x[k];
        w *= fabs(P.detJ);
        if (doF) {
            cell_res(p, o);
            add_res<MC>(o, w, P, R);
        }
        if (doJ)
            static_for<0, NF>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                RowJ rj;
                cell_jac<I>(p, rj);
                add_row<MC, I>(rj, w, P, P, Aown[I]);
            });
    }
    fold<NR>(&R[0][0], &Rt[0][0], &Rabs[0][0]);
    fold<NA>(&Aown[0][0][0][0], &At[0][0][0][0], &Aabs[0][0][0][0]);
    for (int lf = 0; lf < NLF; ++lf) {
        const int nb = m.nbr[c * NLF + lf];
        p.fa = m.facet_area[c * NLF + lf];
        if (nb < 0) {
            const int cls = -nb - 1;
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], t[2];
                double w = facet_point(lf, qf, idx, t);
                eval_point(X, idx, P);
                interpolate<NF>(U, P, p.u, p.gu);
                if (NAUX > 0) interpolate<NA1>(AX, P, p.a, p.ga);
                w *= facet_normal(P, lf, p.n);
                for (int k = 0; k < D; ++k) p.x[k] = P.x[k];
                if (doF) {
                    efacet_res(cls, p, o);
                    add_res<ME>(o, w, P, R);
                }
                if (doJ)
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        RowJ rj;
                        efacet_jac<I>(cls, p, rj);
                        add_row<ME, I>(rj, w, P, P, Aown[I]);
                    });
            }
        } else {
#if ECH_FAMILY_DG
            const double *X2 = m.xcell + (int64_t)nb * (NV * D);
            double U2[NF * NLOC], AX2[NA1 * NLOC];
            load_fields(u, fstride, nb, NF, U2);
            if (NAUX > 0) load_fields(aux, fstride, nb, NAUX, AX2);
            p.cvm = m.cell_volume[nb];
            p.cdm = m.cell_diam[nb];
            const int code = m.code[c * NLF + lf];
            if (doJ) memset(Anb, 0, sizeof(Anb));
            for (int qf = 0; qf < NQF; ++qf) {
                int idx[D], idx2[D], t[2];
                double w = facet_point(lf, qf, idx, t);
                neighbour_point(code, t, idx2);
                eval_point(X2, idx2, P2);
                interpolate<NF>(U2, P2, p.um, p.gum);
                if (NAUX > 0) interpolate<NA1>(AX2, P2, p.am, p.gam);
                eval_point(X, idx, P);
                interpolate<NF>(U, P, p.u, p.gu);
                if (NAUX > 0) interpolate<NA1>(AX, P, p.a, p.ga);
                w *= facet_normal(P, lf, p.n);
                for (int k = 0; k < D; ++k) p.x[k] = P.x[k];
                if (doF) {
                    ifacet_res(p, o);
                    add_res<MIP>(o, w, P, R);
                }
                if (doJ)
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        RowJ rj, rjm;
                        ifacet_jac<I>(p, rj, rjm);
                        add_row<MIP, I>(rj, w, P, P, Aown[I]);
                        add_row<MIM, I>(rjm, w, P, P2, Anb[I]);
                    });
            }
            if (doJ) {
                const int pos = m.slot[c * NLF + lf];
                static_for<0, NF>([&](auto ic) {
                    constexpr int I = decltype(ic)::value;
                    for (int a = 0; a < NLOC; ++a) {
                        double *row = vals + rowptr[(int64_t)I * fstride + c * NLOC + a];
                        for (int j = 0; j < NF; ++j)
                            if (NBR_BLOCK[I][j]) {
                                memcpy(row + block_offset<I>(j, ncol, pos), Anb[I][a][j], sizeof(double) * NLOC);
                                if (vs)
                                    for (int b = 0; b < NLOC; ++b)
                                        (vs + (row - vals))[block_offset<I>(j, ncol, pos) + b] = fabs(Anb[I][a][j][b]);
                            }
                    }
                });
            }
#endif
        }
        fold<NR>(&R[0][0], &Rt[0][0], &Rabs[0][0]);
        fold<NA>(&Aown[0][0][0][0], &At[0][0][0][0], &Aabs[0][0][0][0]);
    }
    if (doJ) {
        const int pos = m.slot_self[c];
        static_for<0, NF>([&](auto ic) {
            constexpr int I = decltype(ic)::value;
            for (int a = 0; a < NLOC; ++a) {
                double *row = vals + rowptr[(int64_t)I * fstride + c * NLOC + a];
                for (int j = 0; j < NF; ++j)
                    if (OWN_BLOCK[I][j]) {
                        memcpy(row + block_offset<I>(j, ncol, pos), At[I][a][j], sizeof(double) * NLOC);
                        if (vs) memcpy(vs + (row - vals) + block_offset<I>(j, ncol, pos), Aabs[I][a][j], sizeof(double) * NLOC);
                    }
            }
        });
    }
    if (doF)
        for (int i = 0; i < NF; ++i)
            for (int a = 0; a < NLOC; ++a) {
                F[i * fstride + c * NLOC + a] = Rt[i][a];
                if (Fs) Fs[i * fstride + c * NLOC + a] = Rabs[i][a];
            }
}

}  // namespace port

extern "C" {

struct cpu_port_mesh {
    int64_t ncell, ncell_total;
    const double *xcell;
    const int32_t *nbr;
    const uint8_t *code, *slot, *slot_self, *ncol;
    const double *cell_volume, *cell_diam, *facet_area;
};

// d, p, nf, naux, nconst, nq, nclasses, family
int cpu_port_describe(int32_t *info) {
    info[0] = echgen::D;
    info[1] = echgen::P;
    info[2] = echgen::NF;
    info[3] = ECH_NAUX;
    info[4] = ECH_NCONST;
    info[5] = echgen::NQ;
    info[6] = ECH_NCLS;
    info[7] = ECH_FAMILY_DG;
    return 8;
}

// Residual into F (if non-NULL) and CSR Jacobian values into vals (if non-NULL) for the cells [cell_begin, cell_end);
// vals_scale / F_scale (optional): element-level scales of the same entries (sums of absolute contributions);
// nthreads <= 0: OpenMP default.  All pointers are host pointers; layouts are those of include/echem_b200_module.h.
int cpu_port_assemble(const cpu_port_mesh *m, const double *u, const double *aux, const double *consts,
                      const int64_t *rowptr, double *vals, double *F, double *vals_scale, double *F_scale,
                      int64_t cell_begin, int64_t cell_end, int nthreads) {
    port::Mesh M{m->ncell, m->ncell_total, m->xcell, m->nbr, m->code, m->slot, m->slot_self, m->ncol,
                 m->cell_volume, m->cell_diam, m->facet_area};
    if (cell_end > m->ncell) cell_end = m->ncell;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t c = cell_begin; c < cell_end; ++c) port::assemble_cell(M, c, u, aux, consts, rowptr, vals, F, vals ? vals_scale : nullptr, F ? F_scale : nullptr);
    return 0;
}

int cpu_port_max_threads(void) { return omp_get_max_threads(); }
}
