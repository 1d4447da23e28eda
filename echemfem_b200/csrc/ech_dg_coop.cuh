// Cooperative DG assembly kernels: the NLOC lanes that own the rows of one cell share the
// quadrature-point work of that cell through shared memory.
//
// ncu on the row-per-thread kernels of ech_dg.cuh (profiles/r1a_*) showed the FP64 pipe, not HBM,
// as the limiter (45 % fp64 pipe active at 14 % DRAM), with ~3/4 of the FP64 instructions spent on
// geometry / basis / interpolation / pointwise physics that every row of a cell recomputed.  Here:
//
//   lane (g, a) of a warp  <->  cell g of the warp, test node a          (G = NLOC lanes per cell)
//   round r:  lane a EVALUATES quadrature task r*G + a of its cell (cell point or facet point:
//             geometry, bases of both sides, state, pointwise f / g coefficients for ALL equations)
//             and writes one record to shared memory;  __syncwarp();
//             then every lane ACCUMULATES, for its own rows (all equations, test node a), the
//             contributions of the G records of its cell.
//
// so each point is evaluated once per cell instead of once per row.  Records are laid out
// [slot][field][cell-in-warp]: the lanes of one cell read the same word (broadcast) and the
// cells of a warp read consecutive words (no bank conflicts).  Only warp-level synchronisation
// is used.  Everything else (owner-computes rows, exactly-once vector stores, constant-memory
// 1D tables, compile-time coupling masks) is as in ech_dg.cuh.
#pragma once
#include "ech_dg.cuh"

namespace echdg {

constexpr int G = NLOC;
constexpr bool COOP_OK = NLOC <= 32;
constexpr int CPW = COOP_OK ? 32 / NLOC : 1;           // cells per warp
constexpr int NTASK = NQC + NLF * NQF;
constexpr int NROUND = (NTASK + G - 1) / G;

// record layout (in doubles)
constexpr int ROWJ_LEN = NF * (1 + 2 * D + D * D);
constexpr int OFF_PHI = 0;
constexpr int OFF_GPHI = NLOC;
constexpr int OFF_PHI2 = NLOC * (1 + D);
constexpr int OFF_GPHI2 = NLOC * (2 + D);
constexpr int OFF_JP = 2 * NLOC * (1 + D);
constexpr int OFF_JM = OFF_JP + NF * ROWJ_LEN;
constexpr int JREC_LEN = (OFF_JM + NF * ROWJ_LEN) | 1;
constexpr int RREC_LEN = (NLOC * (1 + D) + NF * (1 + D)) | 1;

__host__ __device__ constexpr int o_g0(int j) { return j; }
__host__ __device__ constexpr int o_g1(int j, int l) { return NF + j * D + l; }
__host__ __device__ constexpr int o_g2(int j, int k) { return NF + NF * D + j * D + k; }
__host__ __device__ constexpr int o_g3(int j, int k, int l) { return NF + 2 * NF * D + (j * D + k) * D + l; }

constexpr int coop_warps(int reclen) {
    const int per_warp = reclen * 32 * 8;               // bytes: G * CPW <= 32 records
    int w = (100 * 1024) / per_warp;
    return w < 1 ? 1 : (w > 4 ? 4 : w);
}
constexpr int JWARPS = coop_warps(JREC_LEN);
constexpr int RWARPS = coop_warps(RREC_LEN);

// accessor of one record: field k of (slot s, cell g) of this warp
struct RecRef {
    double *base;   // warp base + g
    int reclen;
    __device__ __forceinline__ double &operator()(int slot, int k) const { return base[(slot * reclen + k) * CPW]; }
};

// evaluation context of one lane
struct CellCtx {
    double X[NV][D];
    double U[NF][NLOC];
    double AX[ECH_NA1][NLOC];
};

template <class M, int I>
__device__ __forceinline__ void put_rowj(const RecRef &R, int slot, int off, const RowJ &r, double w) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (M::G0[I][j]) R(slot, off + o_g0(j)) = w * r.g0[j];
        if constexpr (M::G1[I][j]) {
#pragma unroll
            for (int l = 0; l < D; ++l) R(slot, off + o_g1(j, l)) = w * r.g1[j][l];
        }
        if constexpr (M::G2[I][j]) {
#pragma unroll
            for (int k = 0; k < D; ++k) R(slot, off + o_g2(j, k)) = w * r.g2[j][k];
        }
        if constexpr (M::G3[I][j] == 1) R(slot, off + o_g3(j, 0, 0)) = w * r.g3[j][0][0];
        if constexpr (M::G3[I][j] == 2) {
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int l = 0; l < D; ++l) R(slot, off + o_g3(j, k, l)) = w * r.g3[j][k][l];
        }
    });
}

__device__ __forceinline__ void put_basis(const RecRef &R, int slot, int off_phi, int off_gphi,
                                          const double (&phi)[NLOC], const double (&gphi)[NLOC][D]) {
#pragma unroll
    for (int b = 0; b < NLOC; ++b) {
        R(slot, off_phi + b) = phi[b];
#pragma unroll
        for (int k = 0; k < D; ++k) R(slot, off_gphi + b * D + k) = gphi[b][k];
    }
}

// rows (I, a) += record, for the trial basis stored at (off_phi, off_gphi) and coefficients at off
template <class M, int I>
__device__ __forceinline__ void take_rowj(const RecRef &R, int slot, int off, double phia, const double (&gphia)[D],
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
            if constexpr (h0) Aj += R(slot, off + o_g0(j)) * phia;
            if constexpr (h2) {
#pragma unroll
                for (int k = 0; k < D; ++k) Aj += R(slot, off + o_g2(j, k)) * gphia[k];
            }
            if constexpr (h1) {
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += R(slot, off + o_g1(j, l)) * phia;
            }
            if constexpr (h3 == 1) {
                const double s = R(slot, off + o_g3(j, 0, 0));
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += s * gphia[l];
            }
            if constexpr (h3 == 2) {
#pragma unroll
                for (int l = 0; l < D; ++l)
#pragma unroll
                    for (int k = 0; k < D; ++k) Bj[l] += R(slot, off + o_g3(j, k, l)) * gphia[k];
            }
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

__device__ __forceinline__ void get_basis(const RecRef &R, int slot, int off_phi, int off_gphi, double (&phi)[NLOC],
                                          double (&gphi)[NLOC][D]) {
#pragma unroll
    for (int b = 0; b < NLOC; ++b) {
        phi[b] = R(slot, off_phi + b);
#pragma unroll
        for (int k = 0; k < D; ++k) gphi[b][k] = R(slot, off_gphi + b * D + k);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(JWARPS * 32) jacobian_coop_kernel(const ech_dg_args A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, a = lane - g * G;
    const bool active_lane = g < CPW;
    const int64_t c = ((int64_t)blockIdx.x * JWARPS + warp) * CPW + g;
    const bool active = active_lane && c < A.ncell;
    const int64_t fstride = A.ncell_total * NLOC;
    RecRef R{smem + (size_t)warp * (G * JREC_LEN * CPW) + (active_lane ? g : 0), JREC_LEN};

    CellCtx cx;
    double cv = 0.0, cd = 0.0;
    int ncol = 0;
    if (active) {
        load_cell_coords(A.xcell, c, cx.X);
        load_cell_fields<NF>(A.u, fstride, c, cx.U);
        if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, c, cx.AX);
        cv = A.cell_volume[c];
        cd = A.cell_diam[c];
        ncol = A.ncol[c];
    }

    double acc[NF][NF][NLOC], accn[NF][NF][NLOC];
    static_for<0, NF>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        static_for<0, NF>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr bool own = OWN_BLOCK[I][j], nb = NBR_BLOCK[I][j];
            if constexpr (own) {
#pragma unroll
                for (int b = 0; b < NLOC; ++b) acc[I][j][b] = 0.0;
            }
            if constexpr (nb) {
#pragma unroll
                for (int b = 0; b < NLOC; ++b) accn[I][j][b] = 0.0;
            }
        });
    });

    const int64_t rows_per_field = A.ncell * NLOC;

    for (int round = 0; round < NROUND; ++round) {
        // ---------------- evaluate my task of this round
        const int task = round * G + a;
        if (active && task < NTASK) {
            Pt p;
            p.K = A.consts;
            p.cv = cv;
            p.cd = cd;
            Geo gm;
            double phi[NLOC], gphi[NLOC][D];
            int idx[D];
            double w;
            if (task < NQC) {
                cell_point(task, idx, w);
                geometry(cx.X, idx, gm);
                basis(idx, gm.invJ, phi, gphi);
                interpolate<NF>(cx.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cx.AX, phi, gphi, p.a, p.ga);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
                w *= fabs(gm.detJ);
                put_basis(R, a, OFF_PHI, OFF_GPHI, phi, gphi);
                static_for<0, NF>([&](auto ic) {
                    constexpr int I = decltype(ic)::value;
                    RowJ rj;
                    cell_jac<I>(p, rj);
                    put_rowj<MC, I>(R, a, OFF_JP + I * ROWJ_LEN, rj, w);
                });
            } else {
                const int ft = task - NQC;
                const int lf = ft / NQF, qf = ft - lf * NQF;
                const int nb = A.nbr[c * NLF + lf];
                p.fa = A.facet_area[c * NLF + lf];
                int tt[D > 1 ? D - 1 : 1];
                double ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(cx.X, idx, gm);
                basis(idx, gm.invJ, phi, gphi);
                interpolate<NF>(cx.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cx.AX, phi, gphi, p.a, p.ga);
                facet_normal(gm, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
                w *= ds;
                put_basis(R, a, OFF_PHI, OFF_GPHI, phi, gphi);
                if (nb < 0) {
                    const int cls = -nb - 1;
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        RowJ rj;
                        efacet_jac<I>(cls, p, rj);
                        put_rowj<ME, I>(R, a, OFF_JP + I * ROWJ_LEN, rj, w);
                    });
                } else {
                    if constexpr (ECH_FAMILY_DG) {
                        double X2[NV][D], U2[NF][NLOC], AX2[ECH_NA1][NLOC];
                        load_cell_coords(A.xcell, nb, X2);
                        load_cell_fields<NF>(A.u, fstride, nb, U2);
                        if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, nb, AX2);
                        p.cvm = A.cell_volume[nb];
                        p.cdm = A.cell_diam[nb];
                        int idx2[D];
                        neighbour_point(A.code[c * NLF + lf], tt, idx2);
                        Geo g2;
                        double phi2[NLOC], gphi2[NLOC][D];
                        geometry(X2, idx2, g2);
                        basis(idx2, g2.invJ, phi2, gphi2);
                        interpolate<NF>(U2, phi2, gphi2, p.um, p.gum);
                        if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX2, phi2, gphi2, p.am, p.gam);
                        put_basis(R, a, OFF_PHI2, OFF_GPHI2, phi2, gphi2);
                        static_for<0, NF>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            RowJ rj, rjm;
                            ifacet_jac<I>(p, rj, rjm);
                            put_rowj<MIP, I>(R, a, OFF_JP + I * ROWJ_LEN, rj, w);
                            put_rowj<MIM, I>(R, a, OFF_JM + I * ROWJ_LEN, rjm, w);
                        });
                    }
                }
            }
        }
        __syncwarp();
        // ---------------- accumulate the records of my cell into my rows
        if (active) {
            for (int s = 0; s < G; ++s) {
                const int tk = round * G + s;
                if (tk >= NTASK) break;
                double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
                get_basis(R, s, OFF_PHI, OFF_GPHI, phi, gphi);
                phia = R(s, OFF_PHI + a);
#pragma unroll
                for (int k = 0; k < D; ++k) gphia[k] = R(s, OFF_GPHI + a * D + k);
                if (tk < NQC) {
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        take_rowj<MC, I>(R, s, OFF_JP + I * ROWJ_LEN, phia, gphia, phi, gphi, acc[I]);
                    });
                } else {
                    const int ft = tk - NQC;
                    const int lf = ft / NQF, qf = ft - lf * NQF;
                    const int nb = A.nbr[c * NLF + lf];
                    if (nb < 0) {
                        static_for<0, NF>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            take_rowj<ME, I>(R, s, OFF_JP + I * ROWJ_LEN, phia, gphia, phi, gphi, acc[I]);
                        });
                    } else {
                        if constexpr (ECH_FAMILY_DG) {
                            static_for<0, NF>([&](auto ic) {
                                constexpr int I = decltype(ic)::value;
                                take_rowj<MIP, I>(R, s, OFF_JP + I * ROWJ_LEN, phia, gphia, phi, gphi, acc[I]);
                            });
                            get_basis(R, s, OFF_PHI2, OFF_GPHI2, phi, gphi);
                            static_for<0, NF>([&](auto ic) {
                                constexpr int I = decltype(ic)::value;
                                take_rowj<MIM, I>(R, s, OFF_JM + I * ROWJ_LEN, phia, gphia, phi, gphi, accn[I]);
                            });
                            if (qf == NQF - 1) {
                                const int pos = A.slot[c * NLF + lf];
                                static_for<0, NF>([&](auto ic) {
                                    constexpr int I = decltype(ic)::value;
                                    double *__restrict__ vals =
                                        A.out + A.rowptr[(int64_t)I * rows_per_field + c * NLOC + a];
                                    static_for<0, NF>([&](auto jc) {
                                        constexpr int j = decltype(jc)::value;
                                        constexpr bool nbj = NBR_BLOCK[I][j];
                                        if constexpr (nbj) {
                                            store_block(vals + block_offset<I, j>(ncol, pos), accn[I][j]);
#pragma unroll
                                            for (int b = 0; b < NLOC; ++b) accn[I][j][b] = 0.0;
                                        }
                                    });
                                });
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (active) {
        const int pos = A.slot_self[c];
        static_for<0, NF>([&](auto ic) {
            constexpr int I = decltype(ic)::value;
            double *__restrict__ vals = A.out + A.rowptr[(int64_t)I * rows_per_field + c * NLOC + a];
            static_for<0, NF>([&](auto jc) {
                constexpr int j = decltype(jc)::value;
                constexpr bool own = OWN_BLOCK[I][j];
                if constexpr (own) store_block(vals + block_offset<I, j>(ncol, pos), acc[I][j]);
            });
        });
    }
}

// ------------------------------------------------------------------------------------------------
// residual: records carry the test-side basis and w * (f0, f1) of every equation
constexpr int RO_F0 = NLOC * (1 + D);
constexpr int RO_F1 = RO_F0 + NF;

template <class M>
__device__ __forceinline__ void put_res(const RecRef &R, int slot, const Res &o, double w) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr bool h0 = M::F0[i], h1 = M::F1[i];
        if constexpr (h0)
            R(slot, RO_F0 + i) = w * o.f0[i];
        else
            R(slot, RO_F0 + i) = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if constexpr (h1)
                R(slot, RO_F1 + i * D + k) = w * o.f1[i][k];
            else
                R(slot, RO_F1 + i * D + k) = 0.0;
        }
    });
}

__global__ void __launch_bounds__(RWARPS * 32) residual_coop_kernel(const ech_dg_args A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, a = lane - g * G;
    const bool active_lane = g < CPW;
    const int64_t c = ((int64_t)blockIdx.x * RWARPS + warp) * CPW + g;
    const bool active = active_lane && c < A.ncell;
    const int64_t fstride = A.ncell_total * NLOC;
    RecRef R{smem + (size_t)warp * (G * RREC_LEN * CPW) + (active_lane ? g : 0), RREC_LEN};

    CellCtx cx;
    double cv = 0.0, cd = 0.0;
    if (active) {
        load_cell_coords(A.xcell, c, cx.X);
        load_cell_fields<NF>(A.u, fstride, c, cx.U);
        if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, c, cx.AX);
        cv = A.cell_volume[c];
        cd = A.cell_diam[c];
    }
    double r[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) r[i] = 0.0;

    for (int round = 0; round < NROUND; ++round) {
        const int task = round * G + a;
        if (active && task < NTASK) {
            Pt p;
            p.K = A.consts;
            p.cv = cv;
            p.cd = cd;
            Geo gm;
            double phi[NLOC], gphi[NLOC][D];
            int idx[D];
            double w;
            Res o;
            if (task < NQC) {
                cell_point(task, idx, w);
                geometry(cx.X, idx, gm);
                basis(idx, gm.invJ, phi, gphi);
                interpolate<NF>(cx.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cx.AX, phi, gphi, p.a, p.ga);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
                cell_res(p, o);
                put_basis(R, a, OFF_PHI, OFF_GPHI, phi, gphi);
                put_res<MC>(R, a, o, w * fabs(gm.detJ));
            } else {
                const int ft = task - NQC;
                const int lf = ft / NQF, qf = ft - lf * NQF;
                const int nb = A.nbr[c * NLF + lf];
                p.fa = A.facet_area[c * NLF + lf];
                int tt[D > 1 ? D - 1 : 1];
                double ds;
                facet_point(lf, qf, idx, tt, w);
                geometry(cx.X, idx, gm);
                basis(idx, gm.invJ, phi, gphi);
                interpolate<NF>(cx.U, phi, gphi, p.u, p.gu);
                if constexpr (NAUX > 0) interpolate<ECH_NA1>(cx.AX, phi, gphi, p.a, p.ga);
                facet_normal(gm, lf, p.n, ds);
#pragma unroll
                for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
                put_basis(R, a, OFF_PHI, OFF_GPHI, phi, gphi);
                if (nb < 0) {
                    efacet_res(-nb - 1, p, o);
                    put_res<ME>(R, a, o, w * ds);
                } else {
                    if constexpr (ECH_FAMILY_DG) {
                        double X2[NV][D], U2[NF][NLOC], AX2[ECH_NA1][NLOC];
                        load_cell_coords(A.xcell, nb, X2);
                        load_cell_fields<NF>(A.u, fstride, nb, U2);
                        if constexpr (NAUX > 0) load_cell_fields<ECH_NA1>(A.aux, fstride, nb, AX2);
                        p.cvm = A.cell_volume[nb];
                        p.cdm = A.cell_diam[nb];
                        int idx2[D];
                        neighbour_point(A.code[c * NLF + lf], tt, idx2);
                        Geo g2;
                        double phi2[NLOC], gphi2[NLOC][D];
                        geometry(X2, idx2, g2);
                        basis(idx2, g2.invJ, phi2, gphi2);
                        interpolate<NF>(U2, phi2, gphi2, p.um, p.gum);
                        if constexpr (NAUX > 0) interpolate<ECH_NA1>(AX2, phi2, gphi2, p.am, p.gam);
                        ifacet_res(p, o);
                        put_res<MIP>(R, a, o, w * ds);
                    }
                }
            }
        }
        __syncwarp();
        if (active) {
            for (int s = 0; s < G; ++s) {
                if (round * G + s >= NTASK) break;
                const double phia = R(s, OFF_PHI + a);
                double gphia[D];
#pragma unroll
                for (int k = 0; k < D; ++k) gphia[k] = R(s, OFF_GPHI + a * D + k);
#pragma unroll
                for (int i = 0; i < NF; ++i) {
                    double t = R(s, RO_F0 + i) * phia;
#pragma unroll
                    for (int k = 0; k < D; ++k) t += R(s, RO_F1 + i * D + k) * gphia[k];
                    r[i] += t;
                }
            }
        }
        __syncwarp();
    }
    if (active) {
        const int64_t ostride = A.ncell * NLOC;
#pragma unroll
        for (int i = 0; i < NF; ++i) A.out[i * ostride + c * NLOC + a] = r[i];
    }
}

// ------------------------------------------------------------------------------------------------
inline int coop_setup() {
    static int done = 0;
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(jacobian_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         JWARPS * G * JREC_LEN * CPW * 8);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(residual_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             RWARPS * G * RREC_LEN * CPW * 8);
    if (e != cudaSuccess) return (int)e;
    done = 1;
    return 0;
}

inline int launch_jacobian_coop(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    int e = coop_setup();
    if (e) return e;
    const int64_t cells_per_block = (int64_t)JWARPS * CPW;
    const int64_t grid = (a->ncell + cells_per_block - 1) / cells_per_block;
    jacobian_coop_kernel<<<(unsigned)grid, JWARPS * 32, JWARPS * G * JREC_LEN * CPW * 8, s>>>(*a);
    return (int)cudaGetLastError();
}

inline int launch_residual_coop(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    int e = coop_setup();
    if (e) return e;
    const int64_t cells_per_block = (int64_t)RWARPS * CPW;
    const int64_t grid = (a->ncell + cells_per_block - 1) / cells_per_block;
    residual_coop_kernel<<<(unsigned)grid, RWARPS * 32, RWARPS * G * RREC_LEN * CPW * 8, s>>>(*a);
    return (int)cudaGetLastError();
}

}  // namespace echdg
