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
#include <stdio.h>
#include "ech_dg.cuh"

namespace echdg {

constexpr int G = NLOC;
constexpr bool COOP_OK = NLOC <= 32;
constexpr int CPW = COOP_OK ? 32 / NLOC : 1;           // cells per warp
constexpr int NTASK = NQC + NLF * NQF;
constexpr int NROUND = (NTASK + G - 1) / G;

// ---- record layout (in doubles).  A record holds the trial bases of both sides followed by the
// weight-scaled coefficients g0..g3 of every equation, COMPACTED by the compile-time coupling masks.
template <class M>
struct Layout {
    int o0[NF][NF], o1[NF][NF], o2[NF][NF], o3[NF][NF];
    int len;
    __host__ __device__ constexpr Layout() : o0{}, o1{}, o2{}, o3{}, len(0) {
        int n = 0;
        for (int i = 0; i < NF; ++i)
            for (int j = 0; j < NF; ++j) {
                o0[i][j] = o1[i][j] = o2[i][j] = o3[i][j] = -1;
                // scalars first would leave holes; D-vectors start on even offsets (128-bit shared access)
                if (M::G1[i][j]) { n = (D == 2) ? ((n + 1) & ~1) : n; o1[i][j] = n; n += D; }
                if (M::G2[i][j]) { n = (D == 2) ? ((n + 1) & ~1) : n; o2[i][j] = n; n += D; }
                if (M::G3[i][j] == 2) { n = (D == 2) ? ((n + 1) & ~1) : n; o3[i][j] = n; n += D * D; }
                if (M::G0[i][j]) { o0[i][j] = n; n += 1; }
                if (M::G3[i][j] == 1) { o3[i][j] = n; n += 1; }
            }
        len = n;
    }
};
constexpr Layout<MC> LAY_C{};
constexpr Layout<ME> LAY_E{};
constexpr Layout<MIP> LAY_IP{};
constexpr Layout<MIM> LAY_IM{};
template <class M> struct LayoutOf;
template <> struct LayoutOf<MC> { static __host__ __device__ constexpr const Layout<MC> &get() { return LAY_C; } };
template <> struct LayoutOf<ME> { static __host__ __device__ constexpr const Layout<ME> &get() { return LAY_E; } };
template <> struct LayoutOf<MIP> { static __host__ __device__ constexpr const Layout<MIP> &get() { return LAY_IP; } };
template <> struct LayoutOf<MIM> { static __host__ __device__ constexpr const Layout<MIM> &get() { return LAY_IM; } };

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int OFF_PHI = 0;
constexpr int OFF_GPHI = NLOC;
constexpr int OFF_PHI2 = NLOC * (1 + D);
constexpr int OFF_GPHI2 = NLOC * (2 + D);
constexpr int OFF_JP = 2 * NLOC * (1 + D);
constexpr int OFF_JM = OFF_JP + ((LAY_IP.len + 1) & ~1);
constexpr int JREC_LEN = (cmax(cmax(OFF_JP + LAY_C.len, OFF_JP + LAY_E.len), OFF_JM + LAY_IM.len) + 1) & ~1;
constexpr int RREC_LEN = (NLOC * (1 + D) + NF * (1 + D) + 1) & ~1;

// neighbour data staged per (cell, facet): vertex coordinates, state, aux, volume, diameter
constexpr int NBD_X = 0;
constexpr int NBD_U = NV * D;
constexpr int NBD_A = NBD_U + NF * NLOC;
constexpr int NBD_CV = NBD_A + ECH_NAUX * NLOC;
constexpr int NBD_LEN = (NBD_CV + 2) | 1;
// per-cell facet metadata (ints): nbr, code, slot
constexpr int META_LEN = 3 * NLF;

constexpr int NBD_SLOTS = NLF + 1;   // neighbours + the cell itself
__host__ __device__ constexpr int warp_smem_doubles(int reclen) { return (reclen * 32 + CPW * NBD_SLOTS * NBD_LEN + (CPW * META_LEN + 1) / 2 + 1) & ~1; }
constexpr int coop_warps(int reclen) {
    const int per_warp = warp_smem_doubles(reclen) * 8;
#ifdef ECH_COOP_WARPS
    return ECH_COOP_WARPS + 0 * per_warp;
#else
    int w = (56 * 1024) / per_warp;
    return w < 1 ? 1 : (w > 4 ? 4 : w);
#endif
}
constexpr int JWARPS = coop_warps(JREC_LEN);
constexpr int RWARPS = coop_warps(RREC_LEN);

// accessor of one record: field k of (slot s, cell g) lives at k*32 + s*CPW + g, so that the lanes of
// a warp write / read consecutive words (no bank conflicts; same-cell lanes broadcast)
// Record addressing.  Fields are stored as 16-byte pairs; pair m of the record in slot s of cell g
// lives at 16 * (m*32 + s*CPW + g) bytes.
//  * PRODUCER role (evaluate): lane L evaluates slot L / CPW of cell L % CPW, so a warp-wide 128-bit
//    store of pair m is one contiguous 512-byte row -> conflict-free.
//  * CONSUMER role (accumulate): lane L owns test node L % G of cell L / G; a 128-bit load of slot s
//    touches the CPW consecutive pairs of one 128-byte row, same-cell lanes broadcast.
// Measured (accumulate / evaluate phase, ms): 64-bit words 2.15 / 0.54; one lane numbering for both
// roles: cell-minor 2.2 / 0.53, cell-major 0.9 / 0.89; split roles: see profiles/.
struct RecRef {
    double *base;   // warp base + 2 * g
    __device__ __forceinline__ double &operator()(int slot, int k) const {
        return base[((k >> 1) * 32 + slot * CPW) * 2 + (k & 1)];
    }
    __device__ __forceinline__ double2 pair(int slot, int k) const {
        return *reinterpret_cast<const double2 *>(base + ((k >> 1) * 32 + slot * CPW) * 2);
    }
    __device__ __forceinline__ void set_pair(int slot, int k, double x, double y) const {
        *reinterpret_cast<double2 *>(base + ((k >> 1) * 32 + slot * CPW) * 2) = make_double2(x, y);
    }
};

// evaluation context of one lane
struct CellCtx {
    double X[NV][D];
    double U[NF][NLOC];
    double AX[ECH_NA1][NLOC];
};

// the producer assembles its record in registers (compile-time offsets) and flushes it as 128-bit pairs,
// so that every shared-memory store of a warp is conflict-free
template <class M, int I, int LEN>
__device__ __forceinline__ void put_rowj(double (&buf)[LEN], int off, const RowJ &r, double w) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int p0 = LayoutOf<M>::get().o0[I][j], p1 = LayoutOf<M>::get().o1[I][j];
        constexpr int p2 = LayoutOf<M>::get().o2[I][j], p3 = LayoutOf<M>::get().o3[I][j];
        constexpr int h3 = M::G3[I][j];
        if constexpr (p0 >= 0) buf[off + p0] = w * r.g0[j];
        if constexpr (p1 >= 0) {
#pragma unroll
            for (int l = 0; l < D; ++l) buf[off + p1 + l] = w * r.g1[j][l];
        }
        if constexpr (p2 >= 0) {
#pragma unroll
            for (int k = 0; k < D; ++k) buf[off + p2 + k] = w * r.g2[j][k];
        }
        if constexpr (h3 == 1) buf[off + p3] = w * r.g3[j][0][0];
        if constexpr (h3 == 2) {
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int l = 0; l < D; ++l) buf[off + p3 + k * D + l] = w * r.g3[j][k][l];
        }
    });
}

template <int LEN>
__device__ __forceinline__ void flush_record(const RecRef &R, int slot, const double (&buf)[LEN], int from, int to) {
#pragma unroll
    for (int m = 0; m < LEN; m += 2) {
        if (m >= from && m < to) R.set_pair(slot, m, buf[m], buf[m + 1]);
    }
}

template <int LEN>
__device__ __forceinline__ void buf_basis(double (&buf)[LEN], int off_phi, int off_gphi, const double (&phi)[NLOC],
                                          const double (&gphi)[NLOC][D]) {
#pragma unroll
    for (int b = 0; b < NLOC; ++b) {
        buf[off_phi + b] = phi[b];
#pragma unroll
        for (int k = 0; k < D; ++k) buf[off_gphi + b * D + k] = gphi[b][k];
    }
}

__device__ __forceinline__ void put_basis(const RecRef &R, int slot, int off_phi, int off_gphi,
                                          const double (&phi)[NLOC], const double (&gphi)[NLOC][D]) {
    // off_phi and off_gphi are even when NLOC is even; both blocks are then written as pairs
    if constexpr (NLOC % 2 == 0) {
#pragma unroll
        for (int b = 0; b < NLOC; b += 2) R.set_pair(slot, off_phi + b, phi[b], phi[b + 1]);
        const double *gp = &gphi[0][0];
#pragma unroll
        for (int m = 0; m < NLOC * D; m += 2) R.set_pair(slot, off_gphi + m, gp[m], gp[m + 1]);
    } else {
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            R(slot, off_phi + b) = phi[b];
#pragma unroll
            for (int k = 0; k < D; ++k) R(slot, off_gphi + b * D + k) = gphi[b][k];
        }
    }
}

// rows (I, a) += record, for the trial basis (phi, gphi) and the coefficients stored at off
template <class M, int I>
__device__ __forceinline__ void take_rowj(const RecRef &R, int slot, int off, double phia, const double (&gphia)[D],
                                          const double (&phi)[NLOC], const double (&gphi)[NLOC][D],
                                          double (&acc)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int p0 = LayoutOf<M>::get().o0[I][j], p1 = LayoutOf<M>::get().o1[I][j];
        constexpr int p2 = LayoutOf<M>::get().o2[I][j], p3 = LayoutOf<M>::get().o3[I][j];
        constexpr bool h0 = p0 >= 0, h1 = p1 >= 0, h2 = p2 >= 0;
        constexpr int h3 = M::G3[I][j];
        if constexpr (h0 || h1 || h2 || h3 != 0) {
            double Aj = 0.0, Bj[D];
#pragma unroll
            for (int l = 0; l < D; ++l) Bj[l] = 0.0;
            if constexpr (h0) Aj += R(slot, off + p0) * phia;
            if constexpr (h2) {
                if constexpr (D == 2) {
                    const double2 v = R.pair(slot, off + p2);
                    Aj += v.x * gphia[0];
                    Aj += v.y * gphia[1];
                } else {
#pragma unroll
                    for (int k = 0; k < D; ++k) Aj += R(slot, off + p2 + k) * gphia[k];
                }
            }
            if constexpr (h1) {
                if constexpr (D == 2) {
                    const double2 v = R.pair(slot, off + p1);
                    Bj[0] += v.x * phia;
                    Bj[1] += v.y * phia;
                } else {
#pragma unroll
                    for (int l = 0; l < D; ++l) Bj[l] += R(slot, off + p1 + l) * phia;
                }
            }
            if constexpr (h3 == 1) {
                const double sc = R(slot, off + p3);
#pragma unroll
                for (int l = 0; l < D; ++l) Bj[l] += sc * gphia[l];
            }
            if constexpr (h3 == 2) {
#pragma unroll
                for (int l = 0; l < D; ++l)
#pragma unroll
                    for (int k = 0; k < D; ++k) Bj[l] += R(slot, off + p3 + k * D + l) * gphia[k];
            }
#pragma unroll
            for (int b = 0; b < NLOC; ++b) {
                double sacc = acc[j][b];
                if constexpr (h0 || h2) sacc += Aj * phi[b];
                if constexpr (h1 || h3 != 0) {
#pragma unroll
                    for (int l = 0; l < D; ++l) sacc += Bj[l] * gphi[b][l];
                }
                acc[j][b] = sacc;
            }
        }
    });
}

__device__ __forceinline__ void get_basis(const RecRef &R, int slot, int off_phi, int off_gphi, double (&phi)[NLOC],
                                          double (&gphi)[NLOC][D]) {
    if constexpr (NLOC % 2 == 0) {
#pragma unroll
        for (int b = 0; b < NLOC; b += 2) {
            const double2 v = R.pair(slot, off_phi + b);
            phi[b] = v.x;
            phi[b + 1] = v.y;
        }
        double *gp = &gphi[0][0];
#pragma unroll
        for (int m = 0; m < NLOC * D; m += 2) {
            const double2 v = R.pair(slot, off_gphi + m);
            gp[m] = v.x;
            gp[m + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            phi[b] = R(slot, off_phi + b);
#pragma unroll
            for (int k = 0; k < D; ++k) gphi[b][k] = R(slot, off_gphi + b * D + k);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-warp shared memory: [records | neighbour data | facet metadata]
struct WarpSmem {
    double *rec;     // reclen * 32
    double *nbd;     // [CPW][NLF][NBD_LEN]
    int *meta;       // [CPW][3][NLF]   nbr, code, slot
};

__device__ __forceinline__ WarpSmem warp_smem(double *smem, int warp, int reclen) {
    double *base = smem + (size_t)warp * warp_smem_doubles(reclen);
    WarpSmem w;
    w.rec = base;
    w.nbd = base + reclen * 32;
    w.meta = reinterpret_cast<int *>(w.nbd + CPW * NBD_SLOTS * NBD_LEN);
    return w;
}

// stage facet metadata and the neighbours' data of my cell (lanes of a cell split the facets);
// all global loads are issued back to back, so their latency is paid once per warp
__device__ __forceinline__ void stage_neighbours(const ech_dg_args &A, const WarpSmem &W, int g, int a, int64_t c,
                                                 bool active, int64_t fstride) {
    if (active) {
        for (int lf = a; lf < NBD_SLOTS; lf += G) {
            int nb;
            if (lf < NLF) {
                nb = A.nbr[c * NLF + lf];
                int *m = W.meta + g * META_LEN;
                m[lf] = nb;
                m[NLF + lf] = A.code[c * NLF + lf];
                m[2 * NLF + lf] = A.slot[c * NLF + lf];
            } else {
                nb = (int)c;
            }
            double *d = W.nbd + (g * NBD_SLOTS + lf) * NBD_LEN;
            if (nb >= 0) {
                const double2 *px = reinterpret_cast<const double2 *>(A.xcell + (int64_t)nb * (NV * D));
#pragma unroll
                for (int i = 0; i < NV * D / 2; ++i) {
                    const double2 v = __ldg(px + i);
                    d[NBD_X + 2 * i] = v.x;
                    d[NBD_X + 2 * i + 1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int b = 0; b < NLOC; ++b) d[NBD_U + j * NLOC + b] = __ldg(A.u + j * fstride + (int64_t)nb * NLOC + b);
                if constexpr (NAUX > 0) {
#pragma unroll
                    for (int j = 0; j < NAUX; ++j)
#pragma unroll
                        for (int b = 0; b < NLOC; ++b)
                            d[NBD_A + j * NLOC + b] = __ldg(A.aux + j * fstride + (int64_t)nb * NLOC + b);
                }
                d[NBD_CV] = A.cell_volume[nb];
                d[NBD_CV + 1] = A.cell_diam[nb];
            }
        }
    }
    __syncwarp();
}

// own-side evaluation at a table point from the staged cell data
__device__ __forceinline__ void own_side(const double *d, const int (&idx)[D], Pt &p, Geo &gm, double (&phi)[NLOC],
                                         double (&gphi)[NLOC][D]) {
    double X[NV][D], U[NF][NLOC];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < D; ++k) X[v][k] = d[NBD_X + v * D + k];
#pragma unroll
    for (int j = 0; j < NF; ++j)
#pragma unroll
        for (int b = 0; b < NLOC; ++b) U[j][b] = d[NBD_U + j * NLOC + b];
    p.cv = d[NBD_CV];
    p.cd = d[NBD_CV + 1];
    geometry(X, idx, gm);
    basis(idx, gm.invJ, phi, gphi);
    interpolate<NF>(U, phi, gphi, p.u, p.gu);
    if constexpr (NAUX > 0) {
        double AX[ECH_NA1][NLOC];
#pragma unroll
        for (int j = 0; j < ECH_NA1; ++j)
#pragma unroll
            for (int b = 0; b < NLOC; ++b) AX[j][b] = d[NBD_A + j * NLOC + b];
        interpolate<ECH_NA1>(AX, phi, gphi, p.a, p.ga);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
}

template <int NFLD>
__device__ __forceinline__ void read_fields(const double *d, double (&U)[NFLD][NLOC]) {
#pragma unroll
    for (int j = 0; j < NFLD; ++j)
#pragma unroll
        for (int b = 0; b < NLOC; ++b) U[j][b] = d[j * NLOC + b];
}

// evaluate the neighbour side of a facet point from the staged data
__device__ __forceinline__ void neighbour_side(const double *d, int code, const int (&tt)[D > 1 ? D - 1 : 1], Pt &p,
                                               double (&phi2)[NLOC], double (&gphi2)[NLOC][D]) {
    double X2[NV][D], U2[NF][NLOC];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < D; ++k) X2[v][k] = d[NBD_X + v * D + k];
    read_fields<NF>(d + NBD_U, U2);
    p.cvm = d[NBD_CV];
    p.cdm = d[NBD_CV + 1];
    int idx2[D];
    neighbour_point(code, tt, idx2);
    Geo g2;
    geometry(X2, idx2, g2);
    basis(idx2, g2.invJ, phi2, gphi2);
    interpolate<NF>(U2, phi2, gphi2, p.um, p.gum);
    if constexpr (NAUX > 0) {
        double AX2[ECH_NA1][NLOC];
        read_fields<ECH_NA1>(d + NBD_A, AX2);
        interpolate<ECH_NA1>(AX2, phi2, gphi2, p.am, p.gam);
    }
}

#ifndef ECH_JMINB
#define ECH_JMINB 1
#endif
// all records of a facet fall into one round, and rounds do not mix cell and facet records
constexpr bool FACET_IN_ROUND = (NQC % G == 0) && (G % NQF == 0);

template <class M>
__device__ __forceinline__ void take_all(const RecRef &R, int s, int off, double phia, const double (&gphia)[D],
                                         const double (&phi)[NLOC], const double (&gphi)[NLOC][D],
                                         double (&acc)[NF][NF][NLOC]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        take_rowj<M, I>(R, s, off, phia, gphia, phi, gphi, acc[I]);
    });
}

template <bool OWN>
__device__ __forceinline__ void zero_blocks(double (&acc)[NF][NF][NLOC]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        static_for<0, NF>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr bool on = OWN ? OWN_BLOCK[I][j] : NBR_BLOCK[I][j];
            if constexpr (on) {
#pragma unroll
                for (int b = 0; b < NLOC; ++b) acc[I][j][b] = 0.0;
            }
        });
    });
}

template <bool OWN>
__device__ __forceinline__ void store_blocks(double *__restrict__ out, const int64_t (&rp)[NF], int ncol, int pos,
                                             const double (&acc)[NF][NF][NLOC]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        double *__restrict__ vals = out + rp[I];
        static_for<0, NF>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr bool on = OWN ? OWN_BLOCK[I][j] : NBR_BLOCK[I][j];
#ifdef ECH_DBG_NO_STORE
            if constexpr (on) { if (acc[I][j][0] == 1.2345e300) store_block(vals + block_offset<I, j>(ncol, pos), acc[I][j]); }
#else
            if constexpr (on) store_block(vals + block_offset<I, j>(ncol, pos), acc[I][j]);
#endif
        });
    });
}

__device__ __forceinline__ void test_basis(const RecRef &R, int s, int a, double &phia, double (&gphia)[D]) {
    phia = R(s, OFF_PHI + a);
#pragma unroll
    for (int k = 0; k < D; ++k) gphia[k] = R(s, OFF_GPHI + a * D + k);
}

// WITH_RES: fused evaluation of the residual at the same state (one Newton step needs F(u) and J(u);
// geometry, bases, state interpolation and staging are shared, the residual costs 6 more record words)
constexpr int RES_LEN = (NF * (1 + D) + 1) & ~1;
constexpr int OFF_RES = JREC_LEN;
constexpr int JFREC_LEN = JREC_LEN + RES_LEN;

template <class M, int LEN>
__device__ __forceinline__ void put_res_at(double (&buf)[LEN], const Res &o, double w) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr bool h0 = M::F0[i], h1 = M::F1[i];
        if constexpr (h0) buf[OFF_RES + i] = w * o.f0[i];
        if constexpr (h1) {
#pragma unroll
            for (int k = 0; k < D; ++k) buf[OFF_RES + NF + i * D + k] = w * o.f1[i][k];
        }
    });
}

template <bool WITH_RES>
__global__ void __launch_bounds__(JWARPS * 32, ECH_JMINB) jacobian_coop_kernel(const ech_dg_args A) {
    constexpr int RECLEN = WITH_RES ? JFREC_LEN : JREC_LEN;
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, a = lane - g * G;       // cell-major lane numbering
    const bool active_lane = g < CPW;
    const int64_t c = A.cell_begin + ((int64_t)blockIdx.x * JWARPS + warp) * CPW + g;
    const bool active = active_lane && c < A.ncell;
    const int64_t fstride = A.ncell_total * NLOC;
    const int64_t rows_per_field = A.ncell_total * NLOC;
    const WarpSmem W = warp_smem(smem, warp, RECLEN);
    const RecRef R{W.rec + 2 * (active_lane ? g : 0)};
    const int gg = active_lane ? g : 0;
    const int *meta = W.meta + gg * META_LEN;
    // producer role: this lane evaluates slot ae of cell ge of the warp
    const int ge = lane % CPW, ae = lane / CPW;
    const int64_t ce = A.cell_begin + ((int64_t)blockIdx.x * JWARPS + warp) * CPW + ge;
    const bool active_e = ae < G && ce < A.ncell;
    const RecRef RE{W.rec + 2 * ge};
    const int *meta_e = W.meta + ge * META_LEN;
    const double *own = W.nbd + (ge * NBD_SLOTS + NLF) * NBD_LEN;

    int ncol = 0, pos_self = 0;
    int64_t rp[NF];
    if (active) {
        ncol = A.ncol[c];
        pos_self = A.slot_self[c];
#pragma unroll
        for (int i = 0; i < NF; ++i) rp[i] = A.rowptr[(int64_t)i * rows_per_field + c * NLOC + a];
    }
    stage_neighbours(A, W, g, a, c, active, fstride);

    double acc[NF][NF][NLOC];
    zero_blocks<true>(acc);
    double res[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) res[i] = 0.0;
    double accn_persist[FACET_IN_ROUND ? 1 : NF][FACET_IN_ROUND ? 1 : NF][FACET_IN_ROUND ? 1 : NLOC];
    if constexpr (!FACET_IN_ROUND) zero_blocks<false>(reinterpret_cast<double(&)[NF][NF][NLOC]>(accn_persist));

    for (int round = 0; round < NROUND; ++round) {
        // ---------------- evaluate my task of this round
        const int task = round * G + ae;
#ifdef ECH_DBG_NO_EVAL
        if (false) {
#else
        if (active_e && task < NTASK) {
#endif
            Pt p;
            p.K = A.consts;
            Geo gm;
            double phi[NLOC], gphi[NLOC][D];
            int idx[D];
            double w;
            double buf[RECLEN];
#pragma unroll
            for (int m = 0; m < RECLEN; ++m) buf[m] = 0.0;
            if (task < NQC) {
                cell_point(task, idx, w);
                own_side(own, idx, p, gm, phi, gphi);
                w *= fabs(gm.detJ);
                buf_basis(buf, OFF_PHI, OFF_GPHI, phi, gphi);
                static_for<0, NF>([&](auto ic) {
                    constexpr int I = decltype(ic)::value;
                    RowJ rj;
                    cell_jac<I>(p, rj);
                    put_rowj<MC, I>(buf, OFF_JP, rj, w);
                });
                if constexpr (WITH_RES) {
                    Res o;
                    cell_res(p, o);
                    put_res_at<MC>(buf, o, w);
                    flush_record(RE, ae, buf, OFF_RES, OFF_RES + RES_LEN);
                }
                flush_record(RE, ae, buf, OFF_PHI, OFF_PHI2);
                flush_record(RE, ae, buf, OFF_JP, OFF_JP + ((LAY_C.len + 1) & ~1));
            } else {
                const int ft = task - NQC;
                const int lf = ft / NQF, qf = ft - lf * NQF;
                const int nb = meta_e[lf];
                int tt[D > 1 ? D - 1 : 1];
                double ds;
                facet_point(lf, qf, idx, tt, w);
                if (nb >= 0) {
                    if constexpr (ECH_FAMILY_DG) {
                        double phi2[NLOC], gphi2[NLOC][D];
                        neighbour_side(W.nbd + (ge * NBD_SLOTS + lf) * NBD_LEN, meta_e[NLF + lf], tt, p, phi2, gphi2);
                        buf_basis(buf, OFF_PHI2, OFF_GPHI2, phi2, gphi2);
                    }
                }
                own_side(own, idx, p, gm, phi, gphi);
                p.fa = A.facet_area[ce * NLF + lf];
                facet_normal(gm, lf, p.n, ds);
                w *= ds;
                buf_basis(buf, OFF_PHI, OFF_GPHI, phi, gphi);
                if (nb < 0) {
                    const int cls = -nb - 1;
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        RowJ rj;
                        efacet_jac<I>(cls, p, rj);
                        put_rowj<ME, I>(buf, OFF_JP, rj, w);
                    });
                    if constexpr (WITH_RES) {
                        Res o;
                        efacet_res(cls, p, o);
                        put_res_at<ME>(buf, o, w);
                        flush_record(RE, ae, buf, OFF_RES, OFF_RES + RES_LEN);
                    }
                    flush_record(RE, ae, buf, OFF_PHI, OFF_PHI2);
                    flush_record(RE, ae, buf, OFF_JP, OFF_JP + ((LAY_E.len + 1) & ~1));
                } else {
                    if constexpr (ECH_FAMILY_DG) {
                        static_for<0, NF>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            RowJ rj, rjm;
                            ifacet_jac<I>(p, rj, rjm);
                            put_rowj<MIP, I>(buf, OFF_JP, rj, w);
                            put_rowj<MIM, I>(buf, OFF_JM, rjm, w);
                        });
                        if constexpr (WITH_RES) {
                            Res o;
                            ifacet_res(p, o);
                            put_res_at<MIP>(buf, o, w);
                        }
                        flush_record(RE, ae, buf, 0, RECLEN);
                    }
                }
            }
        }
        __syncwarp();
        // ---------------- accumulate the records of my cell into my rows
#ifdef ECH_DBG_NO_ACCUM
        if (false) {
#else
        if (active) {
#endif
            if constexpr (WITH_RES) {
                // residual rows (all equations, test node a): every record of my cell contributes
                for (int s = 0; s < G; ++s) {
                    if (round * G + s >= NTASK) break;
                    double phia, gphia[D];
                    test_basis(R, s, a, phia, gphia);
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        double t = R(s, OFF_RES + i) * phia;
#pragma unroll
                        for (int k2 = 0; k2 < D; ++k2) t += R(s, OFF_RES + NF + i * D + k2) * gphia[k2];
                        res[i] += t;
                    }
                }
            }
            if constexpr (FACET_IN_ROUND) {
                if (round < NQC / G) {
#pragma unroll
                    for (int s = 0; s < G; ++s) {
                        double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
                        get_basis(R, s, OFF_PHI, OFF_GPHI, phi, gphi);
                        test_basis(R, s, a, phia, gphia);
                        take_all<MC>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                    }
                } else {
                    const int lf0 = (round * G - NQC) / NQF;
#pragma unroll
                    for (int fi = 0; fi < G / NQF; ++fi) {
                        const int lf = lf0 + fi;
                        if (lf < NLF) {
                            const int nb = meta[lf];
                            if (nb < 0) {
#pragma unroll
                                for (int qf = 0; qf < NQF; ++qf) {
                                    const int s = fi * NQF + qf;
                                    double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
                                    get_basis(R, s, OFF_PHI, OFF_GPHI, phi, gphi);
                                    test_basis(R, s, a, phia, gphia);
                                    take_all<ME>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                                }
                            } else {
                                if constexpr (ECH_FAMILY_DG) {
                                    double accn[NF][NF][NLOC];
                                    zero_blocks<false>(accn);
#pragma unroll
                                    for (int qf = 0; qf < NQF; ++qf) {
                                        const int s = fi * NQF + qf;
                                        double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
                                        test_basis(R, s, a, phia, gphia);
                                        get_basis(R, s, OFF_PHI2, OFF_GPHI2, phi, gphi);
                                        take_all<MIM>(R, s, OFF_JM, phia, gphia, phi, gphi, accn);
                                        get_basis(R, s, OFF_PHI, OFF_GPHI, phi, gphi);
                                        take_all<MIP>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                                    }
                                    store_blocks<false>(A.out, rp, ncol, meta[2 * NLF + lf], accn);
                                }
                            }
                        }
                    }
                }
            } else {
                double(&accn)[NF][NF][NLOC] = reinterpret_cast<double(&)[NF][NF][NLOC]>(accn_persist);
                for (int s = 0; s < G; ++s) {
                    const int tk = round * G + s;
                    if (tk >= NTASK) break;
                    double phi[NLOC], gphi[NLOC][D], phia, gphia[D];
                    get_basis(R, s, OFF_PHI, OFF_GPHI, phi, gphi);
                    test_basis(R, s, a, phia, gphia);
                    if (tk < NQC) {
                        take_all<MC>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                    } else {
                        const int ft = tk - NQC;
                        const int lf = ft / NQF, qf = ft - lf * NQF;
                        const int nb = meta[lf];
                        if (nb < 0) {
                            take_all<ME>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                        } else {
                            if constexpr (ECH_FAMILY_DG) {
                                take_all<MIP>(R, s, OFF_JP, phia, gphia, phi, gphi, acc);
                                get_basis(R, s, OFF_PHI2, OFF_GPHI2, phi, gphi);
                                take_all<MIM>(R, s, OFF_JM, phia, gphia, phi, gphi, accn);
                                if (qf == NQF - 1) {
                                    store_blocks<false>(A.out, rp, ncol, meta[2 * NLF + lf], accn);
                                    zero_blocks<false>(accn);
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (active) {
        store_blocks<true>(A.out, rp, ncol, pos_self, acc);
        if constexpr (WITH_RES) {
#pragma unroll
            for (int i = 0; i < NF; ++i) A.out2[i * rows_per_field + c * NLOC + a] = res[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// residual: records carry the test-side basis and w * (f0, f1) of every equation
constexpr int RO_F0 = NLOC * (1 + D);
constexpr int RO_F1 = RO_F0 + NF;

template <class M, int LEN>
__device__ __forceinline__ void put_res(double (&buf)[LEN], const Res &o, double w) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr bool h0 = M::F0[i], h1 = M::F1[i];
        if constexpr (h0)
            buf[RO_F0 + i] = w * o.f0[i];
        else
            buf[RO_F0 + i] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if constexpr (h1)
                buf[RO_F1 + i * D + k] = w * o.f1[i][k];
            else
                buf[RO_F1 + i * D + k] = 0.0;
        }
    });
}

__global__ void __launch_bounds__(RWARPS * 32) residual_coop_kernel(const ech_dg_args A) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, a = lane - g * G;       // cell-major lane numbering
    const bool active_lane = g < CPW;
    const int64_t c = A.cell_begin + ((int64_t)blockIdx.x * RWARPS + warp) * CPW + g;
    const bool active = active_lane && c < A.ncell;
    const int64_t fstride = A.ncell_total * NLOC;
    const WarpSmem W = warp_smem(smem, warp, RREC_LEN);
    const RecRef R{W.rec + 2 * (active_lane ? g : 0)};
    const int ge = lane % CPW, ae = lane / CPW;     // producer role (see RecRef)
    const int64_t ce = A.cell_begin + ((int64_t)blockIdx.x * RWARPS + warp) * CPW + ge;
    const bool active_e = ae < G && ce < A.ncell;
    const RecRef RE{W.rec + 2 * ge};
    const int *meta_e = W.meta + ge * META_LEN;
    const double *own = W.nbd + (ge * NBD_SLOTS + NLF) * NBD_LEN;
    stage_neighbours(A, W, g, a, c, active, fstride);
    double r[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) r[i] = 0.0;

    for (int round = 0; round < NROUND; ++round) {
        const int task = round * G + ae;
        if (active_e && task < NTASK) {
            Pt p;
            p.K = A.consts;
            Geo gm;
            double phi[NLOC], gphi[NLOC][D];
            int idx[D];
            double w;
            Res o;
            double buf[RREC_LEN];
#pragma unroll
            for (int m = 0; m < RREC_LEN; ++m) buf[m] = 0.0;
            if (task < NQC) {
                cell_point(task, idx, w);
                own_side(own, idx, p, gm, phi, gphi);
                cell_res(p, o);
                buf_basis(buf, OFF_PHI, OFF_GPHI, phi, gphi);
                put_res<MC>(buf, o, w * fabs(gm.detJ));
            } else {
                const int ft = task - NQC;
                const int lf = ft / NQF, qf = ft - lf * NQF;
                const int nb = meta_e[lf];
                p.fa = A.facet_area[ce * NLF + lf];
                int tt[D > 1 ? D - 1 : 1];
                double ds;
                facet_point(lf, qf, idx, tt, w);
                own_side(own, idx, p, gm, phi, gphi);
                facet_normal(gm, lf, p.n, ds);
                buf_basis(buf, OFF_PHI, OFF_GPHI, phi, gphi);
                if (nb < 0) {
                    efacet_res(-nb - 1, p, o);
                    put_res<ME>(buf, o, w * ds);
                } else {
                    if constexpr (ECH_FAMILY_DG) {
                        double phi2[NLOC], gphi2[NLOC][D];
                        neighbour_side(W.nbd + (ge * NBD_SLOTS + lf) * NBD_LEN, meta_e[NLF + lf], tt, p, phi2, gphi2);
                        ifacet_res(p, o);
                        put_res<MIP>(buf, o, w * ds);
                    }
                }
            }
            flush_record(RE, ae, buf, 0, RREC_LEN);
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
        const int64_t ostride = A.ncell_total * NLOC;
#pragma unroll
        for (int i = 0; i < NF; ++i) A.out[i * ostride + c * NLOC + a] = r[i];
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int JSMEM = JWARPS * warp_smem_doubles(JREC_LEN) * 8;
constexpr int JFSMEM = JWARPS * warp_smem_doubles(JFREC_LEN) * 8;
constexpr int RSMEM = RWARPS * warp_smem_doubles(RREC_LEN) * 8;

static int coop_setup() {
    static int done = 0;
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(jacobian_coop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, JFSMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(jacobian_coop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, JSMEM);
    if (e != cudaSuccess) {
        fprintf(stderr, "echem_b200: cudaFuncSetAttribute(jacobian_coop_kernel, %d B) failed: %s\n", JSMEM,
                cudaGetErrorString(e));
        return (int)e;
    }
    e = cudaFuncSetAttribute(residual_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM);
    if (e != cudaSuccess) {
        fprintf(stderr, "echem_b200: cudaFuncSetAttribute(residual_coop_kernel, %d B) failed: %s\n", RSMEM,
                cudaGetErrorString(e));
        return (int)e;
    }
    done = 1;
    return 0;
}

static int launch_jacobian_coop(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    int e = coop_setup();
    if (e) return e;
    const int64_t cells_per_block = (int64_t)JWARPS * CPW;
    if (a->cell_begin < 0) return (int)cudaErrorInvalidValue;
    if (a->cell_begin >= a->ncell) return 0;
    const int64_t grid = (a->ncell - a->cell_begin + cells_per_block - 1) / cells_per_block;   // cells [cell_begin, ncell)
    if (a->out2)
        jacobian_coop_kernel<true><<<(unsigned)grid, JWARPS * 32, JFSMEM, s>>>(*a);
    else
        jacobian_coop_kernel<false><<<(unsigned)grid, JWARPS * 32, JSMEM, s>>>(*a);
    e = (int)cudaGetLastError();
    if (e) fprintf(stderr, "echem_b200: jacobian_coop_kernel<<<%lld, %d, %d>>> failed: %s\n", (long long)grid,
                   JWARPS * 32, JSMEM, cudaGetErrorString((cudaError_t)e));
    return e;
}

static int launch_residual_coop(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    int e = coop_setup();
    if (e) return e;
    const int64_t cells_per_block = (int64_t)RWARPS * CPW;
    if (a->cell_begin < 0) return (int)cudaErrorInvalidValue;
    if (a->cell_begin >= a->ncell) return 0;
    const int64_t grid = (a->ncell - a->cell_begin + cells_per_block - 1) / cells_per_block;   // cells [cell_begin, ncell)
    residual_coop_kernel<<<(unsigned)grid, RWARPS * 32, RSMEM, s>>>(*a);
    e = (int)cudaGetLastError();
    if (e) fprintf(stderr, "echem_b200: residual_coop_kernel<<<%lld, %d, %d>>> failed: %s\n", (long long)grid,
                   RWARPS * 32, RSMEM, cudaGetErrorString((cudaError_t)e));
    return e;
}

}  // namespace echdg
