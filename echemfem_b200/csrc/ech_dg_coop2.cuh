// Cooperative DG Jacobian (+ fused residual) kernel, second generation: REFERENCE-SPACE records.
//
// ncu on the first cooperative kernel (ech_dg_coop.cuh, profiles/r1f_*) shows the LSU data pipe as
// the limiter: one wavefront per 128 B of shared-memory record traffic plus one per 32-byte sector
// of the CSR stores.  A third of every record was the physical trial/test bases of the point.  Here
// the producer folds the inverse Jacobians of the test side (always the own cell) and of the trial
// side (own cell or neighbour) into the pointwise coefficients,
//
//     g1^[n]    = w * sum_l  Jt^-1[n][l] g1[l]            (trial gradient)
//     g2^[m]    = w * sum_k  Jo^-1[m][k] g2[k]            (test gradient)
//     g3^[m][n] = w * sum_kl Jo^-1[m][k] g3[k][l] Jt^-1[n][l]
//
// so that the consumer contracts them with REFERENCE-element bases.  Those are compile-time
// constants for every own-side point (cell points and own-facet points are compile-time task
// numbers once the slot loop is unrolled): no basis values travel through shared memory, and
// structurally zero basis values (Gauss-Legendre DQ nodes are collocated with the cell rule;
// basis functions off a facet vanish on it) drop out of the accumulation at compile time.  Only
// the neighbour-side trial basis of an interior-facet point (its local facet and orientation are
// run-time mesh data) is still carried by the record, in reference space (table products, no
// geometry).  The test-side values of a lane come from 1D factors it keeps in registers.
//
// Same ownership / store discipline as ech_dg_coop.cuh (every CSR value written exactly once).
#pragma once
#include "ech_dg_coop.cuh"

namespace echdg {
namespace v2 {

// ---------------------------------------------------------------- compile-time task geometry
struct TaskPt {
    int kind;        // 0 cell point, 1 facet point
    int lf, qf;
    int idx[D];      // 1D table index per direction
};

__host__ __device__ constexpr TaskPt task_point(int tk) {
    TaskPt t{};
    if (tk < NQC) {
        t.kind = 0;
        t.lf = -1;
        t.qf = tk;
        for (int k = 0; k < D; ++k) t.idx[k] = (tk / ipow(NQ, D - 1 - k)) % NQ;
    } else {
        const int ft = tk - NQC;
        t.kind = 1;
        t.lf = ft / NQF;
        t.qf = ft % NQF;
        const int kf = t.lf >> 1, s = t.lf & 1;
        int tt[2] = {0, 0};
        for (int m = 0; m < D - 1; ++m) tt[m] = (t.qf / ipow(NQ, D - 2 - m)) % NQ;
        for (int k = 0; k < D; ++k) {
            const int m = (k < kf) ? k : k - 1;
            t.idx[k] = (k == kf) ? NQ + s : (m == 0 ? tt[0] : tt[1]);
        }
    }
    return t;
}

__host__ __device__ constexpr double ref_phi(const TaskPt &t, int b) {
    double v = 1.0;
    for (int k = 0; k < D; ++k) v *= CT_L[t.idx[k]][digit(b, k)];
    return v;
}
__host__ __device__ constexpr double ref_dphi(const TaskPt &t, int b, int m) {
    double v = 1.0;
    for (int k = 0; k < D; ++k) v *= (k == m) ? CT_DL[t.idx[k]][digit(b, k)] : CT_L[t.idx[k]][digit(b, k)];
    return v;
}

// ---------------------------------------------------------------- record layout (doubles)
template <class M>
struct Layout2 {
    int o0[NF][NF], o1[NF][NF], o2[NF][NF], o3[NF][NF];
    int rowbeg[NF + 1];     // coefficients of equation i occupy [rowbeg[i], rowbeg[i+1]), even bounds (flushed per equation)
    int klen;               // [0, klen): K = Jo^-1 Jt^-T, shared by every isotropic (scalar g3) coupling of the record
    int len;
    __host__ __device__ constexpr Layout2() : o0{}, o1{}, o2{}, o3{}, rowbeg{}, klen(0), len(0) {
        bool iso = false;
        for (int i = 0; i < NF; ++i)
            for (int j = 0; j < NF; ++j) iso = iso || M::G3[i][j] == 1;
        klen = iso ? ((D * D + 1) & ~1) : 0;
        int n = klen;
        for (int i = 0; i < NF; ++i) {
            n = (n + 1) & ~1;
            rowbeg[i] = n;
            for (int j = 0; j < NF; ++j) {
                o0[i][j] = o1[i][j] = o2[i][j] = o3[i][j] = -1;
                if (M::G1[i][j]) { n = (D == 2) ? ((n + 1) & ~1) : n; o1[i][j] = n; n += D; }
                if (M::G2[i][j]) { n = (D == 2) ? ((n + 1) & ~1) : n; o2[i][j] = n; n += D; }
                if (M::G3[i][j] == 2) { n = (D == 2) ? ((n + 1) & ~1) : n; o3[i][j] = n; n += D * D; }
                if (M::G0[i][j]) { o0[i][j] = n; n += 1; }
                if (M::G3[i][j] == 1) { o3[i][j] = n; n += 1; }
            }
        }
        n = (n + 1) & ~1;
        rowbeg[NF] = n;
        len = n;
    }
};
constexpr Layout2<MC> L2_C{};
constexpr Layout2<ME> L2_E{};
constexpr Layout2<MIP> L2_IP{};
constexpr Layout2<MIM> L2_IM{};
template <class M> struct L2Of;
template <> struct L2Of<MC> { static __host__ __device__ constexpr const Layout2<MC> &get() { return L2_C; } };
template <> struct L2Of<ME> { static __host__ __device__ constexpr const Layout2<ME> &get() { return L2_E; } };
template <> struct L2Of<MIP> { static __host__ __device__ constexpr const Layout2<MIP> &get() { return L2_IP; } };
template <> struct L2Of<MIM> { static __host__ __device__ constexpr const Layout2<MIM> &get() { return L2_IM; } };

constexpr int even(int n) { return (n + 1) & ~1; }

template <bool WITH_RES>
struct Rec {
    static constexpr int OFF_RES = 0;
    static constexpr int RLEN = WITH_RES ? even(NF * (1 + D)) : 0;
    static constexpr int OFF_JP = RLEN;
    static constexpr int END_C = OFF_JP + even(L2_C.len);
    static constexpr int END_E = OFF_JP + even(L2_E.len);
    static constexpr int OFF_JM = OFF_JP + even(L2_IP.len);
    static constexpr int END_I = OFF_JM + even(L2_IM.len);
    static constexpr int LEN = cmax(cmax(END_C, END_E), ECH_FAMILY_DG ? END_I : 2);
};


// rounds in which the points of facet lf are consumed
constexpr int facet_first_round(int lf) { return (NQC + lf * NQF) / G; }
constexpr int facet_last_round(int lf) { return (NQC + lf * NQF + NQF - 1) / G; }
constexpr bool any_straddle() {
    for (int lf = 0; lf < NLF; ++lf)
        if (facet_first_round(lf) != facet_last_round(lf)) return true;
    return false;
}
constexpr bool STRADDLE = any_straddle();

// Large elements (3D, many fields): NF*NF*NLOC own-block accumulators per lane do not fit the register file
// next to the point evaluation (ptxas: 7 KB of spills per thread for three fields on hexes).  They then live in
// shared memory, [equation][field][node][lane] (conflict-free), and the consumer walks the records of a round
// once per equation with only that equation's NF*NLOC accumulators in registers.
// Measured on configs[3] (hexes, 3 fields, 12.6 M dofs): spills drop from 7 KB to 0.6 KB per thread, but the fused
// step is SLOWER (27.7 ms against 24.4 ms): the per-equation passes recompute the test / neighbour bases and the
// extra 18 KB per warp cost occupancy.  Opt-in (-DECH_SMEM_ACC) until the producer side of 3D is leaner.
#ifdef ECH_SMEM_ACC
constexpr bool SMEM_ACC = !STRADDLE && (NF * NF * NLOC > 48);
#else
constexpr bool SMEM_ACC = false;
#endif
constexpr int ACC_DOUBLES = SMEM_ACC ? NF * NF * NLOC * 32 : 0;
// neighbour blocks one equation at a time (see the consumer).  Measured on the 3D three-ion case (12.6 M dofs):
// stack 7 KB -> 1.5-1.9 KB per thread, but 27.3 ms against 24.3 ms (the recomputed neighbour basis costs more than the
// spills it removes), so it is off by default
#ifndef ECH_NBR_PER_EQ
#define ECH_NBR_PER_EQ 0
#endif
constexpr bool NBR_PER_EQ = ECH_NBR_PER_EQ != 0 && !SMEM_ACC && !STRADDLE && (NF * NF * NLOC > 48);


// ---------------------------------------------------------------- staged cell data (asynchronous copies)
// Per warp: [records | cell data | facet areas | runtime constants | facet metadata].
// Cell data of (cell g, slot s) -- slot lf = neighbour across facet lf, slot NLF = the cell itself:
//   X[NV*D] U[NF*NLOC] A[NAUX*NLOC] CV[2]   padded to an even length (16-byte units for cp.async / LDS.128),
// cell stride = 10 mod 16 doubles, so that the 8 cells a quarter-warp reads at equal offsets fall into
// distinct 16-byte bank groups.
constexpr bool V16 = (NLOC % 2 == 0);          // 16-byte copies and 128-bit shared loads are possible
constexpr int S_X = 0;
constexpr int S_U = NV * D;
constexpr int S_A = S_U + NF * NLOC;
constexpr int S_CV = S_A + NAUX * NLOC;
constexpr int SLOT_LEN = even(S_CV + 2);
constexpr int cell_stride() {
    const int base = NBD_SLOTS * SLOT_LEN;
    return base + ((10 - base % 16) % 16 + 16) % 16;
}
constexpr int CELL_STRIDE = cell_stride();
constexpr int NCONST_PAD = even(ECH_NCONST);

// ECH_COOP2_PIPE (default 1): staging with the PRODUCER lane numbering (lane = slot * CPW + cell): every lane loads
// ONE neighbour id and copies that neighbour's data; the eight lanes of a quarter-warp write eight different cells
// at equal offsets, which the cell stride (10 mod 16 doubles) spreads over distinct bank groups -- the consumer
// numbering of round 1 wrote 4.3 x the ideal number of shared-memory wavefronts (LDGSTS, ncu source page).
// ECH_COOP2_NBUF = 2 double-buffers the staged cell data, so that the data of batch i+1 arrives while batch i is
// evaluated and the ids of batch i+2 are loaded (both dependent round trips of the staging off the critical path).
// Measured on cfg2 (profiles/r2i_coop2_staging_variants.txt): the extra 7 KB per warp cost the eighth resident
// warp per SM (2.93 ms against 2.55 ms); with 2 warps per CTA to amortise the per-CTA reserve it is 3.2 ms.  The
// kernel is occupancy-bound (time ~ 1 / resident warps), so the default is ONE buffer, the next batch staged
// during the last round, persistent warps over 2 waves: 2.51 ms against 2.61 ms for the round-1 staging.
#ifndef ECH_COOP2_PIPE
#define ECH_COOP2_PIPE 1
#endif
#ifndef ECH_COOP2_NBUF
#define ECH_COOP2_NBUF 1
#endif
// (only where the lanes of a warp divide evenly among the cells: for Q2, G = 9, the remapped staging measured
// 12 % slower than the round-1 one -- profiles/r2q_bench_cfg2_Q2.json -- and is not used)
constexpr bool PIPE = ECH_COOP2_PIPE != 0 && G >= NLF && (32 % G == 0);
constexpr int NBUF = (PIPE && ECH_COOP2_NBUF == 2) ? 2 : 1;       // 1: the next batch is staged during the last round
constexpr bool EARLY = NBUF == 2;

struct Smem2 {
    double *rec;       // RECLEN * 32
    double *nbd;       // [NBUF][CPW][CELL_STRIDE]
    double *fa;        // [NBUF][even(CPW * NLF)] facet areas of my cells
    double *kc;        // [NCONST_PAD] runtime constants
    double *acc;       // [NF][NF][NLOC][32] own-block accumulators (SMEM_ACC only)
    int *meta;         // [2][CPW][META_LEN]  (double-buffered: the next batch is staged while this one is consumed)
};
constexpr int FA_LEN = even(CPW * NLF);
__host__ __device__ constexpr int warp_smem2_doubles(int reclen) {
    return even(reclen * 32 + NBUF * (CPW * CELL_STRIDE + FA_LEN) + NCONST_PAD + ACC_DOUBLES + CPW * META_LEN);
}
__device__ __forceinline__ Smem2 warp_smem2(double *smem, int warp, int reclen) {
    double *base = smem + (size_t)warp * warp_smem2_doubles(reclen);
    Smem2 w;
    w.rec = base;
    w.nbd = base + reclen * 32;
    w.fa = w.nbd + NBUF * CPW * CELL_STRIDE;
    w.kc = w.fa + NBUF * FA_LEN;
    w.acc = w.kc + NCONST_PAD;
    w.meta = reinterpret_cast<int *>(w.acc + ACC_DOUBLES);
    return w;
}
constexpr int coop2_warps(int reclen) {
#ifdef ECH_COOP_WARPS
    return ECH_COOP_WARPS + 0 * reclen;
#else
    // one warp per CTA: warps never synchronise with each other, and single-warp CTAs retire and start
    // independently (measured on cfg2: 2.64 ms against 2.72-2.76 ms for 2 or 4 warps per CTA)
    return 1 + 0 * reclen;
#endif
}

__device__ __forceinline__ void cp_async16(double *dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(double *dst, const double *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
template <int N>
__device__ __forceinline__ void cp_run(double *dst, const double *src, int k, int units) {
    // unit k of a run of N doubles: 16-byte units when the run is pair-aligned, 8-byte units otherwise
    (void)units;
    if constexpr (V16 && N % 2 == 0)
        cp_async16(dst + 2 * k, src + 2 * k);
    else
        cp_async8(dst + k, src + k);
}
template <int N>
constexpr int run_units() { return (V16 && N % 2 == 0) ? N / 2 : N; }

// copy the data of cell nb into one slot; with SPLIT, only the units k with k % NPART == part
template <int SPLIT, int NPART>
__device__ __forceinline__ void copy_slot(const ech_dg_args &A, double *d, int nb, int64_t fstride, int part) {
    constexpr int UX = run_units<NV * D>(), UU = run_units<NLOC>();
    const double *px = A.xcell + (int64_t)nb * (NV * D);
    int k = 0;
    static_for<0, UX>([&](auto uc) {
        constexpr int u = decltype(uc)::value;
        if (!SPLIT || (k % NPART) == part) cp_run<NV * D>(d + S_X, px, u, UX);
        ++k;
    });
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const double *pu = A.u + j * fstride + (int64_t)nb * NLOC;
        static_for<0, UU>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            if (!SPLIT || (k % NPART) == part) cp_run<NLOC>(d + S_U + j * NLOC, pu, u, UU);
            ++k;
        });
    });
    static_for<0, NAUX>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const double *pa = A.aux + j * fstride + (int64_t)nb * NLOC;
        static_for<0, UU>([&](auto uc) {
            constexpr int u = decltype(uc)::value;
            if (!SPLIT || (k % NPART) == part) cp_run<NLOC>(d + S_A + j * NLOC, pa, u, UU);
            ++k;
        });
    });
    if (!SPLIT || (k % NPART) == part) cp_async8(d + S_CV, A.cell_volume + nb);
    ++k;
    if (!SPLIT || (k % NPART) == part) cp_async8(d + S_CV + 1, A.cell_diam + nb);
}

// Staging of one batch (my cell + its facet neighbours), split in two so that the two dependent round trips
// of the NEXT batch overlap the work on the current one:
//   load_ids   : neighbour ids of my cell (all lanes) + facet code / column slot (lane lf % G) -> registers
//   issue_stage: facet metadata -> shared memory (buffer mb), then all copies as cp.async (no registers, no
//                shared-memory store instructions); lane a copies neighbour slot lf = a (+ G ...) whole, the
//                self slot is split over the lanes.
// loads that must be ISSUED where they are written (the compiler otherwise sinks them to their first use, which
// turns the latency they were meant to overlap into a stall there)
__device__ __forceinline__ int ldg_s32(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_u8(const uint8_t *p) {
    int v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t ldg_s64(const int64_t *p) {
    int64_t v;
    asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void load_ids(const ech_dg_args &A, int a, int64_t c, bool active, int (&ids)[NLF],
                                         int (&cs)[2 * NLF]) {
#pragma unroll
    for (int lf = 0; lf < NLF; ++lf) {
        ids[lf] = -1;
        cs[lf] = cs[NLF + lf] = 0;
    }
    if (active) {
#pragma unroll
        for (int lf = 0; lf < NLF; ++lf) ids[lf] = ldg_s32(A.nbr + c * NLF + lf);
#pragma unroll
        for (int lf = 0; lf < NLF; ++lf)
            if (lf % G == a) {
                cs[lf] = ldg_u8(A.code + c * NLF + lf);
                cs[NLF + lf] = ldg_u8(A.slot + c * NLF + lf);
            }
    }
}

__device__ __forceinline__ void issue_stage(const ech_dg_args &A, const Smem2 &W, int mb, int g, int a, int64_t c,
                                            bool active, const int (&ids)[NLF], const int (&cs)[2 * NLF],
                                            int64_t fstride) {
    if (active) {
        int *m = W.meta + (mb * CPW + g) * META_LEN;
#pragma unroll
        for (int lf = 0; lf < NLF; ++lf) {
            if (lf % G == a) {
                m[lf] = ids[lf];
                m[NLF + lf] = cs[lf];
                m[2 * NLF + lf] = cs[NLF + lf];
                cp_async8(W.fa + g * NLF + lf, A.facet_area + c * NLF + lf);
            }
        }
        for (int lf = a; lf < NLF; lf += G) {
            int nb = ids[0];
#pragma unroll
            for (int t = 1; t < NLF; ++t) nb = (lf == t) ? ids[t] : nb;
            if (nb >= 0) copy_slot<0, 1>(A, W.nbd + g * CELL_STRIDE + lf * SLOT_LEN, nb, fstride, 0);
        }
        copy_slot<1, G>(A, W.nbd + g * CELL_STRIDE + NLF * SLOT_LEN, (int)c, fstride, a);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- staging with the producer lane numbering (PIPE): lane (ge, ae) = (cell of the warp, slot) copies the data of the
// neighbour across facet ae of cell ge whole, and its share (ae of G) of the cell's own data
struct StageIds {
    int nb, code, slot;     // neighbour across facet ae of cell ge (nb = -1: none / boundary class folded below)
};
__device__ __forceinline__ StageIds load_ids_e(const ech_dg_args &A, int ae, int64_t ce, bool active_e) {
    StageIds r{-1, 0, 0};
    if (active_e && ae < NLF) {
        r.nb = ldg_s32(A.nbr + ce * NLF + ae);
        r.code = ldg_u8(A.code + ce * NLF + ae);
        r.slot = ldg_u8(A.slot + ce * NLF + ae);
    }
    return r;
}
__device__ __forceinline__ void issue_stage_e(const ech_dg_args &A, const Smem2 &W, int mb, int ge, int ae, int64_t ce,
                                              bool active_e, const StageIds &id, int64_t fstride) {
    if (active_e) {
        double *nbd = W.nbd + (EARLY ? mb : 0) * (CPW * CELL_STRIDE) + ge * CELL_STRIDE;
        if (ae < NLF) {
            int *m = W.meta + (mb * CPW + ge) * META_LEN;
            m[ae] = id.nb;
            m[NLF + ae] = id.code;
            m[2 * NLF + ae] = id.slot;
            cp_async8(W.fa + (EARLY ? mb : 0) * FA_LEN + ge * NLF + ae, A.facet_area + ce * NLF + ae);
            if (id.nb >= 0) copy_slot<0, 1>(A, nbd + ae * SLOT_LEN, id.nb, fstride, 0);
        }
        copy_slot<1, G>(A, nbd + NLF * SLOT_LEN, (int)ce, fstride, ae);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// N consecutive doubles from shared memory (128-bit loads when the run is pair-aligned)
template <int N>
__device__ __forceinline__ void lds_run(const double *d, double *out) {
    constexpr bool wide = V16 && N % 2 == 0;
    if constexpr (wide) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(d + i);
            out[i] = v.x;
            out[i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = d[i];
    }
}

constexpr int J2WARPS = coop2_warps(Rec<true>::LEN);

// ---------------------------------------------------------------- producer helpers
__device__ __forceinline__ void basis_ref(const int (&idx)[D], double (&phi)[NLOC], double (&dref)[NLOC][D]) {
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
        double v = 1.0, dr[D];
#pragma unroll
        for (int m = 0; m < D; ++m) dr[m] = 1.0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const int bk = digit(b, k);
            v *= L[k][bk];
#pragma unroll
            for (int m = 0; m < D; ++m) dr[m] *= (m == k) ? dL[k][bk] : L[k][bk];
        }
        phi[b] = v;
#pragma unroll
        for (int m = 0; m < D; ++m) dref[b][m] = dr[m];
    }
}

// value and PHYSICAL gradient of NFLD fields from the reference basis: grad[a] = sum_m invJ[m][a] * (sum_b U_b dref_b[m])
template <int NFLD>
__device__ __forceinline__ void interp_ref(const double *d, const double (&phi)[NLOC], const double (&dref)[NLOC][D],
                                           const double (&invJ)[D][D], double (&u)[NFLD], double (&gu)[NFLD][D]) {
#pragma unroll
    for (int j = 0; j < NFLD; ++j) {
        double s = 0.0, gr[D];
#pragma unroll
        for (int m = 0; m < D; ++m) gr[m] = 0.0;
        double U[NLOC];
        lds_run<NLOC>(d + j * NLOC, U);
#pragma unroll
        for (int b = 0; b < NLOC; ++b) {
            const double ub = U[b];
            s += ub * phi[b];
#pragma unroll
            for (int m = 0; m < D; ++m) gr[m] += ub * dref[b][m];
        }
        u[j] = s;
#pragma unroll
        for (int a = 0; a < D; ++a) {
            double g = 0.0;
#pragma unroll
            for (int m = 0; m < D; ++m) g += invJ[m][a] * gr[m];
            gu[j][a] = g;
        }
    }
}

__device__ __forceinline__ void read_coords(const double *d, double (&X)[NV][D]) {
    lds_run<NV * D>(d + S_X, &X[0][0]);
}

// own side of a point from the staged cell data
__device__ __forceinline__ void own_side2(const double *d, const int (&idx)[D], Pt &p, Geo &gm) {
    double X[NV][D];
    read_coords(d, X);
    {
        double cvd[2];
        lds_run<2>(d + S_CV, cvd);
        p.cv = cvd[0];
        p.cd = cvd[1];
    }
    geometry(X, idx, gm);
    double phi[NLOC], dref[NLOC][D];
    basis_ref(idx, phi, dref);
    interp_ref<NF>(d + S_U, phi, dref, gm.invJ, p.u, p.gu);
    if constexpr (NAUX > 0) interp_ref<ECH_NA1>(d + S_A, phi, dref, gm.invJ, p.a, p.ga);
#pragma unroll
    for (int k = 0; k < D; ++k) p.x[k] = gm.x[k];
}

// neighbour side of a facet point: state for the physics, inverse Jacobian for the folding
__device__ __forceinline__ void neighbour_side2(const double *d, int code, const int (&tt)[D > 1 ? D - 1 : 1], Pt &p,
                                                double (&invJn)[D][D]) {
    double X2[NV][D];
    read_coords(d, X2);
    {
        double cvd[2];
        lds_run<2>(d + S_CV, cvd);
        p.cvm = cvd[0];
        p.cdm = cvd[1];
    }
    int idx2[D];
    neighbour_point(code, tt, idx2);
    Geo g2;
    geometry(X2, idx2, g2);
    double phi2[NLOC], dref2[NLOC][D];
    basis_ref(idx2, phi2, dref2);
    interp_ref<NF>(d + S_U, phi2, dref2, g2.invJ, p.um, p.gum);
    if constexpr (NAUX > 0) interp_ref<ECH_NA1>(d + S_A, phi2, dref2, g2.invJ, p.am, p.gam);
#pragma unroll
    for (int m = 0; m < D; ++m)
#pragma unroll
        for (int a = 0; a < D; ++a) invJn[m][a] = g2.invJ[m][a];
}

// K[m][n] = sum_k Jo[m][k] Jt[n][k], stored once per record when the mask has isotropic couplings
template <class M, int LEN>
__device__ __forceinline__ void put_k(double (&buf)[LEN], int off, const double (&Jo)[D][D], const double (&Jt)[D][D]) {
    if constexpr (L2Of<M>::get().klen > 0) {
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int n = 0; n < D; ++n) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += Jo[m][k] * Jt[n][k];
                buf[off + m * D + n] = s;
            }
    }
}

// coefficients of equation I -> reference space, weight folded in.  Jo: test side, Jt: trial side.
template <class M, int I, int LEN>
__device__ __forceinline__ void put_rowj2(double (&buf)[LEN], int off, const RowJ &r, double w, const double (&Jo)[D][D],
                                          const double (&Jt)[D][D]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int p0 = L2Of<M>::get().o0[I][j], p1 = L2Of<M>::get().o1[I][j];
        constexpr int p2 = L2Of<M>::get().o2[I][j], p3 = L2Of<M>::get().o3[I][j];
        constexpr int h3 = M::G3[I][j];
        if constexpr (p0 >= 0) buf[off + p0] = w * r.g0[j];
        if constexpr (p1 >= 0) {
#pragma unroll
            for (int n = 0; n < D; ++n) {
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < D; ++l) s += Jt[n][l] * r.g1[j][l];
                buf[off + p1 + n] = w * s;
            }
        }
        if constexpr (p2 >= 0) {
#pragma unroll
            for (int m = 0; m < D; ++m) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += Jo[m][k] * r.g2[j][k];
                buf[off + p2 + m] = w * s;
            }
        }
        if constexpr (h3 == 1) buf[off + p3] = w * r.g3[j][0][0];
        if constexpr (h3 == 2) {
            double t[D][D];    // t[k][n] = sum_l g3[k][l] Jt[n][l]
#pragma unroll
            for (int k = 0; k < D; ++k)
#pragma unroll
                for (int n = 0; n < D; ++n) {
                    double s = 0.0;
#pragma unroll
                    for (int l = 0; l < D; ++l) s += r.g3[j][k][l] * Jt[n][l];
                    t[k][n] = s;
                }
#pragma unroll
            for (int m = 0; m < D; ++m)
#pragma unroll
                for (int n = 0; n < D; ++n) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < D; ++k) s += Jo[m][k] * t[k][n];
                    buf[off + p3 + m * D + n] = w * s;
                }
        }
    });
}

template <class M, int LEN>
__device__ __forceinline__ void put_res2(double (&buf)[LEN], const Res &o, double w, const double (&Jo)[D][D]) {
    static_for<0, NF>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        if constexpr (M::F0[i]) buf[i] = w * o.f0[i];
        if constexpr (M::F1[i]) {
#pragma unroll
            for (int m = 0; m < D; ++m) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) s += Jo[m][k] * o.f1[i][k];
                buf[NF + i * D + m] = w * s;
            }
        }
    });
}

// ---------------------------------------------------------------- consumer helpers
// rows (I, a) += one record.  TK >= 0: trial basis = compile-time reference basis of own-side task TK;
// TK < 0: trial basis (phi2, d2) read from the record (neighbour side).
template <class M, int I, int TK>
__device__ __forceinline__ void take_rowj2(const RecRef &R, int slot, int off, double pa, const double (&da)[D],
                                           const double (&Kd)[D], const double (&phi2)[NLOC],
                                           const double (&d2)[NLOC][D], double (&acc)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int p0 = L2Of<M>::get().o0[I][j], p1 = L2Of<M>::get().o1[I][j];
        constexpr int p2 = L2Of<M>::get().o2[I][j], p3 = L2Of<M>::get().o3[I][j];
        constexpr bool h0 = p0 >= 0, h1 = p1 >= 0, h2 = p2 >= 0, h3 = p3 >= 0;
        constexpr bool iso = M::G3[I][j] == 1;
        if constexpr (h0 || h1 || h2 || h3) {
            double Aj = 0.0, Bj[D];
#pragma unroll
            for (int n = 0; n < D; ++n) Bj[n] = 0.0;
            if constexpr (h0) Aj += R(slot, off + p0) * pa;
            if constexpr (h2) {
                if constexpr (D == 2) {
                    const double2 v = R.pair(slot, off + p2);
                    Aj += v.x * da[0];
                    Aj += v.y * da[1];
                } else {
#pragma unroll
                    for (int m = 0; m < D; ++m) Aj += R(slot, off + p2 + m) * da[m];
                }
            }
            if constexpr (h1) {
                if constexpr (D == 2) {
                    const double2 v = R.pair(slot, off + p1);
                    Bj[0] += v.x * pa;
                    Bj[1] += v.y * pa;
                } else {
#pragma unroll
                    for (int n = 0; n < D; ++n) Bj[n] += R(slot, off + p1 + n) * pa;
                }
            }
            if constexpr (iso) {
                const double sc = R(slot, off + p3);
#pragma unroll
                for (int n = 0; n < D; ++n) Bj[n] += sc * Kd[n];
            }
            if constexpr (h3 && !iso) {
                if constexpr (D == 2) {
                    const double2 r0 = R.pair(slot, off + p3), r1 = R.pair(slot, off + p3 + 2);
                    Bj[0] += r0.x * da[0];
                    Bj[1] += r0.y * da[0];
                    Bj[0] += r1.x * da[1];
                    Bj[1] += r1.y * da[1];
                } else {
#pragma unroll
                    for (int m = 0; m < D; ++m)
#pragma unroll
                        for (int n = 0; n < D; ++n) Bj[n] += R(slot, off + p3 + m * D + n) * da[m];
                }
            }
            if constexpr (TK >= 0) {
                static_for<0, NLOC>([&](auto bc) {
                    constexpr int b = decltype(bc)::value;
                    constexpr double pb = ref_phi(task_point(TK >= 0 ? TK : 0), b);
                    if constexpr ((h0 || h2) && pb != 0.0) acc[j][b] += Aj * pb;
                    if constexpr (h1 || h3) {
                        static_for<0, D>([&](auto nc) {
                            constexpr int n = decltype(nc)::value;
                            constexpr double db = ref_dphi(task_point(TK >= 0 ? TK : 0), b, n);
                            if constexpr (db != 0.0) acc[j][b] += Bj[n] * db;
                        });
                    }
                });
            } else {
#pragma unroll
                for (int b = 0; b < NLOC; ++b) {
                    double sacc = acc[j][b];
                    if constexpr (h0 || h2) sacc += Aj * phi2[b];
                    if constexpr (h1 || h3) {
#pragma unroll
                        for (int n = 0; n < D; ++n) sacc += Bj[n] * d2[b][n];
                    }
                    acc[j][b] = sacc;
                }
            }
        }
    });
}

template <class M, int TK>
__device__ __forceinline__ void take_all2(const RecRef &R, int s, int off, double pa, const double (&da)[D],
                                          const double (&phi2)[NLOC], const double (&d2)[NLOC][D],
                                          double (&acc)[NF][NF][NLOC]) {
    double Kd[D];      // Kd[n] = sum_m K[m][n] da[m]
#pragma unroll
    for (int n = 0; n < D; ++n) Kd[n] = 0.0;
    if constexpr (L2Of<M>::get().klen > 0) {
        if constexpr (D == 2) {
            const double2 r0 = R.pair(s, off), r1 = R.pair(s, off + 2);
            Kd[0] = r0.x * da[0] + r1.x * da[1];
            Kd[1] = r0.y * da[0] + r1.y * da[1];
        } else {
#pragma unroll
            for (int m = 0; m < D; ++m)
#pragma unroll
                for (int n = 0; n < D; ++n) Kd[n] += R(s, off + m * D + n) * da[m];
        }
    }
    static_for<0, NF>([&](auto ic) {
        constexpr int I = decltype(ic)::value;
        take_rowj2<M, I, TK>(R, s, off, pa, da, Kd, phi2, d2, acc[I]);
    });
}

// one equation of take_all2 (used by the shared-memory-accumulator path, which walks the records once per equation)
template <class M, int I, int TK>
__device__ __forceinline__ void take_row2(const RecRef &R, int s, int off, double pa, const double (&da)[D],
                                          const double (&phi2)[NLOC], const double (&d2)[NLOC][D],
                                          double (&acc)[NF][NLOC]) {
    double Kd[D];
#pragma unroll
    for (int n = 0; n < D; ++n) Kd[n] = 0.0;
    if constexpr (L2Of<M>::get().klen > 0) {
#pragma unroll
        for (int m = 0; m < D; ++m)
#pragma unroll
            for (int n = 0; n < D; ++n) Kd[n] += R(s, off + m * D + n) * da[m];
    }
    take_rowj2<M, I, TK>(R, s, off, pa, da, Kd, phi2, d2, acc);
}

template <bool OWN, int I>
__device__ __forceinline__ void store_row_blocks(double *__restrict__ out, int64_t rpI, int ncol, int pos,
                                                 const double (&acc)[NF][NLOC]) {
    double *__restrict__ vals = out + rpI;
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr bool on = OWN ? OWN_BLOCK[I][j] : NBR_BLOCK[I][j];
        if constexpr (on) store_block(vals + block_offset<I, j>(ncol, pos), acc[j]);
    });
}

template <int I>
__device__ __forceinline__ void acc_load(const double *sacc, int lane, double (&acc)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (OWN_BLOCK[I][j]) {
#pragma unroll
            for (int b = 0; b < NLOC; ++b) acc[j][b] = sacc[((I * NF + j) * NLOC + b) * 32 + lane];
        }
    });
}
template <int I>
__device__ __forceinline__ void acc_store(double *sacc, int lane, const double (&acc)[NF][NLOC]) {
    static_for<0, NF>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (OWN_BLOCK[I][j]) {
#pragma unroll
            for (int b = 0; b < NLOC; ++b) sacc[((I * NF + j) * NLOC + b) * 32 + lane] = acc[j][b];
        }
    });
}

// reference basis of the neighbour at the facet point (lf, qf) of my cell: its local facet and orientation are
// run-time mesh data, the 1D factors come from constant memory (no shared-memory traffic)
template <int TK>
__device__ __forceinline__ void nb_basis(int code, double (&phi2)[NLOC], double (&d2)[NLOC][D]) {
    constexpr int qf = task_point(TK).qf;
    int tt[D > 1 ? D - 1 : 1];
    tt[0] = 0;
    if constexpr (D > 1) {
#pragma unroll
        for (int m = 0; m < D - 1; ++m) tt[m] = (qf / ipow(NQ, D - 2 - m)) % NQ;
    }
    int idx2[D];
    neighbour_point(code, tt, idx2);
    basis_ref(idx2, phi2, d2);
}

// reference values of the test function a of this lane at the compile-time task TK: products of 1D table
// entries selected by the digits of a (selects of immediates: no registers are held across the kernel)
template <int T>
__device__ __forceinline__ double sel_L(int ak) {
    double v = CT_L[T][0];
    static_for<1, N1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        constexpr double c = CT_L[T][m];
        v = (ak == m) ? c : v;
    });
    return v;
}
template <int T>
__device__ __forceinline__ double sel_dL(int ak) {
    double v = CT_DL[T][0];
    static_for<1, N1>([&](auto mc) {
        constexpr int m = decltype(mc)::value;
        constexpr double c = CT_DL[T][m];
        v = (ak == m) ? c : v;
    });
    return v;
}
template <int TK>
__device__ __forceinline__ void test_values(const int (&ad)[D], double &pa, double (&da)[D]) {
    pa = 1.0;
#pragma unroll
    for (int m = 0; m < D; ++m) da[m] = 1.0;
    static_for<0, D>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        constexpr int t = task_point(TK).idx[k];
        const double l = sel_L<t>(ad[k]), dl = sel_dL<t>(ad[k]);
        pa *= l;
#pragma unroll
        for (int m = 0; m < D; ++m) da[m] *= (m == k) ? dl : l;
    });
}


// ---------------------------------------------------------------- the kernel
#ifdef ECH_MAXNREG
#define ECH_COOP2_BOUNDS __maxnreg__(ECH_MAXNREG)
#else
#define ECH_COOP2_BOUNDS __launch_bounds__(J2WARPS * 32, ECH_JMINB)
#endif
template <bool WITH_RES>
__global__ void ECH_COOP2_BOUNDS jacobian_coop2_kernel(const ech_dg_args A, int64_t nbatch) {
    using RC = Rec<WITH_RES>;
    constexpr int RECLEN = Rec<true>::LEN;       // one shared-memory geometry for both instantiations
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, a = lane - g * G;       // consumer role: test node a of cell g
    const bool active_lane = g < CPW;
    const int64_t fstride = A.ncell_total * NLOC;
    const int64_t rows_per_field = A.ncell_total * NLOC;
    const Smem2 W = warp_smem2(smem, warp, RECLEN);
    const RecRef R{W.rec + 2 * (active_lane ? g : 0)};
    const int gg = active_lane ? g : 0;
    // producer role: this lane evaluates slot ae of cell ge of the warp
    const int ge = lane % CPW, ae = lane / CPW;
    const RecRef RE{W.rec + 2 * ge};

    int ad[D];      // digits of my test node
#pragma unroll
    for (int k = 0; k < D; ++k) ad[k] = ((active_lane ? a : 0) / ipow(N1, D - 1 - k)) % N1;

    // PERSISTENT warps: a warp walks over batches of CPW cells; the staging of the following batches (two dependent
    // global round trips: neighbour ids, then their data) is in flight while the current one is evaluated
    const int64_t wstride = (int64_t)gridDim.x * J2WARPS;
    // a launch covers the batches of cells [A.cell_begin, A.ncell) (cell_begin is a multiple of CPW)
    int64_t batch = A.cell_begin / CPW + (int64_t)blockIdx.x * J2WARPS + warp;
    for (int k = lane; k < ECH_NCONST; k += 32) cp_async8(W.kc + k, A.consts + k);
    int ids[NLF], cs[2 * NLF];                      // (!PIPE) ids of the next batch, consumer numbering
    StageIds idn{-1, 0, 0};                         // (PIPE) ids of the batch to stage next, producer numbering
    int mb = 0;
    if (batch < nbatch) {
        if constexpr (PIPE) {
            const int64_t ce0 = batch * CPW + ge;
            const bool act0 = ae < G && ce0 < A.ncell;
            const StageIds id0 = load_ids_e(A, ae, ce0, act0);
            issue_stage_e(A, W, 0, ge, ae, ce0, act0, id0, fstride);
            const int64_t ce1 = (batch + wstride) * CPW + ge;
            idn = load_ids_e(A, ae, ce1, batch + wstride < nbatch && ae < G && ce1 < A.ncell);
        } else {
            const int64_t c0 = batch * CPW + g;
            const bool act0 = active_lane && c0 < A.ncell;
            load_ids(A, a, c0, act0, ids, cs);
            issue_stage(A, W, mb, g, a, c0, act0, ids, cs, fstride);
        }
    }

    for (; batch < nbatch; batch += wstride) {
    const int64_t c = batch * CPW + g;
    const bool active = active_lane && c < A.ncell;
    const int64_t ce = batch * CPW + ge;
    const bool active_e = ae < G && ce < A.ncell;
    const int *meta = W.meta + (mb * CPW + gg) * META_LEN;
    const int *meta_e = W.meta + (mb * CPW + ge) * META_LEN;
    const bool has_next = batch + wstride < nbatch;
    const int64_t cn = (batch + wstride) * CPW + g;
    const bool active_n = has_next && active_lane && cn < A.ncell;
    const double *nbd_b = W.nbd + (EARLY ? mb : 0) * (CPW * CELL_STRIDE) + ge * CELL_STRIDE;
    const double *fa_b = W.fa + (EARLY ? mb : 0) * FA_LEN + ge * NLF;
    const double *own = nbd_b + NLF * SLOT_LEN;
    if constexpr (EARLY) {
        if (has_next) {
            const int64_t cen = (batch + wstride) * CPW + ge;
            issue_stage_e(A, W, mb ^ 1, ge, ae, cen, ae < G && cen < A.ncell, idn, fstride);     // data of batch i+1
        }
        const int64_t cenn = (batch + 2 * wstride) * CPW + ge;
        idn = load_ids_e(A, ae, cenn, batch + 2 * wstride < nbatch && ae < G && cenn < A.ncell);   // ids of batch i+2
    } else if constexpr (PIPE) {
        // single buffer: idn holds the ids of batch i+1 (loaded one batch ahead); its data is staged in the last round
    } else {
        load_ids(A, a, cn, active_n, ids, cs);          // first round trip of the next batch (registers)
    }

    int ncol = 0, pos_self = 0;
    int64_t rp[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) rp[i] = 0;
    if (active) {
        ncol = ldg_u8(A.ncol + c);
        pos_self = ldg_u8(A.slot_self + c);
#pragma unroll
        for (int i = 0; i < NF; ++i) rp[i] = ldg_s64(A.rowptr + (int64_t)i * rows_per_field + c * NLOC + a);
    }
    if (EARLY && has_next)
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // everything but the batch just issued
    else
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    double acc[SMEM_ACC ? 1 : NF][SMEM_ACC ? 1 : NF][SMEM_ACC ? 1 : NLOC];
    if constexpr (SMEM_ACC) {
        for (int k = 0; k < NF * NF * NLOC; ++k) W.acc[k * 32 + lane] = 0.0;      // my own column of the warp's array
    } else {
        zero_blocks<true>(reinterpret_cast<double(&)[NF][NF][NLOC]>(acc));
    }
    // neighbour-block accumulators live inside one facet of one round, except for facets whose points
    // straddle two rounds (then they persist across the producer phase)
    double accn_persist[STRADDLE ? NF : 1][STRADDLE ? NF : 1][STRADDLE ? NLOC : 1];
    if constexpr (STRADDLE) zero_blocks<false>(reinterpret_cast<double(&)[NF][NF][NLOC]>(accn_persist));
    double res[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) res[i] = 0.0;

    for (int round = 0; round < NROUND; ++round) {
        // ---------------- produce: my task of this round
        const int task = round * G + ae;
        if (active_e && task < NTASK) {
            Pt p;
            p.K = W.kc;
            Geo gm;
            int idx[D];
            double w;
            double buf[RECLEN];
#pragma unroll
            for (int m = 0; m < RECLEN; ++m) buf[m] = 0.0;
            if (task < NQC) {
                cell_point(task, idx, w);
                own_side2(own, idx, p, gm);
                w *= fabs(gm.detJ);
                if constexpr (WITH_RES) {
                    Res o;
                    cell_res(p, o);
                    put_res2<MC>(buf, o, w, gm.invJ);
                    flush_record(RE, ae, buf, 0, RC::RLEN);
                }
                { put_k<MC>(buf, RC::OFF_JP, gm.invJ, gm.invJ); constexpr int ke = RC::OFF_JP + L2Of<MC>::get().klen; flush_record(RE, ae, buf, RC::OFF_JP, ke); }
                static_for<0, NF>([&](auto ic) {
                    constexpr int I = decltype(ic)::value;
                    RowJ rj;
                    cell_jac<I>(p, rj);
                    put_rowj2<MC, I>(buf, RC::OFF_JP, rj, w, gm.invJ, gm.invJ);
                    { constexpr int fb = RC::OFF_JP + L2Of<MC>::get().rowbeg[I], fe = RC::OFF_JP + L2Of<MC>::get().rowbeg[I + 1]; flush_record(RE, ae, buf, fb, fe); }
                });
            } else {
                const int ft = task - NQC;
                const int lf = ft / NQF, qf = ft - lf * NQF;
                const int nb = meta_e[lf];
                int tt[D > 1 ? D - 1 : 1];
                double ds;
                facet_point(lf, qf, idx, tt, w);
                double invJn[D][D];
#pragma unroll
                for (int m = 0; m < D; ++m)
#pragma unroll
                    for (int k = 0; k < D; ++k) invJn[m][k] = 0.0;
                if (nb >= 0) {
                    if constexpr (ECH_FAMILY_DG) {
                        neighbour_side2(nbd_b + lf * SLOT_LEN, meta_e[NLF + lf], tt, p, invJn);
                    }
                }
                own_side2(own, idx, p, gm);
                p.fa = fa_b[lf];
                facet_normal(gm, lf, p.n, ds);
                w *= ds;
                if (nb < 0) {
                    const int cls = -nb - 1;
                    if constexpr (WITH_RES) {
                        Res o;
                        efacet_res(cls, p, o);
                        put_res2<ME>(buf, o, w, gm.invJ);
                        flush_record(RE, ae, buf, 0, RC::RLEN);
                    }
                    { put_k<ME>(buf, RC::OFF_JP, gm.invJ, gm.invJ); constexpr int ke = RC::OFF_JP + L2Of<ME>::get().klen; flush_record(RE, ae, buf, RC::OFF_JP, ke); }
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        RowJ rj;
                        efacet_jac<I>(cls, p, rj);
                        put_rowj2<ME, I>(buf, RC::OFF_JP, rj, w, gm.invJ, gm.invJ);
                        { constexpr int fb = RC::OFF_JP + L2Of<ME>::get().rowbeg[I], fe = RC::OFF_JP + L2Of<ME>::get().rowbeg[I + 1]; flush_record(RE, ae, buf, fb, fe); }
                    });
                } else {
                    if constexpr (ECH_FAMILY_DG) {
                        if constexpr (WITH_RES) {
                            Res o;
                            ifacet_res(p, o);
                            put_res2<MIP>(buf, o, w, gm.invJ);
                            flush_record(RE, ae, buf, 0, RC::RLEN);
                        }
                        { put_k<MIP>(buf, RC::OFF_JP, gm.invJ, gm.invJ); constexpr int ke = RC::OFF_JP + L2Of<MIP>::get().klen; flush_record(RE, ae, buf, RC::OFF_JP, ke); }
                        { put_k<MIM>(buf, RC::OFF_JM, gm.invJ, invJn); constexpr int ke = RC::OFF_JM + L2Of<MIM>::get().klen; flush_record(RE, ae, buf, RC::OFF_JM, ke); }
                        static_for<0, NF>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            RowJ rj, rjm;
                            ifacet_jac<I>(p, rj, rjm);
                            put_rowj2<MIP, I>(buf, RC::OFF_JP, rj, w, gm.invJ, gm.invJ);
                            { constexpr int fb = RC::OFF_JP + L2Of<MIP>::get().rowbeg[I], fe = RC::OFF_JP + L2Of<MIP>::get().rowbeg[I + 1]; flush_record(RE, ae, buf, fb, fe); }
                            put_rowj2<MIM, I>(buf, RC::OFF_JM, rjm, w, gm.invJ, invJn);
                            { constexpr int fb = RC::OFF_JM + L2Of<MIM>::get().rowbeg[I], fe = RC::OFF_JM + L2Of<MIM>::get().rowbeg[I + 1]; flush_record(RE, ae, buf, fb, fe); }
                        });
                    }
                }
            }
        }
        __syncwarp();
        // every producer is done with the staged cell data: second round trip of the next batch
        if constexpr (!PIPE) {
            if (round == NROUND - 1 && has_next) issue_stage(A, W, mb ^ 1, g, a, cn, active_n, ids, cs, fstride);
        } else if constexpr (!EARLY) {
            if (round == NROUND - 1 && has_next) {
                const int64_t cen = (batch + wstride) * CPW + ge;
                issue_stage_e(A, W, mb ^ 1, ge, ae, cen, ae < G && cen < A.ncell, idn, fstride);
                const int64_t cenn = (batch + 2 * wstride) * CPW + ge;
                idn = load_ids_e(A, ae, cenn, batch + 2 * wstride < nbatch && ae < G && cenn < A.ncell);
            }
        }
        // ---------------- consume: the records of my cell into my rows (slot = compile-time task)
        if (active) {
            static_for<0, NROUND>([&](auto rc) {
                constexpr int RD = decltype(rc)::value;
                if (round == RD) {
                    if constexpr (!SMEM_ACC) {
                    double(&accR)[NF][NF][NLOC] = reinterpret_cast<double(&)[NF][NF][NLOC]>(acc);
                    // residual rows and own-block contributions of cell / exterior-facet points
                    static_for<0, G>([&](auto sc) {
                        constexpr int s = decltype(sc)::value;
                        constexpr int TK = RD * G + s;
                        if constexpr (TK < NTASK) {
                            constexpr TaskPt tp = task_point(TK);
                            double pa, da[D];
                            test_values<TK>(ad, pa, da);
                            if constexpr (WITH_RES) {
#pragma unroll
                                for (int i = 0; i < NF; ++i) {
                                    double t = R(s, i) * pa;
#pragma unroll
                                    for (int m = 0; m < D; ++m) t += R(s, NF + i * D + m) * da[m];
                                    res[i] += t;
                                }
                            }
                            double phi2[NLOC], d2[NLOC][D];     // unused on the own side
                            if constexpr (tp.kind == 0) {
                                take_all2<MC, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accR);
                            } else {
                                if (meta[tp.lf] < 0) take_all2<ME, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accR);
                            }
                        }
                    });
                    // interior facets, facet by facet
                    if constexpr (ECH_FAMILY_DG && NBR_PER_EQ) {
                        // large elements: the neighbour-block accumulators of ALL equations (NF*NF*NLOC doubles) next to
                        // the own-block ones do not fit the register file (3D, three fields: 7 KB of spills per
                        // thread).  Own part for all equations first, then the neighbour part one equation at a time
                        // (NF*NLOC accumulators), the neighbour's reference basis recomputed per equation.
                        static_for<0, NLF>([&](auto lc) {
                            constexpr int lf = decltype(lc)::value;
                            constexpr int r0 = facet_first_round(lf);
                            if constexpr (RD == r0) {
                                if (meta[lf] >= 0) {
                                    static_for<0, NQF>([&](auto qc) {
                                        constexpr int qf = decltype(qc)::value;
                                        constexpr int TK = NQC + lf * NQF + qf;
                                        constexpr int s = TK - RD * G;
                                        double pa, da[D];
                                        test_values<TK>(ad, pa, da);
                                        double phi2[NLOC], d2[NLOC][D];      // unused on the own side
                                        take_all2<MIP, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accR);
                                    });
                                    static_for<0, NF>([&](auto ic) {
                                        constexpr int I = decltype(ic)::value;
                                        double accnI[NF][NLOC];
                                        static_for<0, NF>([&](auto jc) {
                                            constexpr int j = decltype(jc)::value;
                                            if constexpr (NBR_BLOCK[I][j]) {
#pragma unroll
                                                for (int b = 0; b < NLOC; ++b) accnI[j][b] = 0.0;
                                            }
                                        });
                                        static_for<0, NQF>([&](auto qc) {
                                            constexpr int qf = decltype(qc)::value;
                                            constexpr int TK = NQC + lf * NQF + qf;
                                            constexpr int s = TK - RD * G;
                                            double pa, da[D];
                                            test_values<TK>(ad, pa, da);
                                            double phi2[NLOC], d2[NLOC][D];
                                            nb_basis<TK>(meta[NLF + lf], phi2, d2);
                                            take_row2<MIM, I, -1>(R, s, RC::OFF_JM, pa, da, phi2, d2, accnI);
                                        });
                                        store_row_blocks<false, I>(A.out, rp[I], ncol, meta[2 * NLF + lf], accnI);
                                    });
                                }
                            }
                        });
                    } else if constexpr (ECH_FAMILY_DG) {
                        static_for<0, NLF>([&](auto lc) {
                            constexpr int lf = decltype(lc)::value;
                            constexpr int r0 = facet_first_round(lf), r1 = facet_last_round(lf);
                            if constexpr (RD >= r0 && RD <= r1) {
                                if (meta[lf] >= 0) {
                                    double accn_local[r0 == r1 ? NF : 1][r0 == r1 ? NF : 1][r0 == r1 ? NLOC : 1];
                                    double(&accn)[NF][NF][NLOC] = *reinterpret_cast<double(*)[NF][NF][NLOC]>(
                                        r0 == r1 ? &accn_local[0][0][0] : &accn_persist[0][0][0]);
                                    if constexpr (r0 == r1) zero_blocks<false>(accn);
                                    static_for<0, NQF>([&](auto qc) {
                                        constexpr int qf = decltype(qc)::value;
                                        constexpr int TK = NQC + lf * NQF + qf;
                                        if constexpr (TK / G == RD) {
                                            constexpr int s = TK - RD * G;
                                            double pa, da[D];
                                            test_values<TK>(ad, pa, da);
                                            double phi2[NLOC], d2[NLOC][D];
                                            take_all2<MIP, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accR);
                                            nb_basis<TK>(meta[NLF + lf], phi2, d2);
                                            take_all2<MIM, -1>(R, s, RC::OFF_JM, pa, da, phi2, d2, accn);
                                        }
                                    });
                                    if constexpr (RD == r1) {
                                        store_blocks<false>(A.out, rp, ncol, meta[2 * NLF + lf], accn);
                                        if constexpr (r0 != r1) zero_blocks<false>(accn);
                                    }
                                }
                            }
                        });
                    }
                    } else {
                    // ---- shared-memory accumulators: residual first, then one pass over the records per equation
                    if constexpr (WITH_RES) {
                        static_for<0, G>([&](auto sc) {
                            constexpr int s = decltype(sc)::value;
                            constexpr int TK = RD * G + s;
                            if constexpr (TK < NTASK) {
                                double pa, da[D];
                                test_values<TK>(ad, pa, da);
#pragma unroll
                                for (int i = 0; i < NF; ++i) {
                                    double t = R(s, i) * pa;
#pragma unroll
                                    for (int m = 0; m < D; ++m) t += R(s, NF + i * D + m) * da[m];
                                    res[i] += t;
                                }
                            }
                        });
                    }
                    static_for<0, NF>([&](auto ic) {
                        constexpr int I = decltype(ic)::value;
                        double accI[NF][NLOC];
                        acc_load<I>(W.acc, lane, accI);
                        double phi2[NLOC], d2[NLOC][D];
                        static_for<0, G>([&](auto sc) {
                            constexpr int s = decltype(sc)::value;
                            constexpr int TK = RD * G + s;
                            if constexpr (TK < NTASK) {
                                constexpr TaskPt tp = task_point(TK);
                                double pa, da[D];
                                test_values<TK>(ad, pa, da);
                                if constexpr (tp.kind == 0) {
                                    take_row2<MC, I, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accI);
                                } else {
                                    if (meta[tp.lf] < 0) take_row2<ME, I, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accI);
                                }
                            }
                        });
                        if constexpr (ECH_FAMILY_DG) {
                            static_for<0, NLF>([&](auto lc) {
                                constexpr int lf = decltype(lc)::value;
                                constexpr int r0 = facet_first_round(lf);        // == last round (no straddling here)
                                if constexpr (RD == r0) {
                                    if (meta[lf] >= 0) {
                                        double accnI[NF][NLOC];
                                        static_for<0, NF>([&](auto jc) {
                                            constexpr int j = decltype(jc)::value;
                                            if constexpr (NBR_BLOCK[I][j]) {
#pragma unroll
                                                for (int b = 0; b < NLOC; ++b) accnI[j][b] = 0.0;
                                            }
                                        });
                                        static_for<0, NQF>([&](auto qc) {
                                            constexpr int qf = decltype(qc)::value;
                                            constexpr int TK = NQC + lf * NQF + qf;
                                            constexpr int s = TK - RD * G;
                                            double pa, da[D];
                                            test_values<TK>(ad, pa, da);
                                            take_row2<MIP, I, TK>(R, s, RC::OFF_JP, pa, da, phi2, d2, accI);
                                            nb_basis<TK>(meta[NLF + lf], phi2, d2);
                                            take_row2<MIM, I, -1>(R, s, RC::OFF_JM, pa, da, phi2, d2, accnI);
                                        });
                                        store_row_blocks<false, I>(A.out, rp[I], ncol, meta[2 * NLF + lf], accnI);
                                    }
                                }
                            });
                        }
                        acc_store<I>(W.acc, lane, accI);
                    });
                    }
                }
            });
        }
        __syncwarp();
    }
    if (active) {
        if constexpr (SMEM_ACC) {
            static_for<0, NF>([&](auto ic) {
                constexpr int I = decltype(ic)::value;
                double accI[NF][NLOC];
                acc_load<I>(W.acc, lane, accI);
                store_row_blocks<true, I>(A.out, rp[I], ncol, pos_self, accI);
            });
        } else {
            store_blocks<true>(A.out, rp, ncol, pos_self, reinterpret_cast<double(&)[NF][NF][NLOC]>(acc));
        }
        if constexpr (WITH_RES) {
#pragma unroll
            for (int i = 0; i < NF; ++i) A.out2[i * rows_per_field + c * NLOC + a] = res[i];
        }
    }
    mb ^= 1;
    }   // batches
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr int J2SMEM = J2WARPS * warp_smem2_doubles(Rec<true>::LEN) * 8;

static int g_coop2_ctas_per_sm[2] = {0, 0};
static int g_coop2_sms = 0;

static int coop2_setup() {
    static int done = 0;
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(jacobian_coop2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, J2SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(jacobian_coop2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, J2SMEM);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_coop2_ctas_per_sm[1], jacobian_coop2_kernel<true>,
                                                          J2WARPS * 32, J2SMEM);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_coop2_ctas_per_sm[0], jacobian_coop2_kernel<false>,
                                                          J2WARPS * 32, J2SMEM);
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_coop2_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess || g_coop2_ctas_per_sm[0] < 1 || g_coop2_ctas_per_sm[1] < 1) {
        fprintf(stderr, "echem_b200: jacobian_coop2_kernel setup (%d B shared) failed: %s\n", J2SMEM,
                cudaGetErrorString(e));
        return e != cudaSuccess ? (int)e : (int)cudaErrorLaunchOutOfResources;
    }
    done = 1;
    return 0;
}

static int launch_jacobian_coop2(const ech_dg_args *a, cudaStream_t s) {
    if (a->ncell == 0) return 0;
    int e = coop2_setup();
    if (e) return e;
    if (a->cell_begin < 0 || a->cell_begin % CPW) return (int)cudaErrorInvalidValue;
    if (a->cell_begin >= a->ncell) return 0;
    const int64_t nbatch = (a->ncell + CPW - 1) / CPW;       // end batch (exclusive)
    const int64_t need = (nbatch - a->cell_begin / CPW + J2WARPS - 1) / J2WARPS;
    const int fused = a->out2 ? 1 : 0;
    // ECH_COOP2_WAVES = w > 0: persistent grid of w waves of resident CTAs (SMs x CTAs per SM), each warp strides
    // over the batches and stages batch i+1 while it consumes batch i.  Measured on cfg2 (profiles/): 2.78 ms
    // against 2.64 ms for one batch per warp (the default, w = 0) -- with 2 warps per scheduler the start-up
    // stagger of short-lived CTAs de-phases the warps better than the software pipeline hides the staging.
#ifndef ECH_COOP2_WAVES
#define ECH_COOP2_WAVES (PIPE ? 2 : 0)
#endif
    // run-time override for tuning: waves of resident CTAs; 0 = one batch per warp
    static const int waves = getenv("ECH_COOP2_WAVES") ? atoi(getenv("ECH_COOP2_WAVES")) : ECH_COOP2_WAVES;
    const int64_t resident = (int64_t)g_coop2_sms * g_coop2_ctas_per_sm[fused] * waves;
    const int64_t grid = (waves == 0 || need < resident) ? need : resident;
    if (fused)
        jacobian_coop2_kernel<true><<<(unsigned)grid, J2WARPS * 32, J2SMEM, s>>>(*a, nbatch);
    else
        jacobian_coop2_kernel<false><<<(unsigned)grid, J2WARPS * 32, J2SMEM, s>>>(*a, nbatch);
    e = (int)cudaGetLastError();
    if (e) fprintf(stderr, "echem_b200: jacobian_coop2_kernel<<<%lld, %d, %d>>> failed: %s\n", (long long)grid,
                   J2WARPS * 32, J2SMEM, cudaGetErrorString((cudaError_t)e));
    return e;
}

}  // namespace v2
}  // namespace echdg
