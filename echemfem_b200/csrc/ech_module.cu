// Translation unit of one physics module: generated integrands + hand-written DG skeleton.
// Built by echemfem_b200/build.py as
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC
//        -include <generated header> ech_module.cu -o libech_phys_<key>.so
#if ECH_FAMILY_DG
#include "ech_dg_coop2.cuh"
#else
#include "ech_cg.cuh"
#endif

extern "C" {

int ech_mod_describe(int32_t *info, int capacity) {
    using namespace echgen;
    const int need = 8 + 2 * NF * NF;
    if (capacity < need) return -need;
    info[0] = D;
    info[1] = P;
    info[2] = NF;
    info[3] = ECH_NAUX;
    info[4] = ECH_NCONST;
    info[5] = NQ;
    info[6] = ECH_NCLS;
    info[7] = ECH_FAMILY_DG;
    for (int i = 0; i < NF; ++i)
        for (int j = 0; j < NF; ++j) {
            info[8 + i * NF + j] = OWN_BLOCK[i][j];
            info[8 + NF * NF + i * NF + j] = NBR_BLOCK[i][j];
        }
    return need;
}

// variant 0 (default): cooperative kernels (one cell per NLOC lanes, shared quadrature work), Jacobian with
//            reference-space records (ech_dg_coop2.cuh);
// variant 1: row-per-thread kernels (kept for elements with more than 32 local nodes and as a cross-check);
// variant 2: cooperative kernels, Jacobian with physical-space records (ech_dg_coop.cuh, first generation)
static int g_variant = 0;

void ech_mod_set_variant(int v) { g_variant = v; }

#if ECH_FAMILY_DG
static bool use_coop() { return echdg::COOP_OK && g_variant != 1; }

static bool use_coop2() { return use_coop() && g_variant == 0 && echdg::NLOC <= 9; }

// the cooperative kernels (both generations, residual and Jacobian) take a cell range
int ech_mod_range_align() { return use_coop() ? echdg::CPW : 0; }

int ech_mod_residual(const ech_dg_args *a, void *stream) {
    if (use_coop()) return echdg::launch_residual_coop(a, (cudaStream_t)stream);
    if (a->cell_begin) return (int)cudaErrorInvalidValue;
    return echdg::launch_residual(a, (cudaStream_t)stream);
}

int ech_mod_jacobian(const ech_dg_args *a, void *stream) {
    // fused residual + Jacobian when a->out2 is set
    // reference-space records pay up to 9 local nodes (Q1 2D/3D, Q2 2D: cfg2 Q2 12.9 ms against 16.5 ms); beyond
    // that the fully unrolled compile-time bases outgrow the register file (cfg2 Q3: 52 ms against 34 ms)
    if (use_coop2()) return echdg::v2::launch_jacobian_coop2(a, (cudaStream_t)stream);
    if (use_coop()) return echdg::launch_jacobian_coop(a, (cudaStream_t)stream);
    if (a->cell_begin) return (int)cudaErrorInvalidValue;
    if (a->out2) {
        ech_dg_args r = *a;
        r.out = a->out2;
        r.out2 = nullptr;
        int e = echdg::launch_residual(&r, (cudaStream_t)stream);
        if (e) return e;
    }
    return echdg::launch_jacobian(a, (cudaStream_t)stream);
}

// which: 0 residual, 1 Jacobian, 2 fused residual + Jacobian
int ech_mod_launches(int which) {
    if (which == 0) return 1;
    if (use_coop()) return 1;
    return echgen::NF + (which == 2 ? 1 : 0);
}
#else
int ech_mod_range_align() { return 0; }

int ech_mod_residual(const ech_dg_args *a, void *stream) {
    (void)g_variant;
    if (a->cell_begin) return (int)cudaErrorInvalidValue;
    return echcg::launch_cg_residual(a, (cudaStream_t)stream);
}

int ech_mod_jacobian(const ech_dg_args *a, void *stream) {
    if (a->cell_begin) return (int)cudaErrorInvalidValue;
    if (a->out2) {
        ech_dg_args r = *a;
        r.out = a->out2;
        r.out2 = nullptr;
        int e = echcg::launch_cg_residual(&r, (cudaStream_t)stream);
        if (e) return e;
    }
    return echcg::launch_cg_jacobian(a, (cudaStream_t)stream);
}

int ech_mod_launches(int which) { return which == 0 ? 1 : echgen::NF + (which == 2 ? 1 : 0); }
#endif

}  // extern "C"
