// Gaussian kernel reductions as pair-engine Ops (forward).
//
// Replaces the KeOps-generated kernels behind the reference's GaussKernel reductions
// (/root/reference/diffICP/tools/kernel.py:125-168; torch twins :177-215, :284-292) and check_coverage (:324-329).
//
// Scaled coordinates: with kappa = sqrt(log2(e)/2)/sigma and z' = kappa (x_i - y_j) the Gaussian is
// K = 2^(-|z'|^2): one MUFU.EX2 per pair, no per-pair multiply by 1/(2 sigma^2).  All coordinates are taken
// relative to a common origin (the first column point) before scaling, which keeps fp32 rounding of the
// pre-scaled coordinates small when the cloud sits far from 0.  Constants:
//   s = 1/sigma^2,  alpha = s/kappa,  beta = s/kappa^2 = 2 ln 2.
// One templated Op computes any subset MASK of the outputs in a single sweep (one exponential per pair
// for the whole subset).
#pragma once
#include "pair_engine.cuh"

namespace dicp {

enum KsumOut : unsigned {
    K_BASE = 1u, K_REDSCAL = 2u, K_RED = 4u, K_GRAD = 8u, K_DD = 16u, K_GEND = 32u, K_HESS = 64u,
    K_LAP = 128u, K_GRADLAP = 256u, K_MINSQ = 512u, K_DOT = 1024u
};

struct KsumParams {
    const float *x, *y;      // rows (M,D), columns (N,D)
    const float *b, *d;      // column vectors (N,D), column scalars (N,)
    const float *c;          // row vectors (M,D)
    const float* origin;     // D floats: common origin subtracted before scaling (= y)
    float kappa, s, alpha, beta;
    float *o_base, *o_redscal, *o_red, *o_grad, *o_dd, *o_gend, *o_hess, *o_lap, *o_gradlap, *o_minsq, *o_dot;
};

template <int D, unsigned MASK, int R_ = 2>
struct KsumOp {
    using Params = KsumParams;
    static constexpr bool PACKED = false;
    static constexpr int THREADS = 128, MINB = 1, R = R_, TILE = 128;
    static constexpr bool NEED_B = (MASK & (K_RED | K_DD | K_GEND | K_HESS | K_DOT)) != 0;
    static constexpr bool NEED_D = (MASK & K_REDSCAL) != 0;
    static constexpr bool NEED_C = (MASK & (K_GEND | K_HESS)) != 0;
    static constexpr bool NEED_K = (MASK & ~K_MINSQ) != 0;
    static constexpr int COLN = D + (NEED_B ? D : 0) + (NEED_D ? 1 : 0);
    static constexpr int COLF4 = (COLN + 3) / 4;
    // accumulator layout
    static constexpr int A_BASE = 0;
    static constexpr int A_REDSCAL = A_BASE + ((MASK & K_BASE) ? 1 : 0);
    static constexpr int A_RED = A_REDSCAL + ((MASK & K_REDSCAL) ? 1 : 0);
    static constexpr int A_GRAD = A_RED + ((MASK & K_RED) ? D : 0);
    static constexpr int A_DD = A_GRAD + ((MASK & K_GRAD) ? D : 0);
    static constexpr int A_GEND = A_DD + ((MASK & K_DD) ? D : 0);
    static constexpr int A_HESS = A_GEND + ((MASK & K_GEND) ? D : 0);
    static constexpr int A_LAP = A_HESS + ((MASK & K_HESS) ? D : 0);
    static constexpr int A_GRADLAP = A_LAP + ((MASK & K_LAP) ? 1 : 0);
    static constexpr int A_MINSQ = A_GRADLAP + ((MASK & K_GRADLAP) ? D : 0);
    static constexpr int A_DOT = A_MINSQ + ((MASK & K_MINSQ) ? 1 : 0);
    static constexpr int NACC = A_DOT + ((MASK & K_DOT) ? 1 : 0);
    static constexpr int NSCAL = 0;

    struct Row {
        float x[D];
        float c[NEED_C ? D : 1];
    };

    static DICP_HD void pack_col(const Params& p, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = (p.y[(size_t)j * D + k] - p.origin[k]) * p.kappa;
            if (NEED_B) {
#pragma unroll
                for (int k = 0; k < D; ++k) c[D + k] = p.b[(size_t)j * D + k];
            }
            if (NEED_D) c[D + (NEED_B ? D : 0)] = p.d[j];
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
        }
    }
    static DICP_HD void load_row(const Params& p, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) r.x[k] = (p.x[(size_t)i * D + k] - p.origin[k]) * p.kappa;
        if (NEED_C) {
#pragma unroll
            for (int k = 0; k < D; ++k) r.c[k] = p.c[(size_t)i * D + k];
        }
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
        if (MASK & K_MINSQ) a[A_MINSQ] = 3.0e38f;
    }
    static DICP_HD void combine(float* a, const float* b) {
        const float m0 = (MASK & K_MINSQ) ? a[A_MINSQ] : 0.f;
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
        if (MASK & K_MINSQ) a[A_MINSQ] = fminf(m0, b[A_MINSQ]);
    }
    static DICP_HD void pair(const Params& p, const Row& r, const float* c, float* a) {
        float z[D];
        float r2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.x[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
        }
        if (MASK & K_MINSQ) a[A_MINSQ] = fminf(a[A_MINSQ], r2);
        if (!NEED_K) return;
        const float K = ex2_neg(r2);
        if (MASK & K_BASE) a[A_BASE] += K;
        if (MASK & K_REDSCAL) a[A_REDSCAL] = fmaf(K, c[D + (NEED_B ? D : 0)], a[A_REDSCAL]);
        if (MASK & K_RED) {
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_RED + k] = fmaf(K, c[D + k], a[A_RED + k]);
        }
        if (MASK & K_GRAD) {
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_GRAD + k] = fmaf(K, z[k], a[A_GRAD + k]);
        }
        if (MASK & K_DD) {
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_DD + k] = fmaf(K * z[k], c[D + k], a[A_DD + k]);
        }
        if (MASK & K_GEND) {
            float w = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) w = fmaf(r.c[k], c[D + k], w);
            const float Kw = K * w;
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_GEND + k] = fmaf(Kw, z[k], a[A_GEND + k]);
        }
        if (MASK & K_HESS) {
            float e[D];
            float ze = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                e[k] = r.c[k] - c[D + k];
                ze = fmaf(z[k], e[k], ze);
            }
            const float t = K * (p.beta * ze);
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_HESS + k] = fmaf(t, z[k], fmaf(-K, e[k], a[A_HESS + k]));
        }
        if (MASK & K_DOT) {
            float zb = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) zb = fmaf(z[k], c[D + k], zb);
            a[A_DOT] = fmaf(K, zb, a[A_DOT]);
        }
        if (MASK & K_LAP) a[A_LAP] = fmaf(K, fmaf(p.beta, r2, -(float)D), a[A_LAP]);
        if (MASK & K_GRADLAP) {
            const float t = K * fmaf(p.beta, r2, -(float)(D + 2));
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_GRADLAP + k] = fmaf(t, z[k], a[A_GRADLAP + k]);
        }
    }
    static DICP_HD void finish(const Params& p, int i, const Row& r, const float* a, float* /*scal*/) {
        if (MASK & K_BASE) p.o_base[i] = a[A_BASE];
        if (MASK & K_REDSCAL) p.o_redscal[i] = a[A_REDSCAL];
        if (MASK & K_LAP) p.o_lap[i] = p.s * a[A_LAP];
        if (MASK & K_MINSQ) p.o_minsq[i] = a[A_MINSQ] / (p.kappa * p.kappa);
        if (MASK & K_DOT) p.o_dot[i] = p.alpha * a[A_DOT];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (MASK & K_RED) p.o_red[o] = a[A_RED + k];
            if (MASK & K_GRAD) p.o_grad[o] = -p.alpha * a[A_GRAD + k];
            if (MASK & K_DD) p.o_dd[o] = -p.alpha * a[A_DD + k];
            if (MASK & K_GEND) p.o_gend[o] = -p.alpha * a[A_GEND + k];
            if (MASK & K_HESS) p.o_hess[o] = p.s * a[A_HESS + k];
            if (MASK & K_GRADLAP) p.o_gradlap[o] = -p.s * p.alpha * a[A_GRADLAP + k];
        }
    }
};

}  // namespace dicp
