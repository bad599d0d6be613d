// Fused Hamiltonian right-hand side of the LDDMM geodesic ODE and its adjoint (VJP), as pair-engine Ops.
//
// Replaces, with ONE exponential per visited pair, the 2 / 4 / 7 separate KeOps reductions that
// LDDMMModel.ODE issues per evaluation (/root/reference/diffICP/core/LDDMM.py:176-227: KRed, GenDKRed,
// GradKRed, HessKRed, GradLapKRed, LapKRed), and the autograd-generated backward reductions of the same
// (reverse-mode through the integrator loop, tools/optim.py:34-47).
//
// Notation (scaled coordinates, see ops_ksum.cuh): z' = kappa (row - col), K = 2^-|z'|^2,
// s = 1/sigma^2, alpha = s/kappa, beta = s/kappa^2 = 2 ln 2, eta = 1/lambda (logdet model) or 0.
//
// Forward, rows = cols = support points (q,p):
//   vq_i   = sum_j K p_j + eta*alpha * sum_j K z'
//   dp_i   = -Gq_i = sum_j K [alpha w + eta s beta (z'.e) - eta^2 s alpha (beta r'^2 - (D+2))] z' - eta s (p_i S0 - Vp_i)
//            with w = p_i.p_j, e = p_i - p_j, S0 = sum_j K, Vp = sum_j K p_j
//   row scalars: A_i = p_i.Vp_i,  B_i = -alpha p_i.Z_i (Z = sum_j K z'),  C_i = s (beta sum_j K r'^2 - D S0)
//            H = A/2 - eta B - eta^2 C/2,   dcost(x=None) = B + eta C
// Forward, rows = data points x, cols = (q,p):
//   vx_k   = sum_j K p_j + eta*alpha * sum_j K z'
//   dcost  = alpha sum_kj K (p_j.z') + eta s sum_kj K (beta r'^2 - D)
// Adjoint: formulas above each Op, derivation in DESIGN.md §5.3; all checked against torch autograd of the oracle
// (tests/test_host_emulation.py).
//
// Every `pair` is written ONCE against the lane-generic interface of common.cuh: instantiated with V = F2 it processes
// TWO columns per call with packed-fp32 instructions (FFMA2 / FADD2 / FMUL2; row-side scalars ride the broadcast
// operand form), which is what the device kernel (pair_kernel_p) uses; with V = float it is the one-column form used for
// an odd trailing column and by the CPU emulation of the tests.  PACKED Ops declare NF, the (even) number of floats of a
// column record.
#pragma once
#include "pair_engine.cuh"

// tuning knobs (overridable at compile time for sweeps: -DDICP_RHS_R=4 ...)
#ifndef DICP_RHS_R
#define DICP_RHS_R 2
#endif
#ifndef DICP_RHS_THREADS
#define DICP_RHS_THREADS 64
#endif
#ifndef DICP_RHS_TILE
#define DICP_RHS_TILE 128
#endif
#ifndef DICP_RHS_MINB
#define DICP_RHS_MINB 1
#endif

namespace dicp {

struct RhsParams {
    const float *q, *p, *x;        // state: support points (M,D), momenta (M,D), data points (Nx,D)
    const float *a, *u, *wx;       // cotangents of vq (M,D), dp (M,D), vx (Nx,D)
    const float *gc;               // device scalar: cotangent of dcost (null => 0)
    const float *origin;           // D floats (= q)
    float kappa, s, alpha, beta, eta;
    float *vq, *dp, *vx;           // forward outputs
    float *gq, *gp, *gx;           // adjoint outputs
    int accumulate;                // adjoint outputs: 0 overwrite, 1 add to existing
    int col0;                      // (q,q) ops: index offset of the COLUMN points relative to the row points (blocked symmetric
                                   // evaluation of very large sets: rows from one super-block, columns from another); else 0
};

#define DICP_RHS_COMMON(NF_)                                                                                   \
    using Params = RhsParams;                                                                                 \
    static constexpr bool PACKED = true;                                                                      \
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;      \
    static constexpr int NF = (NF_);                                                                          \
    static constexpr int COLF4 = (NF + 3) / 4;                                                                \
    static DICP_HD void init(float* a) {                                                                      \
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;                                                            \
    }                                                                                                         \
    static DICP_HD void combine(float* a, const float* b) {                                                   \
        for (int k = 0; k < NACC; ++k) a[k] += b[k];                                                          \
    }                                                                                                         \
    /* packed accumulators of pair_kernel_p: plain sums, one partial sum per half */                          \
    static constexpr bool PAD_NULL = false;                                                                   \
    static DICP_HD void init_packed(F2* a) {                                                                  \
        for (int k = 0; k < NACC; ++k) a[k] = f2(0.f, 0.f);                                                   \
    }                                                                                                         \
    static DICP_HD void unpack_acc(const F2* a, float* out) {                                                 \
        for (int k = 0; k < NACC; ++k) out[k] = f2_sum(a[k]);                                                 \
    }

// column record (q', p): used by RhsQQ, RhsXQ, AdjXQx*
template <int D>
DICP_HD void pack_qp(const RhsParams& P, int j, int N, float* c, int nfloat) {
    for (int k = 0; k < nfloat; ++k) c[k] = 0.f;
    if (j < N) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            c[k] = (P.q[(size_t)j * D + k] - P.origin[k]) * P.kappa;
            c[D + k] = P.p[(size_t)j * D + k];
        }
    }
}
// column record (q', p, a, u): used by AdjQQ*
template <int D>
DICP_HD void pack_qpau(const RhsParams& P, int j, int N, float* c, int nfloat) {
    for (int k = 0; k < nfloat; ++k) c[k] = 0.f;
    if (j < N) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = ((size_t)j + (size_t)P.col0) * D + k;
            c[k] = (P.q[o] - P.origin[k]) * P.kappa;
            c[D + k] = P.p[o];
            c[2 * D + k] = P.a[o];
            c[3 * D + k] = P.u[o];
        }
    }
}
// column record (x', wx): used by AdjXQq*
template <int D>
DICP_HD void pack_xw(const RhsParams& P, int j, int N, float* c, int nfloat) {
    for (int k = 0; k < nfloat; ++k) c[k] = 0.f;
    if (j < N) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            c[k] = (P.x[(size_t)j * D + k] - P.origin[k]) * P.kappa;
            c[D + k] = P.wx[(size_t)j * D + k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward, (q,q)
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, bool ETA, int R_ = DICP_RHS_R>
struct RhsQQ {
    static constexpr int A_V = 0, A_T = D, A_Z = 2 * D;
    static constexpr bool NEEDZ = DIV || ETA;
    static constexpr int A_S0 = A_Z + (NEEDZ ? D : 0);
    static constexpr int A_R2 = A_S0 + (ETA ? 1 : 0);
    static constexpr int NACC = A_R2 + (ETA ? 1 : 0);
    static constexpr int NACC_COL = NACC;
    static constexpr int NSCAL = 3;   // A, B, C
    DICP_RHS_COMMON(2 * D)
    struct Row { float q[D], p[D]; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* a) {
        V z[D];
        V r2, w;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            w = k == 0 ? vmul(vbc<V>(r.p[k]), c[D + k]) : vfma(vbc<V>(r.p[k]), c[D + k], w);
        }
        const V K = vex2n(r2);
        V coef;
        if (ETA) {
            V ze;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const V ek = vsub(vbc<V>(r.p[k]), c[D + k]);
                ze = k == 0 ? vmul(z[k], ek) : vfma(z[k], ek, ze);
            }
            // alpha w + eta s beta (z'.e) - eta^2 s alpha (beta r'^2 - (D+2))
            const float c1 = P.eta * P.s * P.beta, c2 = P.eta * P.eta * P.s * P.alpha;
            const V t1 = vfma(vbc<V>(-c2 * P.beta), r2, vbc<V>(c2 * (float)(D + 2)));
            coef = vfma(vbc<V>(P.alpha), w, vfma(vbc<V>(c1), ze, t1));
            a[A_S0] = vadd(a[A_S0], K);
            a[A_R2] = vfma(K, r2, a[A_R2]);
        } else {
            coef = w;   // alpha applied in finish
        }
        const V Kc = vmul(K, coef);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[A_V + k] = vfma(K, c[D + k], a[A_V + k]);
            a[A_T + k] = vfma(Kc, z[k], a[A_T + k]);
            if (NEEDZ) a[A_Z + k] = vfma(K, z[k], a[A_Z + k]);
        }
    }
    // Both orientations of an unordered pair from ONE evaluation (symmetric engine): K, r'^2, w and the z' coefficient are
    // even under the swap of the two points, z' is odd: the column's sums get K p_i, -Kc z', -K z', K, K r'^2.
    template <class V, bool MASKED = false>
    static DICP_HD void pair_sym(const Params& P, const Row& r, const V* c, V* a, V* ca, V km = V()) {
        V z[D];
        V r2, w;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            w = k == 0 ? vmul(vbc<V>(r.p[k]), c[D + k]) : vfma(vbc<V>(r.p[k]), c[D + k], w);
        }
        V K = vex2n(r2);
        if (MASKED) K = vmul(K, km);
        V coef;
        if (ETA) {
            V ze;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const V ek = vsub(vbc<V>(r.p[k]), c[D + k]);
                ze = k == 0 ? vmul(z[k], ek) : vfma(z[k], ek, ze);
            }
            const float c1 = P.eta * P.s * P.beta, c2 = P.eta * P.eta * P.s * P.alpha;
            const V t1 = vfma(vbc<V>(-c2 * P.beta), r2, vbc<V>(c2 * (float)(D + 2)));
            coef = vfma(vbc<V>(P.alpha), w, vfma(vbc<V>(c1), ze, t1));
            a[A_S0] = vadd(a[A_S0], K);
            ca[A_S0] = vadd(ca[A_S0], K);
            a[A_R2] = vfma(K, r2, a[A_R2]);
            ca[A_R2] = vfma(K, r2, ca[A_R2]);
        } else {
            coef = w;
        }
        const V Kc = vmul(K, coef);
        const V nKc = vmul(Kc, vbc<V>(-1.f));
        V nK;
        if (NEEDZ) nK = vmul(K, vbc<V>(-1.f));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[A_V + k] = vfma(K, c[D + k], a[A_V + k]);
            ca[A_V + k] = vfma(K, vbc<V>(r.p[k]), ca[A_V + k]);
            a[A_T + k] = vfma(Kc, z[k], a[A_T + k]);
            ca[A_T + k] = vfma(nKc, z[k], ca[A_T + k]);
            if (NEEDZ) {
                a[A_Z + k] = vfma(K, z[k], a[A_Z + k]);
                ca[A_Z + k] = vfma(nK, z[k], ca[A_Z + k]);
            }
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* a, float* scal) {
        float A = 0.f, B = 0.f, C = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            float v = a[A_V + k];
            A = fmaf(r.p[k], v, A);
            if (NEEDZ) B = fmaf(r.p[k], a[A_Z + k], B);
            float dpk;
            if (ETA) {
                dpk = a[A_T + k] - P.eta * P.s * (r.p[k] * a[A_S0] - v);
                v = fmaf(P.eta * P.alpha, a[A_Z + k], v);
            } else {
                dpk = P.alpha * a[A_T + k];
            }
            P.vq[o] = v;
            P.dp[o] = dpk;
        }
        if (ETA) C = P.s * (P.beta * a[A_R2] - (float)D * a[A_S0]);
        scal[0] = A;
        scal[1] = -P.alpha * B;
        scal[2] = C;
    }
};

// ------------------------------------------------------------------------------------------------
// forward, (x,q): rows = data points, cols = (q,p)
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, bool ETA, int R_ = DICP_RHS_R>
struct RhsXQ {
    static constexpr int A_V = 0, A_DS = D;
    static constexpr int A_Z = A_DS + (DIV ? 1 : 0);
    static constexpr int A_S0 = A_Z + (ETA ? D : 0);
    static constexpr int A_R2 = A_S0 + (ETA ? 1 : 0);
    static constexpr int NACC = A_R2 + (ETA ? 1 : 0);
    static constexpr int NSCAL = 1;   // dcost contribution
    DICP_RHS_COMMON(2 * D)
    struct Row { float x[D]; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* a) {
        V z[D];
        V r2, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.x[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            if (DIV) pz = k == 0 ? vmul(c[D + k], z[k]) : vfma(c[D + k], z[k], pz);
        }
        const V K = vex2n(r2);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[A_V + k] = vfma(K, c[D + k], a[A_V + k]);
            if (ETA) a[A_Z + k] = vfma(K, z[k], a[A_Z + k]);
        }
        if (DIV) a[A_DS] = vfma(K, pz, a[A_DS]);
        if (ETA) {
            a[A_S0] = vadd(a[A_S0], K);
            a[A_R2] = vfma(K, r2, a[A_R2]);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* a, float* scal) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float v = a[A_V + k];
            if (ETA) v = fmaf(P.eta * P.alpha, a[A_Z + k], v);
            P.vx[(size_t)i * D + k] = v;
        }
        float dc = 0.f;
        if (DIV) dc = P.alpha * a[A_DS];
        if (ETA) dc += P.eta * P.s * (P.beta * a[A_R2] - (float)D * a[A_S0]);
        scal[0] = dc;
    }
};

// ------------------------------------------------------------------------------------------------
// adjoint, (q,q), eta = 0.  rows and cols = (q, p, a, u).
//   gp_i = sum_j K a_j + alpha sum_j K (du.z') p_j                       [- gc alpha sum_j K z'            if DIV]
//   gq_i = - sum_j K [alpha((a_i.p_j)+(a_j.p_i)) + s beta w (du.z')] z' + s sum_j K w du
//                                                                       [- gc s sum_j K (dp - beta (dp.z') z') if DIV]
//   with du = u_i - u_j, dp = p_i - p_j, w = p_i.p_j.
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, int R_ = DICP_RHS_R>
struct AdjQQ {
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D;
    static constexpr int NACC_COL = NACC;
    static constexpr int NSCAL = 0;
    DICP_RHS_COMMON(4 * D)
    struct Row { float q[D], p[D], a[D], u[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qpau<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            r.q[k] = (P.q[o] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[o];
            r.a[k] = P.a[o];
            r.u[k] = P.u[o];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        V z[D], du[D], dp[D];
        V r2, w, apa, duz, dpz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            du[k] = vsub(vbc<V>(r.u[k]), c[3 * D + k]);
            if (DIV) dp[k] = vsub(vbc<V>(r.p[k]), c[D + k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            w = k == 0 ? vmul(vbc<V>(r.p[k]), c[D + k]) : vfma(vbc<V>(r.p[k]), c[D + k], w);
            // (a_i . p_j) + (p_i . a_j)
            apa = k == 0 ? vmul(vbc<V>(r.a[k]), c[D + k]) : vfma(vbc<V>(r.a[k]), c[D + k], apa);
            apa = vfma(vbc<V>(r.p[k]), c[2 * D + k], apa);
            duz = k == 0 ? vmul(du[k], z[k]) : vfma(du[k], z[k], duz);
            if (DIV) dpz = k == 0 ? vmul(dp[k], z[k]) : vfma(dp[k], z[k], dpz);
        }
        const V K = vex2n(r2);
        // ncz = -(coefficient of z' in gq) = -[alpha (ap+pa) + s beta w (du.z')] (+ gc s beta (dp.z') if DIV)
        const V swd = vmul(vmul(vbc<V>(-P.s * P.beta), w), duz);
        V ncz = vfma(vbc<V>(-P.alpha), apa, swd);
        if (DIV) ncz = vfma(vbc<V>(r.gc * P.s * P.beta), dpz, ncz);
        const V Kncz = vmul(K, ncz);
        const V Ksw = vmul(K, vmul(vbc<V>(P.s), w));
        const V Kad = vmul(K, vmul(vbc<V>(P.alpha), duz));
        V nKgz, nKgs;
        if (DIV) {
            nKgz = vmul(K, vbc<V>(-r.gc * P.alpha));
            nKgs = vmul(K, vbc<V>(-r.gc * P.s));
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V gp = vfma(K, c[2 * D + k], acc[A_GP + k]);
            gp = vfma(Kad, c[D + k], gp);
            V gq = vfma(Kncz, z[k], acc[A_GQ + k]);
            gq = vfma(Ksw, du[k], gq);
            if (DIV) {
                gp = vfma(nKgz, z[k], gp);
                gq = vfma(nKgs, dp[k], gq);
            }
            acc[A_GP + k] = gp;
            acc[A_GQ + k] = gq;
        }
    }
    // Both orientations of an unordered pair from ONE evaluation (symmetric engine, sym_engine.cuh): `acc` receives the
    // row's term (row m, column n) exactly as `pair` does, `cacc` the column's term (row n, column m).  Under the swap
    // z', du, dp change sign while K, w, (a.p)+(p.a), (du.z'), (dp.z') do not, so the gq increment is antisymmetric and the
    // gp increment only exchanges (a_n, p_n) for (a_m, p_m) and flips the z' term.
    template <class V, bool MASKED = false>
    static DICP_HD void pair_sym(const Params& P, const Row& r, const V* c, V* acc, V* cacc, V km = V()) {
        V z[D], du[D], dp[D];
        V r2, w, apa, duz, dpz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            du[k] = vsub(vbc<V>(r.u[k]), c[3 * D + k]);
            if (DIV) dp[k] = vsub(vbc<V>(r.p[k]), c[D + k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            w = k == 0 ? vmul(vbc<V>(r.p[k]), c[D + k]) : vfma(vbc<V>(r.p[k]), c[D + k], w);
            apa = k == 0 ? vmul(vbc<V>(r.a[k]), c[D + k]) : vfma(vbc<V>(r.a[k]), c[D + k], apa);
            apa = vfma(vbc<V>(r.p[k]), c[2 * D + k], apa);
            duz = k == 0 ? vmul(du[k], z[k]) : vfma(du[k], z[k], duz);
            if (DIV) dpz = k == 0 ? vmul(dp[k], z[k]) : vfma(dp[k], z[k], dpz);
        }
        V K = vex2n(r2);
        if (MASKED) K = vmul(K, km);                          // padded rows / columns of a ragged tail: K = 0
        const V swd = vmul(vmul(vbc<V>(-P.s * P.beta), w), duz);
        V ncz = vfma(vbc<V>(-P.alpha), apa, swd);
        if (DIV) ncz = vfma(vbc<V>(r.gc * P.s * P.beta), dpz, ncz);
        const V Kncz = vmul(K, ncz);
        const V Ksw = vmul(K, vmul(vbc<V>(P.s), w));
        const V Kad = vmul(K, vmul(vbc<V>(P.alpha), duz));
        V nKgz, nKgs;
        if (DIV) {
            nKgz = vmul(K, vbc<V>(-r.gc * P.alpha));
            nKgs = vmul(K, vbc<V>(-r.gc * P.s));
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V inc = vmul(Kncz, z[k]);                         // gq increment of the row; the column gets its negative
            inc = vfma(Ksw, du[k], inc);
            if (DIV) inc = vfma(nKgs, dp[k], inc);
            acc[A_GQ + k] = vadd(acc[A_GQ + k], inc);
            cacc[A_GQ + k] = vsub(cacc[A_GQ + k], inc);
            V gp = vfma(K, c[2 * D + k], acc[A_GP + k]);
            gp = vfma(Kad, c[D + k], gp);
            V gpc = vfma(K, vbc<V>(r.a[k]), cacc[A_GP + k]);
            gpc = vfma(Kad, vbc<V>(r.p[k]), gpc);
            if (DIV) {
                const V gz = vmul(nKgz, z[k]);
                gp = vadd(gp, gz);
                gpc = vsub(gpc, gz);
            }
            acc[A_GP + k] = gp;
            cacc[A_GP + k] = gpc;
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) {
                P.gp[o] += acc[A_GP + k];
                P.gq[o] += acc[A_GQ + k];
            } else {
                P.gp[o] = acc[A_GP + k];
                P.gq[o] = acc[A_GQ + k];
            }
        }
    }
    // column side of a rectangular pass between two super-blocks (sym_engine.cuh): ADDED to the outputs of column point j
    static constexpr int RECT_R = 2;
    static DICP_HD void finish_col(const Params& P, int j, const float* cacc) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = ((size_t)j + (size_t)P.col0) * D + k;
            P.gp[o] += cacc[A_GP + k];
            P.gq[o] += cacc[A_GQ + k];
        }
    }
};

// ------------------------------------------------------------------------------------------------
// adjoint of the (x,q) pass w.r.t. x: rows = (x_k, wx_k), cols = (q, p)      (eta = 0)
//   gx_k = - sum_j K [alpha (wx_k.p_j) + gc s beta (p_j.z')] z' + gc s sum_j K p_j
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, int R_ = DICP_RHS_R>
struct AdjXQx {
    static constexpr int NACC = D, NSCAL = 0;
    DICP_RHS_COMMON(2 * D)
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        V z[D];
        V r2, wp, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.x[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            wp = k == 0 ? vmul(vbc<V>(r.w[k]), c[D + k]) : vfma(vbc<V>(r.w[k]), c[D + k], wp);
            if (DIV) pz = k == 0 ? vmul(c[D + k], z[k]) : vfma(c[D + k], z[k], pz);
        }
        const V K = vex2n(r2);
        V ncz = vmul(vbc<V>(-P.alpha), wp);
        if (DIV) ncz = vfma(vbc<V>(-r.gc * P.s * P.beta), pz, ncz);
        const V Kncz = vmul(K, ncz);
        V Kg;
        if (DIV) Kg = vmul(K, vbc<V>(r.gc * P.s));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V g = vfma(Kncz, z[k], acc[k]);
            if (DIV) g = vfma(Kg, c[D + k], g);
            acc[k] = g;
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) P.gx[o] += acc[k]; else P.gx[o] = acc[k];
        }
    }
};

// ------------------------------------------------------------------------------------------------
// adjoint of the (x,q) pass w.r.t. (q,p): rows = (q_j, p_j), cols = (x_k, wx_k)   (eta = 0)
//   zeta' = kappa (q_j - x_k)
//   gq_j = - sum_k K [alpha (wx_k.p_j) - gc s beta (p_j.zeta')] zeta' - gc s p_j sum_k K
//   gp_j =   sum_k K wx_k - gc alpha sum_k K zeta'
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, int R_ = DICP_RHS_R>
struct AdjXQq {
    static constexpr int A_GP = 0, A_GQ = D, A_S0 = 2 * D;
    static constexpr int NACC = 2 * D + (DIV ? 1 : 0), NSCAL = 0;
    DICP_RHS_COMMON(2 * D)
    struct Row { float q[D], p[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_xw<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        V z[D];
        V r2, wp, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            wp = k == 0 ? vmul(c[D + k], vbc<V>(r.p[k])) : vfma(c[D + k], vbc<V>(r.p[k]), wp);
            if (DIV) pz = k == 0 ? vmul(vbc<V>(r.p[k]), z[k]) : vfma(vbc<V>(r.p[k]), z[k], pz);
        }
        const V K = vex2n(r2);
        V ncz = vmul(vbc<V>(-P.alpha), wp);
        if (DIV) ncz = vfma(vbc<V>(r.gc * P.s * P.beta), pz, ncz);
        const V Kncz = vmul(K, ncz);
        V nKg;
        if (DIV) nKg = vmul(K, vbc<V>(-r.gc * P.alpha));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V gp = vfma(K, c[D + k], acc[A_GP + k]);
            if (DIV) gp = vfma(nKg, z[k], gp);
            acc[A_GP + k] = gp;
            acc[A_GQ + k] = vfma(Kncz, z[k], acc[A_GQ + k]);
        }
        if (DIV) acc[A_S0] = vadd(acc[A_S0], K);
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            float gq = acc[A_GQ + k];
            if (DIV) gq = fmaf(-r.gc * P.s * acc[A_S0], r.p[k], gq);
            if (P.accumulate) {
                P.gp[o] += acc[A_GP + k];
                P.gq[o] += gq;
            } else {
                P.gp[o] = acc[A_GP + k];
                P.gq[o] = gq;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// adjoint of the (x,q) pass, BOTH sides from one evaluation of the pair (rectangular ring engine, sym_engine.cuh):
// rows = (x_k, wx_k), cols = (q_j, p_j), z' = kappa (x_k - q_j) = -zeta'.  With wp = wx_k.p_j, pz = p_j.z',
// c = -alpha wp - gc s beta pz (the SAME coefficient in AdjXQx and AdjXQq, because p_j.zeta' = -pz):
//   row side    gx_k += K c z' + gc s K p_j                      (= AdjXQx)
//   column side gq_j -= K c z' ;  gp_j += K wx_k + gc alpha K z' ;  S0_j += K        (= AdjXQq; finish adds -gc s S0_j p_j)
// 36 operations per pair instead of 22 + 26 in two passes (D = 3, with the divergence cost).
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV>
struct AdjXQ {
    using Params = RhsParams;
    static constexpr bool PACKED = true;
    static constexpr int NF = 2 * D, COLF4 = (NF + 3) / 4;
    static constexpr int NACC = D;                                   // row side: gx
    static constexpr int C_GP = 0, C_GQ = D, C_S0 = 2 * D;
    static constexpr int NACC_COL = 2 * D + (DIV ? 1 : 0);           // column side: gp, gq, S0
    static constexpr int NSCAL = 0;
    static constexpr int RECT_R = 4;                                 // rows per lane in the ring engine (light rows)
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    template <class V, bool MASKED = false>
    static DICP_HD void pair_sym(const Params& P, const Row& r, const V* c, V* acc, V* cacc, V km = V()) {
        V z[D];
        V r2, wp, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.x[k]), c[k]);
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
            wp = k == 0 ? vmul(vbc<V>(r.w[k]), c[D + k]) : vfma(vbc<V>(r.w[k]), c[D + k], wp);
            if (DIV) pz = k == 0 ? vmul(c[D + k], z[k]) : vfma(c[D + k], z[k], pz);
        }
        V K = vex2n(r2);
        if (MASKED) K = vmul(K, km);
        V ncz = vmul(vbc<V>(-P.alpha), wp);
        if (DIV) ncz = vfma(vbc<V>(-r.gc * P.s * P.beta), pz, ncz);
        const V Kncz = vmul(K, ncz);
        V Kg, Kga;
        if (DIV) {
            Kg = vmul(K, vbc<V>(r.gc * P.s));
            Kga = vmul(K, vbc<V>(r.gc * P.alpha));
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const V inc = vmul(Kncz, z[k]);
            V gx = vadd(acc[k], inc);
            V gp = vfma(K, vbc<V>(r.w[k]), cacc[C_GP + k]);
            if (DIV) {
                gx = vfma(Kg, c[D + k], gx);
                gp = vfma(Kga, z[k], gp);
            }
            acc[k] = gx;
            cacc[C_GP + k] = gp;
            cacc[C_GQ + k] = vsub(cacc[C_GQ + k], inc);
        }
        if (DIV) cacc[C_S0] = vadd(cacc[C_S0], K);
    }
    // row side: gx_k
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) P.gx[(size_t)i * D + k] = acc[k];
    }
    // column side: ADDED to gq_j, gp_j (which hold the (q,q) pass)
    static DICP_HD void finish_col(const Params& P, int j, const float* cacc) {
        const float gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)j * D + k;
            float gq = cacc[C_GQ + k];
            if (DIV) gq = fmaf(-gc * P.s * cacc[C_S0], P.p[o], gq);
            P.gp[o] += cacc[C_GP + k];
            P.gq[o] += gq;
        }
    }
};

// ================================================================================================================
// Adjoint of the logdet model (eta = 1/lambda != 0).  Notation per pair (row m, column n), scaled z' = kappa (row - col):
//   w = p_m.p_n, e = p_m - p_n, du = u_m - u_n, dA = a_m - a_n, t0 = beta r'^2 - D, t1 = beta r'^2 - (D+2), g = gc
//   Phi  = (a_m.p_n + a_n.p_m) + eta alpha (dA.z') + alpha w (du.z') + eta s beta (z'.e)(du.z') - eta s (du.e)
//          - eta^2 s alpha t1 (du.z') - g alpha (e.z') + 2 g eta s t0
//   gq_m = sum_n K { eta s dA + [s w + eta s alpha (z'.e) - eta^2 s^2 t1] du + [eta s alpha (du.z') - g s] e
//                    + [-2 eta^2 s^2 beta (du.z') + 4 g eta s alpha - alpha Phi] z' }
//   gp_m = sum_n K { a_n + alpha (du.z') p_n + eta s beta (du.z') z' - eta s du - g alpha z' }
// ================================================================================================================
template <int D, int R_ = DICP_RHS_R>
struct AdjQQEta {
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D, NACC_COL = NACC, NSCAL = 0;
    DICP_RHS_COMMON(4 * D)
    struct Row { float q[D], p[D], a[D], u[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qpau<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            r.q[k] = (P.q[o] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[o];
            r.a[k] = P.a[o];
            r.u[k] = P.u[o];
        }
        r.gc = (P.gc != nullptr && P.x == nullptr) ? P.gc[0] : 0.f;   // dcost comes from the (q,q) pass only when x is absent
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        const float eta = P.eta, s = P.s, al = P.alpha, be = P.beta, g = r.gc, es = eta * s;
        V z[D], e[D], du[D], dA[D];
        V r2, w, ze, duz, dAz, due, appa;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            e[k] = vsub(vbc<V>(r.p[k]), c[D + k]);
            dA[k] = vsub(vbc<V>(r.a[k]), c[2 * D + k]);
            du[k] = vsub(vbc<V>(r.u[k]), c[3 * D + k]);
            if (k == 0) {
                r2 = vmul(z[k], z[k]);
                w = vmul(vbc<V>(r.p[k]), c[D + k]);
                ze = vmul(z[k], e[k]);
                duz = vmul(du[k], z[k]);
                dAz = vmul(dA[k], z[k]);
                due = vmul(du[k], e[k]);
                appa = vmul(vbc<V>(r.p[k]), c[2 * D + k]);
            } else {
                r2 = vfma(z[k], z[k], r2);
                w = vfma(vbc<V>(r.p[k]), c[D + k], w);
                ze = vfma(z[k], e[k], ze);
                duz = vfma(du[k], z[k], duz);
                dAz = vfma(dA[k], z[k], dAz);
                due = vfma(du[k], e[k], due);
                appa = vfma(vbc<V>(r.p[k]), c[2 * D + k], appa);
            }
            appa = vfma(vbc<V>(r.a[k]), c[D + k], appa);
        }
        const V K = vex2n(r2);
        const V t0 = vfma(vbc<V>(be), r2, vbc<V>(-(float)D));
        const V t1 = vfma(vbc<V>(be), r2, vbc<V>(-(float)(D + 2)));
        // c_du = s w + eta s alpha (z'.e) - (eta s)^2 t1
        const V c_du = vfma(vbc<V>(s), w, vfma(vbc<V>(es * al), ze, vmul(vbc<V>(-es * es), t1)));
        // c_e = eta s alpha (du.z') - g s
        const V c_e = vfma(vbc<V>(es * al), duz, vbc<V>(-g * s));
        // Phi
        V Phi = vfma(vbc<V>(eta * al), dAz, appa);
        Phi = vfma(vmul(vbc<V>(al), w), duz, Phi);
        Phi = vfma(vmul(vbc<V>(es * be), ze), duz, Phi);
        Phi = vfma(vbc<V>(-es), due, Phi);
        Phi = vfma(vmul(vbc<V>(-eta * es * al), t1), duz, Phi);
        Phi = vfma(vbc<V>(-g * al), ze, Phi);
        Phi = vfma(vbc<V>(2.f * g * es), t0, Phi);
        // Cz = -2 (eta s)^2 beta (du.z') + 4 g eta s alpha - alpha Phi
        const V Cz = vfma(vbc<V>(-al), Phi, vfma(vbc<V>(-2.f * es * es * be), duz, vbc<V>(4.f * g * es * al)));
        const V pz = vfma(vbc<V>(es * be), duz, vbc<V>(-g * al));       // coefficient of z' in gp
        const V pp = vmul(vbc<V>(al), duz);                            // coefficient of p_n in gp
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V gq = vmul(vbc<V>(es), dA[k]);
            gq = vfma(c_du, du[k], gq);
            gq = vfma(c_e, e[k], gq);
            gq = vfma(Cz, z[k], gq);
            V gp = vfma(pp, c[D + k], c[2 * D + k]);
            gp = vfma(pz, z[k], gp);
            gp = vfma(vbc<V>(-es), du[k], gp);
            acc[A_GQ + k] = vfma(K, gq, acc[A_GQ + k]);
            acc[A_GP + k] = vfma(K, gp, acc[A_GP + k]);
        }
    }
    // Both orientations of an unordered pair from ONE evaluation (symmetric engine).  Under the swap m <-> n the vectors
    // z', e, du, dA change sign and every scalar (w, z'.e, du.z', dA.z', du.e, t0, t1, Phi) is unchanged: the gq term is
    // antisymmetric; the gp term is  a_m + alpha (du.z') p_m - [eta s beta (du.z') z' - eta s du - g alpha z'].
    template <class V, bool MASKED = false>
    static DICP_HD void pair_sym(const Params& P, const Row& r, const V* c, V* acc, V* cacc, V km = V()) {
        const float eta = P.eta, s = P.s, al = P.alpha, be = P.beta, g = r.gc, es = eta * s;
        V z[D], e[D], du[D], dA[D];
        V r2, w, ze, duz, dAz, due, appa;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            e[k] = vsub(vbc<V>(r.p[k]), c[D + k]);
            dA[k] = vsub(vbc<V>(r.a[k]), c[2 * D + k]);
            du[k] = vsub(vbc<V>(r.u[k]), c[3 * D + k]);
            if (k == 0) {
                r2 = vmul(z[k], z[k]);
                w = vmul(vbc<V>(r.p[k]), c[D + k]);
                ze = vmul(z[k], e[k]);
                duz = vmul(du[k], z[k]);
                dAz = vmul(dA[k], z[k]);
                due = vmul(du[k], e[k]);
                appa = vmul(vbc<V>(r.p[k]), c[2 * D + k]);
            } else {
                r2 = vfma(z[k], z[k], r2);
                w = vfma(vbc<V>(r.p[k]), c[D + k], w);
                ze = vfma(z[k], e[k], ze);
                duz = vfma(du[k], z[k], duz);
                dAz = vfma(dA[k], z[k], dAz);
                due = vfma(du[k], e[k], due);
                appa = vfma(vbc<V>(r.p[k]), c[2 * D + k], appa);
            }
            appa = vfma(vbc<V>(r.a[k]), c[D + k], appa);
        }
        V K = vex2n(r2);
        if (MASKED) K = vmul(K, km);                          // padded rows / columns of a ragged tail: K = 0
        const V nK = vmul(K, vbc<V>(-1.f));
        const V t0 = vfma(vbc<V>(be), r2, vbc<V>(-(float)D));
        const V t1 = vfma(vbc<V>(be), r2, vbc<V>(-(float)(D + 2)));
        const V c_du = vfma(vbc<V>(s), w, vfma(vbc<V>(es * al), ze, vmul(vbc<V>(-es * es), t1)));
        const V c_e = vfma(vbc<V>(es * al), duz, vbc<V>(-g * s));
        V Phi = vfma(vbc<V>(eta * al), dAz, appa);
        Phi = vfma(vmul(vbc<V>(al), w), duz, Phi);
        Phi = vfma(vmul(vbc<V>(es * be), ze), duz, Phi);
        Phi = vfma(vbc<V>(-es), due, Phi);
        Phi = vfma(vmul(vbc<V>(-eta * es * al), t1), duz, Phi);
        Phi = vfma(vbc<V>(-g * al), ze, Phi);
        Phi = vfma(vbc<V>(2.f * g * es), t0, Phi);
        const V Cz = vfma(vbc<V>(-al), Phi, vfma(vbc<V>(-2.f * es * es * be), duz, vbc<V>(4.f * g * es * al)));
        const V pz = vfma(vbc<V>(es * be), duz, vbc<V>(-g * al));
        const V pp = vmul(vbc<V>(al), duz);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V gq = vmul(vbc<V>(es), dA[k]);
            gq = vfma(c_du, du[k], gq);
            gq = vfma(c_e, e[k], gq);
            gq = vfma(Cz, z[k], gq);
            acc[A_GQ + k] = vfma(K, gq, acc[A_GQ + k]);
            cacc[A_GQ + k] = vfma(nK, gq, cacc[A_GQ + k]);
            const V t = vfma(pz, z[k], vmul(vbc<V>(-es), du[k]));            // odd part of the gp term
            const V gpr = vadd(vfma(pp, c[D + k], c[2 * D + k]), t);
            const V gpc = vsub(vfma(pp, vbc<V>(r.p[k]), vbc<V>(r.a[k])), t);
            acc[A_GP + k] = vfma(K, gpr, acc[A_GP + k]);
            cacc[A_GP + k] = vfma(K, gpc, cacc[A_GP + k]);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) {
                P.gp[o] += acc[A_GP + k];
                P.gq[o] += acc[A_GQ + k];
            } else {
                P.gp[o] = acc[A_GP + k];
                P.gq[o] = acc[A_GQ + k];
            }
        }
    }
    // column side of a rectangular pass between two super-blocks (sym_engine.cuh): ADDED to the outputs of column point j
    static constexpr int RECT_R = 2;
    static DICP_HD void finish_col(const Params& P, int j, const float* cacc) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = ((size_t)j + (size_t)P.col0) * D + k;
            P.gp[o] += cacc[A_GP + k];
            P.gq[o] += cacc[A_GQ + k];
        }
    }
};

// logdet, (x,q) pass w.r.t. x: rows = (x_k, wx_k), cols = (q,p)
//   psi  = wx.p + eta alpha (wx.z') + g [alpha (p.z') + eta s t0]
//   gx_k = sum_j K { eta s wx_k + g s p_j + (2 g eta s alpha - alpha psi) z' }
template <int D, int R_ = DICP_RHS_R>
struct AdjXQxEta {
    static constexpr int NACC = D, NSCAL = 0;
    DICP_RHS_COMMON(2 * D)
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = P.gc != nullptr ? P.gc[0] : 0.f;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        const float es = P.eta * P.s, al = P.alpha, g = r.gc;
        V z[D];
        V r2, wp, wz, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.x[k]), c[k]);
            if (k == 0) {
                r2 = vmul(z[k], z[k]);
                wp = vmul(vbc<V>(r.w[k]), c[D + k]);
                wz = vmul(vbc<V>(r.w[k]), z[k]);
                pz = vmul(c[D + k], z[k]);
            } else {
                r2 = vfma(z[k], z[k], r2);
                wp = vfma(vbc<V>(r.w[k]), c[D + k], wp);
                wz = vfma(vbc<V>(r.w[k]), z[k], wz);
                pz = vfma(c[D + k], z[k], pz);
            }
        }
        const V K = vex2n(r2);
        // psi = wp + eta al wz + g al pz + g es (beta r2 - D)
        V psi = vfma(vbc<V>(P.eta * al), wz, wp);
        psi = vfma(vbc<V>(g * al), pz, psi);
        psi = vfma(vbc<V>(g * es * P.beta), r2, vadd(psi, vbc<V>(-g * es * (float)D)));
        const V cz = vfma(vbc<V>(-al), psi, vbc<V>(2.f * g * es * al));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V t = vfma(vbc<V>(g * P.s), c[D + k], vbc<V>(es * r.w[k]));
            t = vfma(cz, z[k], t);
            acc[k] = vfma(K, t, acc[k]);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) P.gx[o] += acc[k]; else P.gx[o] = acc[k];
        }
    }
};

// logdet, (x,q) pass, BOTH sides from one evaluation of the pair (rectangular ring engine): rows = (x_k, wx_k), cols = (q_j, p_j).
// psi is the same in AdjXQxEta and AdjXQqEta (zeta' = -z'), so with t = eta s wx_k + g s p_j + (2 g eta s alpha - alpha psi) z':
//   row side    gx_k += K t
//   column side gq_j -= K t ;  gp_j += K (wx_k + g alpha z')
template <int D>
struct AdjXQE {
    using Params = RhsParams;
    static constexpr bool PACKED = true;
    static constexpr int NF = 2 * D, COLF4 = (NF + 3) / 4;
    static constexpr int NACC = D;
    static constexpr int C_GP = 0, C_GQ = D;
    static constexpr int NACC_COL = 2 * D;
    static constexpr int NSCAL = 0;
    static constexpr int RECT_R = 4;
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_qp<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = P.gc != nullptr ? P.gc[0] : 0.f;
    }
    template <class V, bool MASKED = false>
    static DICP_HD void pair_sym(const Params& P, const Row& r, const V* c, V* acc, V* cacc, V km = V()) {
        const float es = P.eta * P.s, al = P.alpha, g = r.gc;
        V z[D];
        V r2, wp, wz, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.x[k]), c[k]);
            if (k == 0) {
                r2 = vmul(z[k], z[k]);
                wp = vmul(vbc<V>(r.w[k]), c[D + k]);
                wz = vmul(vbc<V>(r.w[k]), z[k]);
                pz = vmul(c[D + k], z[k]);
            } else {
                r2 = vfma(z[k], z[k], r2);
                wp = vfma(vbc<V>(r.w[k]), c[D + k], wp);
                wz = vfma(vbc<V>(r.w[k]), z[k], wz);
                pz = vfma(c[D + k], z[k], pz);
            }
        }
        V K = vex2n(r2);
        if (MASKED) K = vmul(K, km);
        V psi = vfma(vbc<V>(P.eta * al), wz, wp);
        psi = vfma(vbc<V>(g * al), pz, psi);
        psi = vfma(vbc<V>(g * es * P.beta), r2, vadd(psi, vbc<V>(-g * es * (float)D)));
        const V cz = vfma(vbc<V>(-al), psi, vbc<V>(2.f * g * es * al));
        const V Kga = vmul(K, vbc<V>(g * al));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V t = vfma(vbc<V>(g * P.s), c[D + k], vbc<V>(es * r.w[k]));
            t = vfma(cz, z[k], t);
            const V Kt = vmul(K, t);
            acc[k] = vadd(acc[k], Kt);
            cacc[C_GQ + k] = vsub(cacc[C_GQ + k], Kt);
            V gp = vfma(K, vbc<V>(r.w[k]), cacc[C_GP + k]);
            cacc[C_GP + k] = vfma(Kga, z[k], gp);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) P.gx[(size_t)i * D + k] = acc[k];
    }
    static DICP_HD void finish_col(const Params& P, int j, const float* cacc) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)j * D + k;
            P.gp[o] += cacc[C_GP + k];
            P.gq[o] += cacc[C_GQ + k];
        }
    }
};

// logdet, (x,q) pass w.r.t. (q,p): rows = (q_j, p_j), cols = (x_k, wx_k); zeta' = kappa (q_j - x_k)
//   psi  = wx.p - eta alpha (wx.zeta') + g [-alpha (p.zeta') + eta s t0]
//   gq_j = sum_k K { -eta s wx_k - g s p_j + (2 g eta s alpha - alpha psi) zeta' }
//   gp_j = sum_k K { wx_k - g alpha zeta' }
template <int D, int R_ = DICP_RHS_R>
struct AdjXQqEta {
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D, NSCAL = 0;
    DICP_RHS_COMMON(2 * D)
    struct Row { float q[D], p[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { pack_xw<D>(P, j, N, c, COLF4 * 4); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
        r.gc = P.gc != nullptr ? P.gc[0] : 0.f;
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* acc) {
        const float es = P.eta * P.s, al = P.alpha, g = r.gc;
        V z[D];
        V r2, wp, wz, pz;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(vbc<V>(r.q[k]), c[k]);
            if (k == 0) {
                r2 = vmul(z[k], z[k]);
                wp = vmul(c[D + k], vbc<V>(r.p[k]));
                wz = vmul(c[D + k], z[k]);
                pz = vmul(vbc<V>(r.p[k]), z[k]);
            } else {
                r2 = vfma(z[k], z[k], r2);
                wp = vfma(c[D + k], vbc<V>(r.p[k]), wp);
                wz = vfma(c[D + k], z[k], wz);
                pz = vfma(vbc<V>(r.p[k]), z[k], pz);
            }
        }
        const V K = vex2n(r2);
        V psi = vfma(vbc<V>(-P.eta * al), wz, wp);
        psi = vfma(vbc<V>(-g * al), pz, psi);
        psi = vfma(vbc<V>(g * es * P.beta), r2, vadd(psi, vbc<V>(-g * es * (float)D)));
        const V cz = vfma(vbc<V>(-al), psi, vbc<V>(2.f * g * es * al));
#pragma unroll
        for (int k = 0; k < D; ++k) {
            V tq = vfma(vbc<V>(-es), c[D + k], vbc<V>(-g * P.s * r.p[k]));
            tq = vfma(cz, z[k], tq);
            acc[A_GQ + k] = vfma(K, tq, acc[A_GQ + k]);
            const V tp = vfma(vbc<V>(-g * al), z[k], c[D + k]);
            acc[A_GP + k] = vfma(K, tp, acc[A_GP + k]);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) {
                P.gp[o] += acc[A_GP + k];
                P.gq[o] += acc[A_GQ + k];
            } else {
                P.gp[o] = acc[A_GP + k];
                P.gq[o] = acc[A_GQ + k];
            }
        }
    }
};

}  // namespace dicp
