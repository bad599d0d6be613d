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
//
// Adjoint (eta = 0 models; cotangents a of vq, u of dp, wx of vx, gc of dcost, gh of A [= lambda/2 * dL/dH ... see capi]):
//   derived in DESIGN.md §5; checked against torch autograd of the oracle in tests/test_host_emulation.py.
#pragma once
#include "pair_engine.cuh"

// tuning knobs (overridable at compile time for sweeps: -DDICP_RHS_R=4 ...)
#ifndef DICP_RHS_R
#define DICP_RHS_R 2
#endif
#ifndef DICP_RHS_THREADS
#define DICP_RHS_THREADS 128
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
};

// ------------------------------------------------------------------------------------------------
// forward, (q,q)
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, bool ETA, int R_ = DICP_RHS_R>
struct RhsQQ {
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int A_V = 0, A_T = D, A_Z = 2 * D;
    static constexpr bool NEEDZ = DIV || ETA;
    static constexpr int A_S0 = A_Z + (NEEDZ ? D : 0);
    static constexpr int A_R2 = A_S0 + (ETA ? 1 : 0);
    static constexpr int NACC = A_R2 + (ETA ? 1 : 0);
    static constexpr int NSCAL = 3;   // A, B, C
    struct Row { float q[D], p[D]; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                c[k] = (P.q[(size_t)j * D + k] - P.origin[k]) * P.kappa;
                c[D + k] = P.p[(size_t)j * D + k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
        }
    }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* a) {
        float z[D];
        float r2 = 0.f, w = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.q[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            w = fmaf(r.p[k], c[D + k], w);
        }
        const float K = ex2_neg(r2);
        float coef;
        if (ETA) {
            float ze = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) ze = fmaf(z[k], r.p[k] - c[D + k], ze);
            // alpha w + eta s beta (z'.e) - eta^2 s alpha (beta r'^2 - (D+2))
            const float c1 = P.eta * P.s * P.beta, c2 = P.eta * P.eta * P.s * P.alpha;
            coef = fmaf(P.alpha, w, fmaf(c1, ze, -c2 * fmaf(P.beta, r2, -(float)(D + 2))));
            a[A_S0] += K;
            a[A_R2] = fmaf(K, r2, a[A_R2]);
        } else {
            coef = w;   // alpha applied in finish
        }
        const float Kc = K * coef;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[A_V + k] = fmaf(K, c[D + k], a[A_V + k]);
            a[A_T + k] = fmaf(Kc, z[k], a[A_T + k]);
            if (NEEDZ) a[A_Z + k] = fmaf(K, z[k], a[A_Z + k]);
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
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int A_V = 0, A_DS = D;
    static constexpr int A_Z = A_DS + (DIV ? 1 : 0);
    static constexpr int A_S0 = A_Z + (ETA ? D : 0);
    static constexpr int A_R2 = A_S0 + (ETA ? 1 : 0);
    static constexpr int NACC = A_R2 + (ETA ? 1 : 0);
    static constexpr int NSCAL = 1;   // dcost contribution
    struct Row { float x[D]; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { RhsQQ<D, DIV, ETA, R_>::pack_col(P, j, N, c); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* a) {
        float z[D];
        float r2 = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.x[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            if (DIV) pz = fmaf(c[D + k], z[k], pz);
        }
        const float K = ex2_neg(r2);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            a[A_V + k] = fmaf(K, c[D + k], a[A_V + k]);
            if (ETA) a[A_Z + k] = fmaf(K, z[k], a[A_Z + k]);
        }
        if (DIV) a[A_DS] = fmaf(K, pz, a[A_DS]);
        if (ETA) {
            a[A_S0] += K;
            a[A_R2] = fmaf(K, r2, a[A_R2]);
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
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (4 * D + 3) / 4;
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D;
    static constexpr int NSCAL = 0;
    struct Row { float q[D], p[D], a[D], u[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const size_t o = (size_t)j * D + k;
                c[k] = (P.q[o] - P.origin[k]) * P.kappa;
                c[D + k] = P.p[o];
                c[2 * D + k] = P.a[o];
                c[3 * D + k] = P.u[o];
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
        }
    }
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
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        float z[D], du[D];
        float r2 = 0.f, w = 0.f, ap = 0.f, pa = 0.f, duz = 0.f, dpz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.q[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            w = fmaf(r.p[k], c[D + k], w);
            ap = fmaf(r.a[k], c[D + k], ap);          // a_i . p_j
            pa = fmaf(r.p[k], c[2 * D + k], pa);      // p_i . a_j
            du[k] = r.u[k] - c[3 * D + k];
            duz = fmaf(du[k], z[k], duz);
            if (DIV) dpz = fmaf(r.p[k] - c[D + k], z[k], dpz);
        }
        const float K = ex2_neg(r2);
        // coefficient of z' in gq (sign folded: gq -= cz * z')
        float cz = fmaf(P.alpha, ap + pa, P.s * P.beta * w * duz);
        if (DIV) cz = fmaf(-r.gc * P.s * P.beta, dpz, cz);
        const float Kcz = K * cz;
        const float Ksw = K * P.s * w;
        const float Kad = K * P.alpha * duz;
        const float Kgz = DIV ? K * r.gc * P.alpha : 0.f;
        const float Kgs = DIV ? K * r.gc * P.s : 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float gp = fmaf(K, c[2 * D + k], acc[A_GP + k]);
            gp = fmaf(Kad, c[D + k], gp);
            float gq = fmaf(-Kcz, z[k], acc[A_GQ + k]);
            gq = fmaf(Ksw, du[k], gq);
            if (DIV) {
                gp = fmaf(-Kgz, z[k], gp);
                gq = fmaf(-Kgs, r.p[k] - c[D + k], gq);
            }
            acc[A_GP + k] = gp;
            acc[A_GQ + k] = gq;
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
};

// ------------------------------------------------------------------------------------------------
// adjoint of the (x,q) pass w.r.t. x: rows = (x_k, wx_k), cols = (q, p)      (eta = 0)
//   gx_k = - sum_j K [alpha (wx_k.p_j) + gc s beta (p_j.z')] z' + gc s sum_j K p_j
// ------------------------------------------------------------------------------------------------
template <int D, bool DIV, int R_ = DICP_RHS_R>
struct AdjXQx {
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int NACC = D, NSCAL = 0;
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { RhsQQ<D, DIV, false, R_>::pack_col(P, j, N, c); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        float z[D];
        float r2 = 0.f, wp = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.x[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            wp = fmaf(r.w[k], c[D + k], wp);
            if (DIV) pz = fmaf(c[D + k], z[k], pz);
        }
        const float K = ex2_neg(r2);
        float cz = P.alpha * wp;
        if (DIV) cz = fmaf(r.gc * P.s * P.beta, pz, cz);
        const float Kcz = K * cz;
        const float Kg = DIV ? K * r.gc * P.s : 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float g = fmaf(-Kcz, z[k], acc[k]);
            if (DIV) g = fmaf(Kg, c[D + k], g);
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
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int A_GP = 0, A_GQ = D, A_S0 = 2 * D;
    static constexpr int NACC = 2 * D + (DIV ? 1 : 0), NSCAL = 0;
    struct Row { float q[D], p[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                c[k] = (P.x[(size_t)j * D + k] - P.origin[k]) * P.kappa;
                c[D + k] = P.wx[(size_t)j * D + k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
        }
    }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
        r.gc = (DIV && P.gc != nullptr) ? P.gc[0] : 0.f;
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        float z[D];
        float r2 = 0.f, wp = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.q[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            wp = fmaf(c[D + k], r.p[k], wp);
            if (DIV) pz = fmaf(r.p[k], z[k], pz);
        }
        const float K = ex2_neg(r2);
        float cz = P.alpha * wp;
        if (DIV) cz = fmaf(-r.gc * P.s * P.beta, pz, cz);
        const float Kcz = K * cz;
        const float Kg = DIV ? K * r.gc * P.alpha : 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float gp = fmaf(K, c[D + k], acc[A_GP + k]);
            if (DIV) gp = fmaf(-Kg, z[k], gp);
            acc[A_GP + k] = gp;
            acc[A_GQ + k] = fmaf(-Kcz, z[k], acc[A_GQ + k]);
        }
        if (DIV) acc[A_S0] += K;
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


// ================================================================================================================
// Adjoint of the logdet model (eta = 1/lambda != 0).  Derivation in DESIGN.md §5.3; checked against torch autograd of
// the oracle (tests/test_host_emulation.py).  Notation per pair (row m, column n), scaled z' = kappa (row - col):
//   w = p_m.p_n, e = p_m - p_n, du = u_m - u_n, dA = a_m - a_n, t0 = beta r'^2 - D, t1 = beta r'^2 - (D+2), g = gc
//   Phi  = (a_m.p_n + a_n.p_m) + eta alpha (dA.z') + alpha w (du.z') + eta s beta (z'.e)(du.z') - eta s (du.e)
//          - eta^2 s alpha t1 (du.z') - g alpha (e.z') + 2 g eta s t0
//   gq_m = sum_n K { eta s dA + [s w + eta s alpha (z'.e) - eta^2 s^2 t1] du + [eta s alpha (du.z') - g s] e
//                    + [-2 eta^2 s^2 beta (du.z') + 4 g eta s alpha - alpha Phi] z' }
//   gp_m = sum_n K { a_n + alpha (du.z') p_n + eta s beta (du.z') z' - eta s du - g alpha z' }
// ================================================================================================================
template <int D, int R_ = DICP_RHS_R>
struct AdjQQEta {
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (4 * D + 3) / 4;
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D, NSCAL = 0;
    struct Row { float q[D], p[D], a[D], u[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { AdjQQ<D, true, R_>::pack_col(P, j, N, c); }
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
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        const float eta = P.eta, s = P.s, al = P.alpha, be = P.beta, g = r.gc;
        float z[D], e[D], du[D], dA[D];
        float r2 = 0.f, w = 0.f, ze = 0.f, duz = 0.f, dAz = 0.f, due = 0.f, appa = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.q[k] - c[k];
            e[k] = r.p[k] - c[D + k];
            dA[k] = r.a[k] - c[2 * D + k];
            du[k] = r.u[k] - c[3 * D + k];
            r2 = fmaf(z[k], z[k], r2);
            w = fmaf(r.p[k], c[D + k], w);
            ze = fmaf(z[k], e[k], ze);
            duz = fmaf(du[k], z[k], duz);
            dAz = fmaf(dA[k], z[k], dAz);
            due = fmaf(du[k], e[k], due);
            appa = fmaf(r.a[k], c[D + k], fmaf(r.p[k], c[2 * D + k], appa));
        }
        const float K = ex2_neg(r2);
        const float t0 = fmaf(be, r2, -(float)D), t1 = fmaf(be, r2, -(float)(D + 2));
        const float es = eta * s;
        const float c_du = s * w + es * al * ze - es * es * t1;
        const float c_e = es * al * duz - g * s;
        const float Phi = appa + eta * al * dAz + al * w * duz + es * be * ze * duz - es * due - eta * es * al * t1 * duz
                          - g * al * ze + 2.f * g * es * t0;
        const float Cz = -2.f * es * es * be * duz + 4.f * g * es * al - al * Phi;
        const float pz = es * be * duz - g * al;           // coefficient of z' in gp
        const float pp = al * duz;                         // coefficient of p_n in gp
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float gq = es * dA[k] + c_du * du[k] + c_e * e[k] + Cz * z[k];
            const float gp = c[2 * D + k] + pp * c[D + k] + pz * z[k] - es * du[k];
            acc[A_GQ + k] = fmaf(K, gq, acc[A_GQ + k]);
            acc[A_GP + k] = fmaf(K, gp, acc[A_GP + k]);
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
};

// logdet, (x,q) pass w.r.t. x: rows = (x_k, wx_k), cols = (q,p)
//   psi  = wx.p + eta alpha (wx.z') + g [alpha (p.z') + eta s t0]
//   gx_k = sum_j K { eta s wx_k + g s p_j + (2 g eta s alpha - alpha psi) z' }
template <int D, int R_ = DICP_RHS_R>
struct AdjXQxEta {
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int NACC = D, NSCAL = 0;
    struct Row { float x[D], w[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { RhsQQ<D, true, true, R_>::pack_col(P, j, N, c); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.x[k] = (P.x[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.w[k] = P.wx[(size_t)i * D + k];
        }
        r.gc = P.gc != nullptr ? P.gc[0] : 0.f;
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        const float es = P.eta * P.s, al = P.alpha, g = r.gc;
        float z[D];
        float r2 = 0.f, wp = 0.f, wz = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.x[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            wp = fmaf(r.w[k], c[D + k], wp);
            wz = fmaf(r.w[k], z[k], wz);
            pz = fmaf(c[D + k], z[k], pz);
        }
        const float K = ex2_neg(r2);
        const float psi = wp + P.eta * al * wz + g * (al * pz + es * fmaf(P.beta, r2, -(float)D));
        const float cz = 2.f * g * es * al - al * psi;
#pragma unroll
        for (int k = 0; k < D; ++k)
            acc[k] = fmaf(K, es * r.w[k] + g * P.s * c[D + k] + cz * z[k], acc[k]);
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* acc, float*) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const size_t o = (size_t)i * D + k;
            if (P.accumulate) P.gx[o] += acc[k]; else P.gx[o] = acc[k];
        }
    }
};

// logdet, (x,q) pass w.r.t. (q,p): rows = (q_j, p_j), cols = (x_k, wx_k); zeta' = kappa (q_j - x_k)
//   psi  = wx.p - eta alpha (wx.zeta') + g [-alpha (p.zeta') + eta s t0]
//   gq_j = sum_k K { -eta s wx_k - g s p_j + (2 g eta s alpha - alpha psi) zeta' }
//   gp_j = sum_k K { wx_k - g alpha zeta' }
template <int D, int R_ = DICP_RHS_R>
struct AdjXQqEta {
    using Params = RhsParams;
    static constexpr int THREADS = DICP_RHS_THREADS, MINB = DICP_RHS_MINB, R = R_, TILE = DICP_RHS_TILE;
    static constexpr int COLF4 = (2 * D + 3) / 4;
    static constexpr int A_GP = 0, A_GQ = D;
    static constexpr int NACC = 2 * D, NSCAL = 0;
    struct Row { float q[D], p[D], gc; };

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) { AdjXQq<D, true, R_>::pack_col(P, j, N, c); }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.q[k] = (P.q[(size_t)i * D + k] - P.origin[k]) * P.kappa;
            r.p[k] = P.p[(size_t)i * D + k];
        }
        r.gc = P.gc != nullptr ? P.gc[0] : 0.f;
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
    }
    static DICP_HD void combine(float* a, const float* b) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] += b[k];
    }
    static DICP_HD void pair(const Params& P, const Row& r, const float* c, float* acc) {
        const float es = P.eta * P.s, al = P.alpha, g = r.gc;
        float z[D];
        float r2 = 0.f, wp = 0.f, wz = 0.f, pz = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = r.q[k] - c[k];
            r2 = fmaf(z[k], z[k], r2);
            wp = fmaf(c[D + k], r.p[k], wp);
            wz = fmaf(c[D + k], z[k], wz);
            pz = fmaf(r.p[k], z[k], pz);
        }
        const float K = ex2_neg(r2);
        const float psi = wp - P.eta * al * wz + g * (-al * pz + es * fmaf(P.beta, r2, -(float)D));
        const float cz = 2.f * g * es * al - al * psi;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            acc[A_GQ + k] = fmaf(K, -es * c[D + k] - g * P.s * r.p[k] + cz * z[k], acc[A_GQ + k]);
            acc[A_GP + k] = fmaf(K, c[D + k] - g * al * z[k], acc[A_GP + k]);
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
