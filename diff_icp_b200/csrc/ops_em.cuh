// GMM EM step (isotropic, one shared sigma) as pair-engine Ops.
//
// Replaces the six KeOps reductions of GaussianMixtureUnif.EM_step_keops
// (/root/reference/diffICP/core/GMM.py:402-496: row LSE :410-415, sumsoftmaxweight :443, column logsumexp :446,
// weighted logsumexp :455, targets :473, free-energy offset :485-486) / the dense (N,C) block of EM_step_torch
// (:263-317) by at most three sweeps over points x components, each with one exponential per pair:
//
//   EmRow<D, LITE=true>   row pass, T_n = LSE_c t_nc only                       (needed before the column pass)
//   EmCol<D>              column pass: log-domain sufficient statistics of every component
//                         (m_c, S0_c, B_c, A_c) = online-softmax sums over n of gamma_nc, gamma_nc (x_n - mu_c),
//                         gamma_nc |x_n - mu_c|^2, centred on the OLD mu (well conditioned)
//   EmRow<D, LITE=false>  row pass with payload: responsibilities from the OLD parameters, targets Y_n and the
//                         free-energy sums from the NEW ones (old gamma x new theta, GMM.py:303,312-314 / :473,485)
//
// When mu and w are frozen (two-set matching, GMM.py / api/ICP_two_set.py:179-187) or skip_M=True, ONE full row pass
// is the whole EM step.  Everything is evaluated in the log2 domain: with kappa = sqrt(log2(e)/2)/sigma and
// x' = kappa (x - origin), t2_nc = wl2_c - |x'_n - mu'_c|^2 is log2(pi_c N(x_n | mu_c, sigma)) and one MUFU.EX2 per
// pair gives the unnormalised responsibility.  Row / column maxima are tracked online; accumulators are only
// rescaled when the running maximum grows by more than 2^8 (exact in exact arithmetic, no overflow in fp32).
#pragma once
#include "pair_engine.cuh"

namespace dicp {

struct EmParams {
    const float* X;        // (N,D) data points
    const float* mu_old;   // (C,D) centroids that define the responsibilities
    const float* wl2;      // (C)   log2-domain scores (w_c - LSE(w) - D(ln sigma + ln(2 pi)/2)) * log2(e), old params
    const float* mu_new;   // (C,D) centroids used for the targets / free energy
    const float* lpi_new;  // (C)   natural-log mixture weights used for the free energy
    const float* T2;       // (N)   row LSE in log2 units (input of the column pass)
    const float* origin;   // D floats subtracted before scaling (= mu_old)
    float kappa;           // sqrt(log2(e)/2) / sigma_old
    float* o_T2;           // (N)   row LSE, log2 units
    float* o_Y;            // (N,D) quadratic targets
    float* o_rowP;         // (N)   sum_c gamma |mu'_c|^2 - |Y_n|^2            (optional, null = not written)
    float* o_rowQ;         // (N)   sum_c gamma (ln gamma - ln pi'_c)          (optional)
    float* o_sq;           // (N)   |x_n - Y_n|^2                              (optional)
    float* o_stats;        // (C, D+3) column statistics: m (log2), S0, B (D, unscaled), A (unscaled)
    const double* state;   // device-resident EM loop (em_col_small.cuh, EmState): kappa is read from it and the kernel does
                           // nothing once the loop's stop flag is set; null = kappa above, always run
};

// Device-resident state of an EM loop replayed as a CUDA graph (no host read between steps): doubles.
enum EmState { ES_SIGMA = 0, ES_KAPPA, ES_LGN, ES_DONE, ES_HAVE_LAST, ES_LAST_FE, ES_STEPS, ES_CFE, ES_FE, ES_N, ES_TOL,
               ES_SIGMA_NEW, ES_MAXIT, ES_COUNT = 16 };
// kernel prologue: returns false when the loop has stopped; otherwise takes kappa from the state
#define DICP_EM_STATE_PROLOGUE(P)                         \
    if ((P).state != nullptr) {                           \
        if ((P).state[ES_DONE] != 0.0) return;            \
        (P).kappa = (float)(P).state[ES_KAPPA];           \
    }

static constexpr float kRescaleSlack = 8.0f;
static constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------------------
// row pass: rows = points, cols = components
// ------------------------------------------------------------------------------------------------------------
template <int D, bool LITE>
struct EmRow {
    using Params = EmParams;
    // packed-fp32 form (pair_kernel_p): two components per call, one 64-bit register carries the same field of both; the
    // running maximum of a row is ONE scalar shared by the two halves (kept duplicated in the A_M register)
    static constexpr bool PACKED = true;
    static constexpr int THREADS = 128, MINB = 1, R = 2, TILE = 128;
    static constexpr int COLN = LITE ? (D + 1) : (2 * D + 3);
    static constexpr int NF = (COLN + 1) / 2 * 2;              // floats per column record (even)
    static constexpr int COLF4 = (NF + 3) / 4;
    static constexpr bool PAD_NULL = true;                     // a padded component contributes exactly 0
    // accumulators: m, S, then (full) Y(D), M2, L, T, DS
    static constexpr int A_M = 0, A_S = 1, A_Y = 2, A_M2 = 2 + D, A_L = 3 + D, A_T = 4 + D, A_DS = 5 + D;
    static constexpr int NACC = LITE ? 2 : (6 + D);
    static constexpr int NSCAL = LITE ? 0 : 4;   // P, Q, SQ, DS
    struct Row { float x[D]; float raw[D]; };   // raw = x - origin (unscaled)

    static DICP_HD void init_packed(F2* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = f2(0.f, 0.f);
        a[A_M] = f2(-3.0e38f, -3.0e38f);
    }
    static DICP_HD void unpack_acc(const F2* a, float* out) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) out[k] = f2_sum(a[k]);
        out[A_M] = vlane0(a[A_M]);
    }

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = (P.mu_old[(size_t)j * D + k] - P.origin[k]) * P.kappa;
            c[D] = P.wl2[j];
            if (!LITE) {
                float m2 = 0.f;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float v = P.mu_new[(size_t)j * D + k] - P.origin[k];   // centred: well-conditioned rowP
                    c[D + 1 + k] = v;
                    m2 = fmaf(v, v, m2);
                }
                c[2 * D + 1] = m2;
                c[2 * D + 2] = P.lpi_new[j];
            }
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
            c[D] = -1.0e30f;    // padded component: t2 = -inf-ish, never the maximum, e = 0
        }
    }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            r.raw[k] = P.X[(size_t)i * D + k] - P.origin[k];
            r.x[k] = r.raw[k] * P.kappa;
        }
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
        a[A_M] = -3.0e38f;
    }
    static DICP_HD void rescale(float* a, float m_new) {
        const float sc = ex2f_fast(a[A_M] - m_new);     // 0 when a[A_M] is the initial -3e38
#pragma unroll
        for (int k = 1; k < NACC; ++k) a[k] *= sc;
        a[A_M] = m_new;
    }
    static DICP_HD void combine(float* a, const float* b) {
        float bb[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) bb[k] = b[k];
        const float m = fmaxf(a[A_M], bb[A_M]);
        rescale(a, m);
        rescale(bb, m);
#pragma unroll
        for (int k = 1; k < NACC; ++k) a[k] += bb[k];
    }
    // V = float: one component; V = F2: two components (device kernel).  e = 2^(t2 - m) with m the row's running maximum
    // (raised, with a rescale of the sums, whenever a t2 exceeds it by more than kRescaleSlack).
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* a) {
        V r2;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const V z = vsub(vbc<V>(r.x[k]), c[k]);
            r2 = k == 0 ? vmul(z, z) : vfma(z, z, r2);
        }
        const V t2 = vsub(c[D], r2);
        float m = vlane0(a[A_M]);
        const float tmax = vhmax(t2);
        if (tmax > m + kRescaleSlack) {
            const V sc = vbc<V>(ex2f_fast(m - tmax));          // 0 when m is the initial -3e38
#pragma unroll
            for (int k = 1; k < NACC; ++k) a[k] = vmul(a[k], sc);
            a[A_M] = vbc<V>(tmax);
            m = tmax;
        }
        const V e = vex2n(vsub(vbc<V>(m), t2));
        a[A_S] = vadd(a[A_S], e);
        if (!LITE) {
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_Y + k] = vfma(e, c[D + 1 + k], a[A_Y + k]);
            a[A_M2] = vfma(e, c[2 * D + 1], a[A_M2]);
            a[A_L] = vfma(e, c[2 * D + 2], a[A_L]);
            a[A_T] = vfma(e, t2, a[A_T]);
            a[A_DS] = vfma(e, r2, a[A_DS]);
        }
    }
    // R rows against one (pair of) component(s): the same arithmetic as R calls of pair(), organised so that the R rows'
    // chains are independent up to ONE rarely taken branch (a per-row branch makes every row wait for its own compare:
    // measured 1.5-2x slower at 50 components).  Bit-identical to the per-row calls.
    static constexpr bool FUSED_ROWS = true;
    template <class V, int R>
    static DICP_HD void pair_rows(const Params& P, const Row (&r)[R], const V* c, V (&a)[R][NACC]) {
        V r2[R], t2[R];
        float tmax[R];
        bool need = false;
#pragma unroll
        for (int q = 0; q < R; ++q) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const V z = vsub(vbc<V>(r[q].x[k]), c[k]);
                r2[q] = k == 0 ? vmul(z, z) : vfma(z, z, r2[q]);
            }
            t2[q] = vsub(c[D], r2[q]);
            tmax[q] = vhmax(t2[q]);
            need = need || (tmax[q] > vlane0(a[q][A_M]) + kRescaleSlack);
        }
        if (need) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float m = vlane0(a[q][A_M]);
                if (tmax[q] > m + kRescaleSlack) {
                    const V sc = vbc<V>(ex2f_fast(m - tmax[q]));
#pragma unroll
                    for (int k = 1; k < NACC; ++k) a[q][k] = vmul(a[q][k], sc);
                    a[q][A_M] = vbc<V>(tmax[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const V e = vex2n(vsub(a[q][A_M], t2[q]));
            a[q][A_S] = vadd(a[q][A_S], e);
            if (!LITE) {
#pragma unroll
                for (int k = 0; k < D; ++k) a[q][A_Y + k] = vfma(e, c[D + 1 + k], a[q][A_Y + k]);
                a[q][A_M2] = vfma(e, c[2 * D + 1], a[q][A_M2]);
                a[q][A_L] = vfma(e, c[2 * D + 2], a[q][A_L]);
                a[q][A_T] = vfma(e, t2[q], a[q][A_T]);
                a[q][A_DS] = vfma(e, r2[q], a[q][A_DS]);
            }
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row& r, const float* a, float* scal) {
        const float S = a[A_S];
        const float T2 = a[A_M] + lg2f_fast(S);
        P.o_T2[i] = T2;
        if (!LITE) {
            const float inv = 1.0f / S;
            float y2 = 0.f, sq = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float y = a[A_Y + k] * inv;                    // target relative to the origin
                P.o_Y[(size_t)i * D + k] = y + P.origin[k];
                y2 = fmaf(y, y, y2);
                const float dxy = r.raw[k] - y;
                sq = fmaf(dxy, dxy, sq);
            }
            const float rowP = a[A_M2] * inv - y2;
            const float rowQ = (a[A_T] * inv - T2) * kLn2 - a[A_L] * inv;   // sum_c gamma (ln gamma_nc - ln pi'_c)
            const float ds = a[A_DS] * inv / (P.kappa * P.kappa);           // sum_c gamma |x_n - mu_c|^2 (old mu)
            if (P.o_rowP) P.o_rowP[i] = rowP;
            if (P.o_rowQ) P.o_rowQ[i] = rowQ;
            if (P.o_sq) P.o_sq[i] = sq;
            scal[0] = rowP; scal[1] = rowQ; scal[2] = sq; scal[3] = ds;
        }
    }
};

// ------------------------------------------------------------------------------------------------------------
// column pass: rows = components, cols = points (x'_n, T2_n)
// ------------------------------------------------------------------------------------------------------------
template <int D>
struct EmCol {
    using Params = EmParams;
    static constexpr bool PACKED = true;                       // two points per call (see EmRow)
    static constexpr int THREADS = 128, MINB = 1, R = 1, TILE = 128;
    static constexpr int NF = (D + 1 + 1) / 2 * 2;             // x' (D), T2, padded to an even count
    static constexpr int COLF4 = (NF + 3) / 4;
    static constexpr bool PAD_NULL = true;                     // a padded point contributes exactly 0
    static constexpr int A_M = 0, A_S = 1, A_B = 2, A_A = 2 + D;
    static constexpr int NACC = 3 + D, NSCAL = 0;
    struct Row { float mu[D]; float wl2; };

    static DICP_HD void init_packed(F2* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = f2(0.f, 0.f);
        a[A_M] = f2(-3.0e38f, -3.0e38f);
    }
    static DICP_HD void unpack_acc(const F2* a, float* out) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) out[k] = f2_sum(a[k]);
        out[A_M] = vlane0(a[A_M]);
    }

    static DICP_HD void pack_col(const Params& P, int j, int N, float* c) {
#pragma unroll
        for (int k = 0; k < COLF4 * 4; ++k) c[k] = 0.f;
        if (j < N) {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = (P.X[(size_t)j * D + k] - P.origin[k]) * P.kappa;
            c[D] = P.T2[j];
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = DICP_FAR;
            c[D] = 1.0e30f;      // padded point: log2 gamma = -huge, e = 0, never a new maximum
        }
    }
    static DICP_HD void load_row(const Params& P, int i, Row& r) {
#pragma unroll
        for (int k = 0; k < D; ++k) r.mu[k] = (P.mu_old[(size_t)i * D + k] - P.origin[k]) * P.kappa;
        r.wl2 = P.wl2[i];
    }
    static DICP_HD void init(float* a) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) a[k] = 0.f;
        a[A_M] = -3.0e38f;
    }
    static DICP_HD void rescale(float* a, float m_new) {
        const float sc = ex2f_fast(a[A_M] - m_new);
#pragma unroll
        for (int k = 1; k < NACC; ++k) a[k] *= sc;
        a[A_M] = m_new;
    }
    static DICP_HD void combine(float* a, const float* b) {
        float bb[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) bb[k] = b[k];
        const float m = fmaxf(a[A_M], bb[A_M]);
        rescale(a, m);
        rescale(bb, m);
#pragma unroll
        for (int k = 1; k < NACC; ++k) a[k] += bb[k];
    }
    template <class V>
    static DICP_HD void pair(const Params& P, const Row& r, const V* c, V* a) {
        V z[D];
        V r2;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            z[k] = vsub(c[k], vbc<V>(r.mu[k]));          // x'_n - mu'_c
            r2 = k == 0 ? vmul(z[k], z[k]) : vfma(z[k], z[k], r2);
        }
        const V l2 = vsub(vsub(vbc<V>(r.wl2), c[D]), r2);   // log2 gamma_nc  (<= 0 up to rounding)
        float m = vlane0(a[A_M]);
        const float lmax = vhmax(l2);
        if (lmax > m + kRescaleSlack) {
            const V sc = vbc<V>(ex2f_fast(m - lmax));
#pragma unroll
            for (int k = 1; k < NACC; ++k) a[k] = vmul(a[k], sc);
            a[A_M] = vbc<V>(lmax);
            m = lmax;
        }
        const V e = vex2n(vsub(vbc<V>(m), l2));
        a[A_S] = vadd(a[A_S], e);
#pragma unroll
        for (int k = 0; k < D; ++k) a[A_B + k] = vfma(e, z[k], a[A_B + k]);
        a[A_A] = vfma(e, r2, a[A_A]);
    }
    // U (pairs of) points against one component: as U calls of pair() when no running-maximum rescale is due inside the
    // block (one branch for the block, U independent chains); otherwise the calls one by one.  Bit-identical to them.
    template <class V, int U>
    static DICP_HD void pair_cols(const Params& P, const Row& r, const V (&c)[U][NF], V* a) {
        V z[U][D], r2[U], l2[U];
        const float m = vlane0(a[A_M]);
        bool need = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                z[u][k] = vsub(c[u][k], vbc<V>(r.mu[k]));
                r2[u] = k == 0 ? vmul(z[u][k], z[u][k]) : vfma(z[u][k], z[u][k], r2[u]);
            }
            l2[u] = vsub(vsub(vbc<V>(r.wl2), c[u][D]), r2[u]);
            need = need || (vhmax(l2[u]) > m + kRescaleSlack);
        }
        if (need) {
#pragma unroll
            for (int u = 0; u < U; ++u) pair<V>(P, r, c[u], a);
            return;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const V e = vex2n(vsub(a[A_M], l2[u]));
            a[A_S] = vadd(a[A_S], e);
#pragma unroll
            for (int k = 0; k < D; ++k) a[A_B + k] = vfma(e, z[u][k], a[A_B + k]);
            a[A_A] = vfma(e, r2[u], a[A_A]);
        }
    }
    static DICP_HD void finish(const Params& P, int i, const Row&, const float* a, float*) {
        float* o = P.o_stats + (size_t)i * (D + 3);
        const float ik = 1.0f / P.kappa;
        o[0] = a[A_M];
        o[1] = a[A_S];
#pragma unroll
        for (int k = 0; k < D; ++k) o[2 + k] = a[A_B + k] * ik;
        o[2 + D] = a[A_A] * ik * ik;
    }
};

}  // namespace dicp
