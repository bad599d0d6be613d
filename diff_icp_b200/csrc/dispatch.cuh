// Op dispatch shared by the device library (capi.cu) and the CPU emulation used by the tests
// (tests/hostemu/hostemu.cpp).  `Exec` decides where an Op runs:
//   ex.run<Op>(prm, rows, cols, scal_out, accumulate)   pack + pair (+ finish, scalar reduction)
//   ex.scal_fix(scal, eta, withdiv)                     scal[0] = withdiv ? scal[2] + eta*scal[3] : 0
#pragma once
#include <math.h>
#include "ops_ksum.cuh"
#include "ops_rhs.cuh"
#include "ops_em.cuh"

namespace dicp {

struct GaussConst {
    float kappa, s, alpha, beta;
};
inline GaussConst gauss_const(float sigma) {
    GaussConst g;
    const double s = 1.0 / ((double)sigma * (double)sigma);
    const double kappa = sqrt(0.5 * 1.4426950408889634) / (double)sigma;   // sqrt(log2(e)/2)/sigma
    g.kappa = (float)kappa;
    g.s = (float)s;
    g.alpha = (float)(s / kappa);
    g.beta = (float)(s / (kappa * kappa));                                  // = 2 ln 2
    return g;
}

template <int D, class Exec>
int ksum_dispatch(Exec& ex, unsigned mask, const KsumParams& prm, int M, int N) {
    switch (mask) {
#define DICP_CASE(m) case (m): return ex.template run<KsumOp<D, (m)>>(prm, M, N, nullptr, 0);
        DICP_CASE(K_BASE)
        DICP_CASE(K_REDSCAL)
        DICP_CASE(K_RED)
        DICP_CASE(K_GRAD)
        DICP_CASE(K_DD)
        DICP_CASE(K_GEND)
        DICP_CASE(K_HESS)
        DICP_CASE(K_LAP)
        DICP_CASE(K_GRADLAP)
        DICP_CASE(K_MINSQ)
        DICP_CASE(K_DOT)
        DICP_CASE(K_RED | K_GRAD)               // v(x) with the gradient component (LDDMM.py:114)
        DICP_CASE(K_GRAD | K_LAP)               // mdivsum with the gradient component (LDDMM.py:135)
        DICP_CASE(K_RED | K_GRAD | K_LAP)       // Hamiltonian pieces (LDDMM.py:151-153)
        DICP_CASE(K_BASE | K_REDSCAL)
#undef DICP_CASE
        default: return DICP_EUNSUPPORTED;
    }
}

template <class Exec>
int ksum_entry(Exec& ex, int D, unsigned mask, float sigma, const float* x, int64_t M, const float* y, int64_t N,
               const float* b, const float* c, const float* d, float* const* outs) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || M < 0 || N < 1 || M > INT32_MAX || N > INT32_MAX) return DICP_EBADARG;
    if (M == 0) return DICP_OK;
    if (x == nullptr || y == nullptr) return DICP_EBADARG;
    if ((mask & (K_RED | K_DD | K_GEND | K_HESS | K_DOT)) && b == nullptr) return DICP_EBADARG;
    if ((mask & (K_GEND | K_HESS)) && c == nullptr) return DICP_EBADARG;
    if ((mask & K_REDSCAL) && d == nullptr) return DICP_EBADARG;
    for (int k = 0; k < 11; ++k)
        if ((mask >> k & 1u) && outs[k] == nullptr) return DICP_EBADARG;
    GaussConst g = gauss_const(sigma);
    KsumParams prm{x, y, b, d, c, y, g.kappa, g.s, g.alpha, g.beta,
                   outs[0], outs[1], outs[2], outs[3], outs[4], outs[5], outs[6], outs[7], outs[8], outs[9], outs[10]};
    return D == 2 ? ksum_dispatch<2>(ex, mask, prm, (int)M, (int)N) : ksum_dispatch<3>(ex, mask, prm, (int)M, (int)N);
}

template <int D, class Exec>
int rhs_forward_d(Exec& ex, int withlogdet, const RhsParams& prm, int M, int Nx, float* scal) {
    const bool hasx = prm.x != nullptr && Nx > 0;
    const bool eta = prm.eta != 0.f;
    const bool div_qq = withlogdet && !hasx;
    int rc;
    // (q,q) pass: scal[1..3] = A, B, C
    if (ex.use_sym_forward(M)) {                    // every unordered pair once (symmetric engine; measured SLOWER than the
                                                    // general engine for the forward pass on B200: off unless mode 2)
        if (eta) rc = ex.template run_sym<RhsQQ<D, true, true>>(prm, M, scal + 1);
        else if (div_qq) rc = ex.template run_sym<RhsQQ<D, true, false>>(prm, M, scal + 1);
        else rc = ex.template run_sym<RhsQQ<D, false, false>>(prm, M, scal + 1);
    } else if (eta) rc = ex.template run<RhsQQ<D, true, true>>(prm, M, M, scal + 1, 0);
    else if (div_qq) rc = ex.template run<RhsQQ<D, true, false>>(prm, M, M, scal + 1, 0);
    else rc = ex.template run<RhsQQ<D, false, false>>(prm, M, M, scal + 1, 0);
    if (rc != DICP_OK) return rc;
    ex.scal_fix(scal, prm.eta, div_qq ? 1 : 0);
    if (hasx) {
        // (x,q) pass: scal[0] += dcost contribution
        if (eta) rc = ex.template run<RhsXQ<D, true, true>>(prm, Nx, M, scal, 1);
        else if (withlogdet) rc = ex.template run<RhsXQ<D, true, false>>(prm, Nx, M, scal, 1);
        else rc = ex.template run<RhsXQ<D, false, false>>(prm, Nx, M, scal, 1);
    }
    return rc;
}

template <class Exec>
int rhs_forward_entry(Exec& ex, int D, int withlogdet, float sigma, float eta, const float* q, const float* p,
                      int64_t M, const float* x, int64_t Nx, float* vq, float* dp, float* vx, float* scal) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || M < 1 || Nx < 0 || M > INT32_MAX || Nx > INT32_MAX) return DICP_EBADARG;
    if (!q || !p || !vq || !dp || !scal) return DICP_EBADARG;
    if (Nx > 0 && x != nullptr && vx == nullptr) return DICP_EBADARG;
    if (eta != 0.f && !withlogdet) return DICP_EUNSUPPORTED;   // gradcomponent without logdet is not a reference model
    GaussConst g = gauss_const(sigma);
    RhsParams prm{};
    prm.q = q; prm.p = p; prm.x = (Nx > 0 ? x : nullptr); prm.origin = q;
    prm.kappa = g.kappa; prm.s = g.s; prm.alpha = g.alpha; prm.beta = g.beta; prm.eta = eta;
    prm.vq = vq; prm.dp = dp; prm.vx = vx;
    return D == 2 ? rhs_forward_d<2>(ex, withlogdet, prm, (int)M, (int)Nx, scal)
                  : rhs_forward_d<3>(ex, withlogdet, prm, (int)M, (int)Nx, scal);
}

template <int D, class Exec>
int rhs_adjoint_d(Exec& ex, int withlogdet, RhsParams prm, int M, int Nx) {
    const bool hasx = prm.x != nullptr && Nx > 0;
    int rc;
    prm.accumulate = 0;
    if (prm.eta != 0.f) {                           // logdet model (withlogdet is implied)
        rc = ex.use_sym(M) ? ex.template run_sym<AdjQQEta<D>>(prm, M, nullptr) : ex.template run<AdjQQEta<D>>(prm, M, M, nullptr, 0);
        if (rc != DICP_OK || !hasx) return rc;
        if (ex.use_rect(Nx, M)) return ex.template run_rect<AdjXQE<D>>(prm, Nx, M);      // both sides in one ring pass
        rc = ex.template run<AdjXQxEta<D>>(prm, Nx, M, nullptr, 0);
        if (rc != DICP_OK) return rc;
        prm.accumulate = 1;
        return ex.template run<AdjXQqEta<D>>(prm, M, Nx, nullptr, 0);
    }
    const bool div_qq = withlogdet && !hasx;
    if (ex.use_sym(M)) {
        if (div_qq) rc = ex.template run_sym<AdjQQ<D, true>>(prm, M, nullptr);
        else rc = ex.template run_sym<AdjQQ<D, false>>(prm, M, nullptr);
    } else {
        if (div_qq) rc = ex.template run<AdjQQ<D, true>>(prm, M, M, nullptr, 0);
        else rc = ex.template run<AdjQQ<D, false>>(prm, M, M, nullptr, 0);
    }
    if (rc != DICP_OK) return rc;
    if (hasx && ex.use_rect(Nx, M)) {               // both sides of the (x,q) pass from one evaluation of every pair
        if (withlogdet) return ex.template run_rect<AdjXQ<D, true>>(prm, Nx, M);
        return ex.template run_rect<AdjXQ<D, false>>(prm, Nx, M);
    }
    if (hasx) {
        if (withlogdet) rc = ex.template run<AdjXQx<D, true>>(prm, Nx, M, nullptr, 0);
        else rc = ex.template run<AdjXQx<D, false>>(prm, Nx, M, nullptr, 0);
        if (rc != DICP_OK) return rc;
        prm.accumulate = 1;
        if (withlogdet) rc = ex.template run<AdjXQq<D, true>>(prm, M, Nx, nullptr, 0);
        else rc = ex.template run<AdjXQq<D, false>>(prm, M, Nx, nullptr, 0);
    }
    return rc;
}

template <class Exec>
int rhs_adjoint_entry(Exec& ex, int D, int withlogdet, float sigma, float eta, const float* q, const float* p,
                      int64_t M, const float* x, int64_t Nx, const float* a, const float* u, const float* wx,
                      const float* gc, float* gq, float* gp, float* gx) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || M < 1 || Nx < 0 || M > INT32_MAX || Nx > INT32_MAX) return DICP_EBADARG;
    if (!q || !p || !a || !u || !gq || !gp) return DICP_EBADARG;
    if (Nx > 0 && x != nullptr && (wx == nullptr || gx == nullptr)) return DICP_EBADARG;
    GaussConst g = gauss_const(sigma);
    RhsParams prm{};
    prm.q = q; prm.p = p; prm.x = (Nx > 0 ? x : nullptr); prm.origin = q;
    prm.a = a; prm.u = u; prm.wx = wx; prm.gc = gc;
    prm.kappa = g.kappa; prm.s = g.s; prm.alpha = g.alpha; prm.beta = g.beta; prm.eta = eta;
    prm.gq = gq; prm.gp = gp; prm.gx = gx;
    return D == 2 ? rhs_adjoint_d<2>(ex, withlogdet, prm, (int)M, (int)Nx)
                  : rhs_adjoint_d<3>(ex, withlogdet, prm, (int)M, (int)Nx);
}

// ---- GMM EM ---------------------------------------------------------------------------------------------
template <class Exec>
int em_rowpass_entry(Exec& ex, int D, int lite, float sigma_old, const float* X, int64_t N, const float* mu_old,
                     const float* wl2, int64_t C, const float* mu_new, const float* lpi_new, float* T2, float* Y,
                     float* rowP, float* rowQ, float* sq, float* scal4) {
    if ((D != 2 && D != 3) || !(sigma_old > 0.f) || N < 0 || C < 1 || N > INT32_MAX || C > INT32_MAX) return DICP_EBADARG;
    if (N == 0) return DICP_OK;
    if (!X || !mu_old || !wl2 || !T2) return DICP_EBADARG;
    if (!lite && (!mu_new || !lpi_new || !Y || !scal4)) return DICP_EBADARG;
    EmParams prm{};
    prm.X = X; prm.mu_old = mu_old; prm.wl2 = wl2; prm.mu_new = mu_new; prm.lpi_new = lpi_new; prm.origin = mu_old;
    prm.kappa = gauss_const(sigma_old).kappa;
    prm.o_T2 = T2; prm.o_Y = Y; prm.o_rowP = rowP; prm.o_rowQ = rowQ; prm.o_sq = sq;
    if (lite) return D == 2 ? ex.template run<EmRow<2, true>>(prm, (int)N, (int)C, nullptr, 0)
                            : ex.template run<EmRow<3, true>>(prm, (int)N, (int)C, nullptr, 0);
    return D == 2 ? ex.template run<EmRow<2, false>>(prm, (int)N, (int)C, scal4, 0)
                  : ex.template run<EmRow<3, false>>(prm, (int)N, (int)C, scal4, 0);
}

template <class Exec>
int em_colstats_entry(Exec& ex, int D, float sigma_old, const float* X, int64_t N, const float* T2,
                      const float* mu_old, const float* wl2, int64_t C, float* stats) {
    if ((D != 2 && D != 3) || !(sigma_old > 0.f) || N < 1 || C < 1 || N > INT32_MAX || C > INT32_MAX) return DICP_EBADARG;
    if (!X || !T2 || !mu_old || !wl2 || !stats) return DICP_EBADARG;
    EmParams prm{};
    prm.X = X; prm.T2 = T2; prm.mu_old = mu_old; prm.wl2 = wl2; prm.origin = mu_old;
    prm.kappa = gauss_const(sigma_old).kappa;
    prm.o_stats = stats;
    return D == 2 ? ex.template run<EmCol<2>>(prm, (int)C, (int)N, nullptr, 0)
                  : ex.template run<EmCol<3>>(prm, (int)C, (int)N, nullptr, 0);
}

// CPU executor: tests only (tests/hostemu).  Never part of libdicp_b200.so.
struct HostExec {
    bool sym = false;                       // evaluate the (q,q) adjoint pass through Op::pair_sym (tests of the formulas)
    bool use_sym(int) const { return sym; }
    bool use_sym_forward(int) const { return sym; }
    bool use_rect(int, int) const { return sym; }
    template <class Op>
    int run_rect(const typename Op::Params& prm, int Mrows, int Ncols) {
        run_rect_host<Op>(prm, Mrows, Ncols);
        return DICP_OK;
    }
    template <class Op>
    int run_sym(const typename Op::Params& prm, int M, float* scal_out) {
        run_pair_host_sym<Op>(prm, M, scal_out);
        return DICP_OK;
    }
    template <class Op>
    int run(const typename Op::Params& prm, int M, int N, float* scal_out, int accumulate) {
        float tmp[8] = {0};
        run_pair_host<Op>(prm, M, N, tmp);
        if (scal_out)
            for (int k = 0; k < Op::NSCAL; ++k) scal_out[k] = accumulate ? scal_out[k] + tmp[k] : tmp[k];
        return DICP_OK;
    }
    void scal_fix(float* scal, float eta, int withdiv) { scal[0] = withdiv ? scal[2] + eta * scal[3] : 0.f; }
};

}  // namespace dicp
