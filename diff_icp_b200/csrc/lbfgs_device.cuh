// Lock-step L-BFGS with the optimiser state ON THE DEVICE: one warp per frame.
//
// csrc/lbfgs_batch.cu holds one resumable L-BFGS state machine per frame on the HOST (the algorithm and settings of
// torch.optim.LBFGS(max_iter=20, max_eval=100, history_size=100, line_search_fn="strong_wolfe") as the reference uses it,
// /root/reference/diffICP/tools/optim.py:26,56): every lock-step round is a closure kernel, a device-to-host copy of losses
// and gradients, ~0.2 ms of host work and a host-to-device copy of the next trial points.  With the closure of all frames
// in ONE launch (cluster_closure.cuh) that host round trip is a third of the round.  Here the same state machine runs in a
// kernel right after the closure kernel -- one warp per frame, vectors in global memory with element i owned by lane i mod 32
// (so no cross-lane memory traffic: only the dot products meet, through fixed shuffle trees), scalars replicated in registers
// and evaluated uniformly by all lanes, fp64 like the host version -- and writes the next trial points and the active flags
// straight into the closure kernel's input buffers.  The pair (closure kernel, this kernel) is the body of a WHILE conditional
// node of a CUDA graph whose condition "some frame still waits for a closure value" is set by the last CTA of this kernel:
// one graph launch per optimizer.step() of ALL frames, no host round trip per round.
//
// The control flow mirrors lbfgs_batch.cu function by function (feed_one, iterate, bracket_result, zoom_*, line_search_done,
// after_iteration); read that file for the algorithm.  Dot products are summed lane-strided + shuffle tree here and
// sequentially there, so iterates agree to fp64 rounding, not bit for bit.
#pragma once
#include "small_step.cuh"     // last_cta
#include "../../include/dicp_b200.h"

namespace dicp {

static constexpr int kLdWarps = 4;             // frames per CTA
static constexpr int kLdNpl = 6;               // elements per lane at most: n <= 192 (64 support points in 3-D)

enum { LP_IDLE = 0, LP_WAIT_FIRST, LP_WAIT_BRACKET, LP_WAIT_ZOOM, LP_WAIT_PLAIN };
// per-frame slots (the layout the host side reads: include/dicp_b200.h documents the ones it uses)
enum { LI_N = 0, LI_LINESEARCH, LI_PHASE, LI_NITER, LI_CUREVALS, LI_OPTCOND, LI_FUNCEVALS, LI_NITERTOTAL, LI_LSITER, LI_LSEVALS,
       LI_LOW, LI_HIGH, LI_BRN, LI_DONE, LI_INSUF, LI_FIRSTBR, LI_HASBEST, LI_HCOUNT, LI_HHEAD, LI_COUNT = DICP_LBFGS_DEV_NI };
enum { LS_T = 0, LS_HDIAG, LS_LOSS, LS_PREVLOSS, LS_F0, LS_GTD0, LS_DNORM, LS_TPREV, LS_FPREV, LS_GTDPREV, LS_FNEW, LS_GTDNEW,
       LS_BRT0, LS_BRT1, LS_BRF0, LS_BRF1, LS_BRGTD0, LS_BRGTD1, LS_LASTEVAL, LS_BESTLOSS, LS_COUNT = DICP_LBFGS_DEV_ND };
enum { LV_X = 0, LV_D, LV_G, LV_PREVG, LV_XEVAL, LV_XINIT, LV_G0, LV_GPREV, LV_GNEW, LV_BRG0, LV_BRG1, LV_Q,
       LV_COUNT = DICP_LBFGS_DEV_NV };

struct LdWarp {
    const dicp_lbfgs_dev& B;
    int k, lane, n;
    int I[LI_COUNT];
    double S[LS_COUNT];

    DICP_D LdWarp(const dicp_lbfgs_dev& b, int kk, int ln) : B(b), k(kk), lane(ln) {
#pragma unroll
        for (int i = 0; i < LI_COUNT; ++i) I[i] = B.ints[(size_t)k * LI_COUNT + i];
#pragma unroll
        for (int i = 0; i < LS_COUNT; ++i) S[i] = B.dbl[(size_t)k * LS_COUNT + i];
        n = I[LI_N];
    }
    DICP_D void store() {
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < LI_COUNT; ++i) B.ints[(size_t)k * LI_COUNT + i] = I[i];
#pragma unroll
            for (int i = 0; i < LS_COUNT; ++i) B.dbl[(size_t)k * LS_COUNT + i] = S[i];
        }
    }
    DICP_D double* V(int v) const { return B.vec + ((size_t)k * LV_COUNT + v) * B.stride; }
    DICP_D double* dirs(int slot) const { return B.dirs + ((size_t)k * B.history + slot) * B.stride; }
    DICP_D double* stps(int slot) const { return B.stps + ((size_t)k * B.history + slot) * B.stride; }
    // ---- warp-cooperative vector helpers: element i belongs to lane i mod 32 in EVERY vector ------------------------------
    DICP_D double wsum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return __shfl_sync(0xffffffffu, v, 0);              // one value for the whole warp, whatever NaN payloads did
    }
    DICP_D double dot(const double* a, const double* b) const {
        double s = 0.0;
        for (int i = lane; i < n; i += 32) s += a[i] * b[i];
        return wsum(s);
    }
    DICP_D double amax(const double* a, double scale = 1.0) const {       // max |a_i * scale| (NaN of a lane's last element survives)
        double m = 0.0;
        for (int i = lane; i < n; i += 32) { const double w = fabs(a[i] * scale); if (!(w <= m)) m = w; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, m, o); if (!(w <= m)) m = w; }
        // the butterfly with a non-associative "max" may leave different lanes with different values when NaNs are around:
        return __shfl_sync(0xffffffffu, m, 0);
    }
    DICP_D void copy(double* dst, const double* src) const { for (int i = lane; i < n; i += 32) dst[i] = src[i]; }
    // br_* arrays indexed by a run-time 0 / 1
    DICP_D double brt(int i) const { return i ? S[LS_BRT1] : S[LS_BRT0]; }
    DICP_D double brf(int i) const { return i ? S[LS_BRF1] : S[LS_BRF0]; }
    DICP_D double brgtd(int i) const { return i ? S[LS_BRGTD1] : S[LS_BRGTD0]; }
    DICP_D void set_br(int i, double t, double f, double gtd, const double* g) {
        if (i) { S[LS_BRT1] = t; S[LS_BRF1] = f; S[LS_BRGTD1] = gtd; } else { S[LS_BRT0] = t; S[LS_BRF0] = f; S[LS_BRGTD0] = gtd; }
        copy(V(i ? LV_BRG1 : LV_BRG0), g);
    }
};

DICP_D double ld_cubic_min(double x1, double f1, double g1, double x2, double f2, double g2, double lo, double hi) {
    const double d1 = g1 + g2 - 3.0 * (f1 - f2) / (x1 - x2);
    const double sq = d1 * d1 - g1 * g2;
    if (sq >= 0.0) {
        const double d2 = sqrt(sq);
        const double pos = (x1 <= x2) ? x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2))
                                      : x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2));
        double r = pos;
        if (!(r >= lo)) r = lo;
        if (r > hi) r = hi;
        return r;
    }
    return 0.5 * (lo + hi);
}

// x_eval = fp32(base + t d); phase = ph
DICP_D void ld_request(LdWarp& F, const double* base, double t, int ph) {
    double* xe = F.V(LV_XEVAL);
    const double* d = F.V(LV_D);
    for (int i = F.lane; i < F.n; i += 32) xe[i] = (double)(float)(base[i] + t * d[i]);
    F.I[LI_PHASE] = ph;
}

DICP_D void ld_iterate(LdWarp& F);

DICP_D void ld_after_iteration(LdWarp& F, int ls_evals) {
    const dicp_lbfgs_dev& B = F.B;
    F.I[LI_CUREVALS] += ls_evals;
    F.I[LI_FUNCEVALS] += ls_evals;
    if (F.I[LI_NITER] == B.max_iter || F.I[LI_CUREVALS] >= B.max_eval || F.I[LI_OPTCOND]) { F.I[LI_PHASE] = LP_IDLE; return; }
    double m = 0.0;
    {   // max |d_i t| with the host's plain maximum (no NaN special case)
        const double* d = F.V(LV_D);
        for (int i = F.lane; i < F.n; i += 32) { const double w = fabs(d[i] * F.S[LS_T]); if (w > m) m = w; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, m, o); if (w > m) m = w; }
        m = __shfl_sync(0xffffffffu, m, 0);
    }
    if (m <= B.tol_change) { F.I[LI_PHASE] = LP_IDLE; return; }
    if (fabs(F.S[LS_LOSS] - F.S[LS_PREVLOSS]) < B.tol_change) { F.I[LI_PHASE] = LP_IDLE; return; }
    ld_iterate(F);
}

DICP_D void ld_line_search_done(LdWarp& F) {
    const int low = F.I[LI_LOW];
    const double t = F.brt(low);
    F.S[LS_LOSS] = F.brf(low);
    F.copy(F.V(LV_G), F.V(low ? LV_BRG1 : LV_BRG0));
    F.S[LS_T] = t;
    {
        double* x = F.V(LV_X);
        const double* xi = F.V(LV_XINIT);
        const double* d = F.V(LV_D);
        for (int i = F.lane; i < F.n; i += 32) x[i] = (double)(float)(xi[i] + t * d[i]);
    }
    F.I[LI_OPTCOND] = F.amax(F.V(LV_G)) <= F.B.tol_grad;
    ld_after_iteration(F, F.I[LI_LSEVALS]);
}

DICP_D void ld_zoom_continue(LdWarp& F) {
    const dicp_lbfgs_dev& B = F.B;
    if (F.I[LI_DONE] || F.I[LI_LSITER] >= B.max_ls) return ld_line_search_done(F);
    if (fabs(F.S[LS_BRT1] - F.S[LS_BRT0]) * F.S[LS_DNORM] < B.tol_change) return ld_line_search_done(F);
    const double bmax = F.S[LS_BRT0] > F.S[LS_BRT1] ? F.S[LS_BRT0] : F.S[LS_BRT1];
    const double bmin = F.S[LS_BRT0] < F.S[LS_BRT1] ? F.S[LS_BRT0] : F.S[LS_BRT1];
    double t = ld_cubic_min(F.S[LS_BRT0], F.S[LS_BRF0], F.S[LS_BRGTD0], F.S[LS_BRT1], F.S[LS_BRF1], F.S[LS_BRGTD1], bmin, bmax);
    const double eps = 0.1 * (bmax - bmin);
    const double gap = (bmax - t < t - bmin) ? bmax - t : t - bmin;
    if (gap < eps) {
        if (F.I[LI_INSUF] || t >= bmax || t <= bmin) {
            t = (fabs(t - bmax) < fabs(t - bmin)) ? bmax - eps : bmin + eps;
            F.I[LI_INSUF] = 0;
        } else {
            F.I[LI_INSUF] = 1;
        }
    } else {
        F.I[LI_INSUF] = 0;
    }
    F.S[LS_T] = t;
    ld_request(F, F.V(LV_XINIT), t, LP_WAIT_ZOOM);
}

DICP_D void ld_zoom_start(LdWarp& F) {
    F.I[LI_INSUF] = 0;
    const double flast = F.I[LI_BRN] == 2 ? F.S[LS_BRF1] : F.S[LS_BRF0];
    if (F.S[LS_BRF0] <= flast) { F.I[LI_LOW] = 0; F.I[LI_HIGH] = 1; } else { F.I[LI_LOW] = 1; F.I[LI_HIGH] = 0; }
    ld_zoom_continue(F);
}

DICP_D void ld_zoom_result(LdWarp& F) {
    const dicp_lbfgs_dev& B = F.B;
    F.I[LI_LSEVALS] += 1;
    F.I[LI_LSITER] += 1;
    const double t = F.S[LS_T];
    const int low = F.I[LI_LOW], high = F.I[LI_HIGH];
    const double* gnew = F.V(LV_GNEW);
    if (F.S[LS_FNEW] > (F.S[LS_F0] + B.c1 * t * F.S[LS_GTD0]) || F.S[LS_FNEW] >= F.brf(low)) {
        F.set_br(high, t, F.S[LS_FNEW], F.S[LS_GTDNEW], gnew);
        if (F.S[LS_BRF0] <= F.S[LS_BRF1]) { F.I[LI_LOW] = 0; F.I[LI_HIGH] = 1; } else { F.I[LI_LOW] = 1; F.I[LI_HIGH] = 0; }
    } else {
        if (fabs(F.S[LS_GTDNEW]) <= -B.c2 * F.S[LS_GTD0]) {
            F.I[LI_DONE] = 1;
        } else if (F.S[LS_GTDNEW] * (F.brt(high) - F.brt(low)) >= 0) {
            F.set_br(high, F.brt(low), F.brf(low), F.brgtd(low), F.V(low ? LV_BRG1 : LV_BRG0));
        }
        F.set_br(low, t, F.S[LS_FNEW], F.S[LS_GTDNEW], gnew);
    }
    ld_zoom_continue(F);
}

DICP_D void ld_set_bracket2(LdWarp& F, double t) {
    F.I[LI_BRN] = 2;
    F.set_br(0, F.S[LS_TPREV], F.S[LS_FPREV], F.S[LS_GTDPREV], F.V(LV_GPREV));
    F.set_br(1, t, F.S[LS_FNEW], F.S[LS_GTDNEW], F.V(LV_GNEW));
}

DICP_D void ld_bracket_result(LdWarp& F) {
    const dicp_lbfgs_dev& B = F.B;
    F.I[LI_LSEVALS] += 1;
    if (!F.I[LI_FIRSTBR]) {
        F.I[LI_LSITER] += 1;
        if (F.I[LI_LSITER] == B.max_ls) {                 // bracketing gave up: [0, t]
            F.I[LI_BRN] = 2;
            F.set_br(0, 0.0, F.S[LS_F0], F.S[LS_GTD0], F.V(LV_G0));
            F.set_br(1, F.S[LS_T], F.S[LS_FNEW], F.S[LS_GTDNEW], F.V(LV_GNEW));
            return ld_zoom_start(F);
        }
    }
    F.I[LI_FIRSTBR] = 0;
    const double t = F.S[LS_T];
    if (F.S[LS_FNEW] > (F.S[LS_F0] + B.c1 * t * F.S[LS_GTD0]) || (F.I[LI_LSITER] > 1 && F.S[LS_FNEW] >= F.S[LS_FPREV])) {
        ld_set_bracket2(F, t);
        return ld_zoom_start(F);
    }
    if (fabs(F.S[LS_GTDNEW]) <= -B.c2 * F.S[LS_GTD0]) {
        F.I[LI_BRN] = 1;
        F.set_br(0, t, F.S[LS_FNEW], F.S[LS_GTDNEW], F.V(LV_GNEW));
        F.I[LI_DONE] = 1;
        return ld_zoom_start(F);
    }
    if (F.S[LS_GTDNEW] >= 0) {
        ld_set_bracket2(F, t);
        return ld_zoom_start(F);
    }
    const double min_step = t + 0.01 * (t - F.S[LS_TPREV]), max_step = t * 10;
    const double tn = ld_cubic_min(F.S[LS_TPREV], F.S[LS_FPREV], F.S[LS_GTDPREV], t, F.S[LS_FNEW], F.S[LS_GTDNEW], min_step, max_step);
    F.S[LS_TPREV] = t; F.S[LS_FPREV] = F.S[LS_FNEW]; F.S[LS_GTDPREV] = F.S[LS_GTDNEW];
    F.copy(F.V(LV_GPREV), F.V(LV_GNEW));
    F.S[LS_T] = tn;
    ld_request(F, F.V(LV_XINIT), tn, LP_WAIT_BRACKET);
}

// top of one L-BFGS iteration: direction (two-loop recursion over the ring buffer of curvature pairs), initial step, then the
// line search's first evaluation
DICP_D void ld_iterate(LdWarp& F) {
    const dicp_lbfgs_dev& B = F.B;
    const int n = F.n, lane = F.lane;
    double* d = F.V(LV_D);
    double* g = F.V(LV_G);
    double* pg = F.V(LV_PREVG);
    F.I[LI_NITER] += 1;
    F.I[LI_NITERTOTAL] += 1;
    if (F.I[LI_NITERTOTAL] == 1) {
        for (int i = lane; i < n; i += 32) d[i] = -g[i];
        F.I[LI_HCOUNT] = 0;
        F.I[LI_HHEAD] = 0;
        F.S[LS_HDIAG] = 1;
    } else {
        // y = g - prev_g, s = d t go straight into the next ring slot; kept only when y.s > 1e-10
        const int m0 = F.I[LI_HCOUNT], head = F.I[LI_HHEAD];
        const int slot = (head + m0) % B.history;                     // m0 == history: this is the oldest pair's slot
        double* y = F.dirs(slot);
        double* s = F.stps(slot);
        double ys = 0.0, yy = 0.0;
        if (m0 < B.history) {
            for (int i = lane; i < n; i += 32) {
                const double yi = g[i] - pg[i], si = d[i] * F.S[LS_T];
                y[i] = yi; s[i] = si;
                ys += yi * si; yy += yi * yi;
            }
            ys = F.wsum(ys); yy = F.wsum(yy);
            if (ys > 1e-10) {
                F.I[LI_HCOUNT] = m0 + 1;
                F.S[LS_HDIAG] = ys / yy;
                if (lane == 0) B.ro[(size_t)F.k * B.history + slot] = 1.0 / ys;
            }
        } else {
            // full ring: the new pair may only overwrite the oldest one if it is accepted -- test first
            for (int i = lane; i < n; i += 32) {
                const double yi = g[i] - pg[i], si = d[i] * F.S[LS_T];
                ys += yi * si; yy += yi * yi;
            }
            ys = F.wsum(ys); yy = F.wsum(yy);
            if (ys > 1e-10) {
                for (int i = lane; i < n; i += 32) { y[i] = g[i] - pg[i]; s[i] = d[i] * F.S[LS_T]; }
                F.I[LI_HHEAD] = (head + 1) % B.history;
                F.S[LS_HDIAG] = ys / yy;
                if (lane == 0) B.ro[(size_t)F.k * B.history + slot] = 1.0 / ys;
            }
        }
        __syncwarp();                                                  // ro written by lane 0 is read by all lanes below
        // Two-loop recursion with q in REGISTERS (element e * 32 + lane, n <= 32 * kLdNpl) and the next pair's vectors
        // prefetched while the current dot product goes through its shuffle tree: the loop is a chain of dependent
        // reductions, and a global-memory round trip per link would dominate it.
        const int m = F.I[LI_HCOUNT], h0 = F.I[LI_HHEAD];
        const double* ro = B.ro + (size_t)F.k * B.history;
        double* al = B.al + (size_t)F.k * B.history;
        double q[kLdNpl], va[kLdNpl], vb[kLdNpl], na[kLdNpl], nb[kLdNpl];
#pragma unroll
        for (int e = 0; e < kLdNpl; ++e) { const int i = e * 32 + lane; q[e] = i < n ? -g[i] : 0.0; }
        auto load_pair = [&](int j, double (&sv)[kLdNpl], double (&yv)[kLdNpl], double& r) {
            const int sl = (h0 + j) % B.history;
            const double* sd = F.stps(sl);
            const double* yd = F.dirs(sl);
#pragma unroll
            for (int e = 0; e < kLdNpl; ++e) {
                const int i = e * 32 + lane;
                sv[e] = i < n ? sd[i] : 0.0;
                yv[e] = i < n ? yd[i] : 0.0;
            }
            r = ro[sl];
        };
        double rc = 0.0, rn = 0.0;
        if (m > 0) load_pair(m - 1, va, vb, rc);
        for (int j = m - 1; j >= 0; --j) {
            if (j > 0) load_pair(j - 1, na, nb, rn);
            double p = 0.0;
#pragma unroll
            for (int e = 0; e < kLdNpl; ++e) p += va[e] * q[e];
            const double a = F.wsum(p) * rc;
            if (lane == 0) al[(h0 + j) % B.history] = a;
#pragma unroll
            for (int e = 0; e < kLdNpl; ++e) { q[e] -= a * vb[e]; va[e] = na[e]; vb[e] = nb[e]; }
            rc = rn;
        }
#pragma unroll
        for (int e = 0; e < kLdNpl; ++e) q[e] *= F.S[LS_HDIAG];
        __syncwarp();                                                  // al written by lane 0 above
        double ac = 0.0, an = 0.0;
        if (m > 0) { load_pair(0, va, vb, rc); ac = al[h0 % B.history]; }
        for (int j = 0; j < m; ++j) {
            if (j + 1 < m) { load_pair(j + 1, na, nb, rn); an = al[(h0 + j + 1) % B.history]; }
            double p = 0.0;
#pragma unroll
            for (int e = 0; e < kLdNpl; ++e) p += vb[e] * q[e];
            const double be = F.wsum(p) * rc;
#pragma unroll
            for (int e = 0; e < kLdNpl; ++e) { q[e] += (ac - be) * va[e]; va[e] = na[e]; vb[e] = nb[e]; }
            rc = rn; ac = an;
        }
#pragma unroll
        for (int e = 0; e < kLdNpl; ++e) { const int i = e * 32 + lane; if (i < n) d[i] = q[e]; }
    }
    F.copy(pg, g);
    F.S[LS_PREVLOSS] = F.S[LS_LOSS];
    if (F.I[LI_NITERTOTAL] == 1) {
        double s1 = 0.0;
        for (int i = lane; i < n; i += 32) s1 += fabs(g[i]);
        s1 = F.wsum(s1);
        const double r = 1.0 / s1;
        F.S[LS_T] = (r < 1.0 ? r : 1.0) * B.lr;
    } else {
        F.S[LS_T] = B.lr;
    }
    const double gtd = F.dot(g, d);
    if (gtd > -B.tol_change) { F.I[LI_PHASE] = LP_IDLE; return; }
    if (F.I[LI_LINESEARCH]) {
        F.copy(F.V(LV_XINIT), F.V(LV_X));
        F.copy(F.V(LV_G0), g);
        F.S[LS_F0] = F.S[LS_LOSS]; F.S[LS_GTD0] = gtd;
        F.S[LS_DNORM] = F.amax(d);
        F.S[LS_TPREV] = 0; F.S[LS_FPREV] = F.S[LS_LOSS]; F.S[LS_GTDPREV] = gtd;
        F.copy(F.V(LV_GPREV), g);
        F.I[LI_DONE] = 0; F.I[LI_LSITER] = 0; F.I[LI_LSEVALS] = 0; F.I[LI_FIRSTBR] = 1; F.I[LI_BRN] = 0;
        ld_request(F, F.V(LV_XINIT), F.S[LS_T], LP_WAIT_BRACKET);
    } else {
        double* x = F.V(LV_X);
        for (int i = lane; i < n; i += 32) x[i] = (double)(float)(x[i] + F.S[LS_T] * d[i]);
        if (F.I[LI_NITER] != B.max_iter) {
            F.copy(F.V(LV_XEVAL), x);
            F.I[LI_PHASE] = LP_WAIT_PLAIN;
        } else {
            F.I[LI_PHASE] = LP_IDLE;        // after_iteration(0 evaluations): n_iter == max_iter ends the step
        }
    }
}

DICP_D void ld_feed_one(LdWarp& F, double loss, const float* grad) {
    const dicp_lbfgs_dev& B = F.B;
    const int n = F.n, lane = F.lane;
    F.S[LS_LASTEVAL] = loss;
    if (loss < F.S[LS_BESTLOSS]) {
        F.S[LS_BESTLOSS] = loss;
        const double* xe = F.V(LV_XEVAL);
        float* bx = B.best_x + (size_t)F.k * B.stride;
        for (int i = lane; i < n; i += 32) bx[i] = (float)xe[i];
        F.I[LI_HASBEST] = 1;
    }
    switch (F.I[LI_PHASE]) {
        case LP_WAIT_FIRST: {
            F.S[LS_LOSS] = loss;
            double* g = F.V(LV_G);
            for (int i = lane; i < n; i += 32) g[i] = (double)grad[i];
            F.I[LI_CUREVALS] = 1;
            F.I[LI_FUNCEVALS] += 1;
            F.I[LI_NITER] = 0;
            if (F.amax(g) <= B.tol_grad || B.max_iter < 1) { F.I[LI_PHASE] = LP_IDLE; return; }
            return ld_iterate(F);
        }
        case LP_WAIT_BRACKET:
        case LP_WAIT_ZOOM: {
            F.S[LS_FNEW] = loss;
            double* gn = F.V(LV_GNEW);
            for (int i = lane; i < n; i += 32) gn[i] = (double)grad[i];
            F.S[LS_GTDNEW] = F.dot(gn, F.V(LV_D));
            if (F.I[LI_PHASE] == LP_WAIT_BRACKET) return ld_bracket_result(F);
            return ld_zoom_result(F);
        }
        case LP_WAIT_PLAIN: {
            F.S[LS_LOSS] = loss;
            double* g = F.V(LV_G);
            for (int i = lane; i < n; i += 32) g[i] = (double)grad[i];
            F.I[LI_OPTCOND] = F.amax(g) <= B.tol_grad;
            return ld_after_iteration(F, 1);
        }
        default:
            return;
    }
}

// ---- kernels --------------------------------------------------------------------------------------------------------------------

// optimizer.step() begins for the frames with mask[k] != 0: trial point = current point, phase = WAIT_FIRST; the closure
// kernel's inputs (X row, active flag) are written for every frame.
__global__ void __launch_bounds__(kLdWarps * 32) lbfgs_dev_begin_kernel(dicp_lbfgs_dev B, const unsigned char* __restrict__ mask,
                                                                        float* __restrict__ X, long long xstride,
                                                                        int* __restrict__ active) {
    const int k = blockIdx.x * kLdWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= B.K) return;
    const bool on = mask == nullptr || mask[k] != 0;
    int* I = B.ints + (size_t)k * LI_COUNT;
    if (on) {
        const int n = I[LI_N];
        const double* x = B.vec + ((size_t)k * LV_COUNT + LV_X) * B.stride;
        double* xe = B.vec + ((size_t)k * LV_COUNT + LV_XEVAL) * B.stride;
        for (int i = lane; i < n; i += 32) {
            xe[i] = x[i];
            X[(size_t)k * xstride + i] = (float)x[i];
        }
        if (lane == 0) I[LI_PHASE] = LP_WAIT_FIRST;
    }
    if (lane == 0) active[k] = on ? 1 : 0;
    if (k == 0 && lane == 0) { B.counters[0] = 0u; B.counters[1] = 0u; B.counters[2] = 0u; }
}

// One lock-step round after the closure kernel: every waiting frame consumes its (loss, gradient) and either asks for the
// next closure value (X row, active = 1) or ends its step (active = 0).  out: the closure kernel's output rows
// [0, A, B, C, cost(1), data loss, ., . | gradient], loss = fp32(lam_reg (A/2 - eta B - eta^2 C / 2) + cost(1) + data loss) as
// shooting.BatchedClosurePlan computes it on the host.  The last CTA counts the waiting frames and, inside a WHILE node,
// sets the loop condition.
__global__ void __launch_bounds__(kLdWarps * 32) lbfgs_dev_feed_kernel(dicp_lbfgs_dev B, const float* __restrict__ out,
                                                                       long long ostride, int nscal, double lam_reg, double eta,
                                                                       float* __restrict__ X, long long xstride,
                                                                       int* __restrict__ active, int max_rounds, int use_cond,
                                                                       cudaGraphConditionalHandle cond) {
    const int k = blockIdx.x * kLdWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k < B.K && active[k] != 0) {
        LdWarp F(B, k, lane);
        if (F.I[LI_PHASE] != LP_IDLE) {
            const float* o = out + (size_t)k * ostride;
            const double H0 = 0.5 * (double)o[1] - eta * (double)o[2] - 0.5 * eta * eta * (double)o[3];
            const double loss = (double)(float)(lam_reg * H0 + (double)o[4] + (double)o[5]);
            ld_feed_one(F, loss, o + nscal);
            const bool wait = F.I[LI_PHASE] != LP_IDLE;
            if (wait) {
                const double* xe = F.V(LV_XEVAL);
                for (int i = lane; i < F.n; i += 32) X[(size_t)k * xstride + i] = (float)xe[i];
            }
            F.store();
            if (lane == 0) {
                active[k] = wait ? 1 : 0;
                if (wait) atomicAdd(&B.counters[1], 1u);
            }
        } else if (lane == 0) {
            active[k] = 0;
        }
    }
    if (last_cta(&B.counters[0], gridDim.x)) {
        if (threadIdx.x == 0) {
            const unsigned pending = *((volatile unsigned*)&B.counters[1]);
            const unsigned rounds = B.counters[2] + 1u;
            B.counters[2] = rounds;
            B.counters[3] = pending;
            B.counters[0] = 0u;
            B.counters[1] = 0u;
            if (use_cond) cudaGraphSetConditional(cond, (pending > 0u && rounds < (unsigned)max_rounds) ? 1u : 0u);
        }
    }
}

}  // namespace dicp
