// Lock-step L-BFGS for many independent small problems (HOST code; no device work in this file).
//
// DiffPSR.Reg_opt (/root/reference/diffICP/core/PSR.py:521-569) registers the K frames one after the other, each with its
// own torch.optim.LBFGS(max_iter=20, max_eval=100, history_size=100, line_search_fn="strong_wolfe") driven by
// LBFGS_optimization (/root/reference/diffICP/tools/optim.py:10-110).  With a small support set one closure evaluation
// of one frame is far too little work for a B200, so the frames are advanced TOGETHER: every round, each frame that is
// waiting for a closure value contributes its trial point, ONE batched device launch sequence evaluates all of them
// (shooting.BatchedClosurePlan), and every frame's optimiser consumes its own (loss, gradient) and moves on to its next
// trial point.  Frames stay completely independent: each has its own L-BFGS memory, its own strong-Wolfe line search
// (bracketing + zoom with cubic interpolation, the algorithm torch.optim.LBFGS documents and implements; c1 = 1e-4,
// c2 = 0.9, max_ls = 25) and its own stopping tests, so the iterates of a frame are those of the sequential algorithm up
// to floating-point rounding (vectors are kept in fp64 here, parameters are rounded to fp32 when they are evaluated).
//
// The optimiser of a frame is written as a resumable state machine: `advance` runs until the next closure evaluation is
// needed (phase WAIT_*) or the optimiser step is over (phase IDLE).
#include "../../include/dicp_b200.h"

#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <vector>

namespace {

typedef std::vector<double> Vec;

inline double dot(const Vec& a, const Vec& b) {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}
inline double amax(const Vec& a) {
    double m = 0.0;
    for (double v : a) { const double w = std::fabs(v); if (!(w <= m)) m = w; }       // NaN propagates
    return m;
}

// Minimiser of the cubic through (x1,f1,g1), (x2,f2,g2), clamped to [lo, hi].
double cubic_min(double x1, double f1, double g1, double x2, double f2, double g2, double lo, double hi) {
    const double d1 = g1 + g2 - 3.0 * (f1 - f2) / (x1 - x2);
    const double sq = d1 * d1 - g1 * g2;
    if (sq >= 0.0) {
        const double d2 = std::sqrt(sq);
        const double pos = (x1 <= x2) ? x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2.0 * d2))
                                      : x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2.0 * d2));
        double r = pos;
        if (!(r >= lo)) r = lo;                  // also catches NaN
        if (r > hi) r = hi;
        return r;
    }
    return 0.5 * (lo + hi);
}

enum Phase { IDLE = 0, WAIT_FIRST, WAIT_BRACKET, WAIT_ZOOM, WAIT_PLAIN };

struct Frame {
    int n = 0;
    bool line_search = true;
    Vec x;                                  // current parameters
    // optimiser memory (persists over steps, like torch.optim.LBFGS.state)
    Vec d, g, prev_g;
    double t = 0, H_diag = 1, loss = 0, prev_loss = 0;
    std::deque<Vec> old_dirs, old_stps;
    std::deque<double> ro;
    long func_evals = 0, n_iter_total = 0;
    bool have_prev = false;
    // one optimiser step
    int n_iter = 0, current_evals = 0;
    bool opt_cond = false;
    Phase phase = IDLE;
    Vec x_eval;                             // where the next closure value is wanted
    // line search
    Vec x_init, g0, g_prev, g_new, br_g[2];
    double f0 = 0, gtd0 = 0, d_norm = 0, t_prev = 0, f_prev = 0, gtd_prev = 0, f_new = 0, gtd_new = 0;
    double br_t[2] = {0, 0}, br_f[2] = {0, 0}, br_gtd[2] = {0, 0};
    int br_n = 0, ls_iter = 0, ls_evals = 0, low = 0, high = 1;
    bool done = false, insuf = false, first_bracket = true;
    // bookkeeping of the driver loop (tools/optim.py:32-50, 57): last / best closure values
    double last_eval = std::numeric_limits<double>::quiet_NaN();
    double best_loss = std::numeric_limits<double>::infinity();
    std::vector<float> best_x;
    bool has_best = false;
};

struct Batch {
    int K = 0;
    long stride = 0;
    int max_iter = 20, max_eval = 100, history = 100;
    double tol_grad = 1e-7, tol_change = 1e-9, lr = 1.0;
    double c1 = 1e-4, c2 = 0.9;
    int max_ls = 25;
    std::vector<Frame> f;
};

void request(Frame& F, const Vec& base, double t, Phase ph) {
    F.x_eval.resize(F.n);
    for (int i = 0; i < F.n; ++i) F.x_eval[i] = (double)(float)(base[i] + t * F.d[i]);
    F.phase = ph;
}

void zoom_continue(const Batch& B, Frame& F);
void iterate(const Batch& B, Frame& F);

void finish_step(Frame& F) { F.phase = IDLE; }

// break tests at the bottom of one L-BFGS iteration
void after_iteration(const Batch& B, Frame& F, int ls_evals) {
    F.current_evals += ls_evals;
    F.func_evals += ls_evals;
    if (F.n_iter == B.max_iter || F.current_evals >= B.max_eval || F.opt_cond) return finish_step(F);
    double m = 0.0;
    for (int i = 0; i < F.n; ++i) { const double w = std::fabs(F.d[i] * F.t); if (w > m) m = w; }
    if (m <= B.tol_change) return finish_step(F);
    if (std::fabs(F.loss - F.prev_loss) < B.tol_change) return finish_step(F);
    iterate(B, F);
}

void line_search_done(const Batch& B, Frame& F) {
    const double t = F.br_t[F.low];
    F.loss = F.br_f[F.low];
    F.g = F.br_g[F.low];
    F.t = t;
    for (int i = 0; i < F.n; ++i) F.x[i] = (double)(float)(F.x_init[i] + t * F.d[i]);
    F.opt_cond = amax(F.g) <= B.tol_grad;
    after_iteration(B, F, F.ls_evals);
}

void zoom_start(const Batch& B, Frame& F) {
    F.insuf = false;
    if (F.br_f[0] <= F.br_f[F.br_n - 1]) { F.low = 0; F.high = 1; } else { F.low = 1; F.high = 0; }
    zoom_continue(B, F);
}

void zoom_continue(const Batch& B, Frame& F) {
    if (F.done || F.ls_iter >= B.max_ls) return line_search_done(B, F);
    if (std::fabs(F.br_t[1] - F.br_t[0]) * F.d_norm < B.tol_change) return line_search_done(B, F);
    const double bmax = F.br_t[0] > F.br_t[1] ? F.br_t[0] : F.br_t[1];
    const double bmin = F.br_t[0] < F.br_t[1] ? F.br_t[0] : F.br_t[1];
    double t = cubic_min(F.br_t[0], F.br_f[0], F.br_gtd[0], F.br_t[1], F.br_f[1], F.br_gtd[1], bmin, bmax);
    const double eps = 0.1 * (bmax - bmin);
    const double gap = (bmax - t < t - bmin) ? bmax - t : t - bmin;
    if (gap < eps) {
        if (F.insuf || t >= bmax || t <= bmin) {
            t = (std::fabs(t - bmax) < std::fabs(t - bmin)) ? bmax - eps : bmin + eps;
            F.insuf = false;
        } else {
            F.insuf = true;
        }
    } else {
        F.insuf = false;
    }
    F.t = t;
    request(F, F.x_init, t, WAIT_ZOOM);
}

void zoom_result(const Batch& B, Frame& F) {
    F.ls_evals += 1;
    F.ls_iter += 1;
    const double t = F.t;
    if (F.f_new > (F.f0 + B.c1 * t * F.gtd0) || F.f_new >= F.br_f[F.low]) {
        F.br_t[F.high] = t; F.br_f[F.high] = F.f_new; F.br_g[F.high] = F.g_new; F.br_gtd[F.high] = F.gtd_new;
        if (F.br_f[0] <= F.br_f[1]) { F.low = 0; F.high = 1; } else { F.low = 1; F.high = 0; }
    } else {
        if (std::fabs(F.gtd_new) <= -B.c2 * F.gtd0) {
            F.done = true;
        } else if (F.gtd_new * (F.br_t[F.high] - F.br_t[F.low]) >= 0) {
            F.br_t[F.high] = F.br_t[F.low]; F.br_f[F.high] = F.br_f[F.low];
            F.br_g[F.high] = F.br_g[F.low]; F.br_gtd[F.high] = F.br_gtd[F.low];
        }
        F.br_t[F.low] = t; F.br_f[F.low] = F.f_new; F.br_g[F.low] = F.g_new; F.br_gtd[F.low] = F.gtd_new;
    }
    zoom_continue(B, F);
}

void set_bracket2(Frame& F, double t) {
    F.br_n = 2;
    F.br_t[0] = F.t_prev; F.br_t[1] = t;
    F.br_f[0] = F.f_prev; F.br_f[1] = F.f_new;
    F.br_g[0] = F.g_prev; F.br_g[1] = F.g_new;
    F.br_gtd[0] = F.gtd_prev; F.br_gtd[1] = F.gtd_new;
}

void bracket_result(const Batch& B, Frame& F) {
    F.ls_evals += 1;
    if (!F.first_bracket) {
        F.ls_iter += 1;
        if (F.ls_iter == B.max_ls) {                 // bracketing gave up: [0, t]
            F.br_n = 2;
            F.br_t[0] = 0; F.br_t[1] = F.t;
            F.br_f[0] = F.f0; F.br_f[1] = F.f_new;
            F.br_g[0] = F.g0; F.br_g[1] = F.g_new;
            F.br_gtd[0] = F.gtd0; F.br_gtd[1] = F.gtd_new;
            return zoom_start(B, F);
        }
    }
    F.first_bracket = false;
    const double t = F.t;
    if (F.f_new > (F.f0 + B.c1 * t * F.gtd0) || (F.ls_iter > 1 && F.f_new >= F.f_prev)) {
        set_bracket2(F, t);
        return zoom_start(B, F);
    }
    if (std::fabs(F.gtd_new) <= -B.c2 * F.gtd0) {
        F.br_n = 1;
        F.br_t[0] = t; F.br_f[0] = F.f_new; F.br_g[0] = F.g_new; F.br_gtd[0] = F.gtd_new;
        F.done = true;
        return zoom_start(B, F);
    }
    if (F.gtd_new >= 0) {
        set_bracket2(F, t);
        return zoom_start(B, F);
    }
    const double min_step = t + 0.01 * (t - F.t_prev), max_step = t * 10;
    const double tn = cubic_min(F.t_prev, F.f_prev, F.gtd_prev, t, F.f_new, F.gtd_new, min_step, max_step);
    F.t_prev = t; F.f_prev = F.f_new; F.g_prev = F.g_new; F.gtd_prev = F.gtd_new;
    F.t = tn;
    request(F, F.x_init, tn, WAIT_BRACKET);
}

// top of one L-BFGS iteration: direction, initial step, then the line search's first evaluation
void iterate(const Batch& B, Frame& F) {
    const int n = F.n;
    F.n_iter += 1;
    F.n_iter_total += 1;
    if (F.n_iter_total == 1) {
        F.d.resize(n);
        for (int i = 0; i < n; ++i) F.d[i] = -F.g[i];
        F.old_dirs.clear(); F.old_stps.clear(); F.ro.clear();
        F.H_diag = 1;
    } else {
        Vec y(n), s(n);
        for (int i = 0; i < n; ++i) { y[i] = F.g[i] - F.prev_g[i]; s[i] = F.d[i] * F.t; }
        const double ys = dot(y, s);
        if (ys > 1e-10) {
            if ((int)F.old_dirs.size() == B.history) { F.old_dirs.pop_front(); F.old_stps.pop_front(); F.ro.pop_front(); }
            F.H_diag = ys / dot(y, y);
            F.old_dirs.push_back(std::move(y)); F.old_stps.push_back(std::move(s)); F.ro.push_back(1.0 / ys);
        }
        const int m = (int)F.old_dirs.size();
        std::vector<double> al(m);
        Vec q(n);
        for (int i = 0; i < n; ++i) q[i] = -F.g[i];
        for (int i = m - 1; i >= 0; --i) {
            al[i] = dot(F.old_stps[i], q) * F.ro[i];
            for (int j = 0; j < n; ++j) q[j] -= al[i] * F.old_dirs[i][j];
        }
        for (int j = 0; j < n; ++j) q[j] *= F.H_diag;
        for (int i = 0; i < m; ++i) {
            const double be = dot(F.old_dirs[i], q) * F.ro[i];
            for (int j = 0; j < n; ++j) q[j] += (al[i] - be) * F.old_stps[i][j];
        }
        F.d = std::move(q);
    }
    F.prev_g = F.g;
    F.have_prev = true;
    F.prev_loss = F.loss;
    if (F.n_iter_total == 1) {
        double s1 = 0.0;
        for (double v : F.g) s1 += std::fabs(v);
        const double r = 1.0 / s1;
        F.t = (r < 1.0 ? r : 1.0) * B.lr;
    } else {
        F.t = B.lr;
    }
    const double gtd = dot(F.g, F.d);
    if (gtd > -B.tol_change) return finish_step(F);
    if (F.line_search) {
        F.x_init = F.x;
        F.g0 = F.g; F.f0 = F.loss; F.gtd0 = gtd;
        F.d_norm = amax(F.d);
        F.t_prev = 0; F.f_prev = F.loss; F.g_prev = F.g; F.gtd_prev = gtd;
        F.done = false; F.ls_iter = 0; F.ls_evals = 0; F.first_bracket = true; F.br_n = 0;
        request(F, F.x_init, F.t, WAIT_BRACKET);
    } else {
        for (int i = 0; i < n; ++i) F.x[i] = (double)(float)(F.x[i] + F.t * F.d[i]);
        if (F.n_iter != B.max_iter) {
            F.x_eval = F.x;
            F.phase = WAIT_PLAIN;
        } else {
            after_iteration(B, F, 0);
        }
    }
}

void feed_one(const Batch& B, Frame& F, double loss, const float* grad) {
    const int n = F.n;
    F.last_eval = loss;
    if (loss < F.best_loss) {
        F.best_loss = loss;
        F.best_x.resize(n);
        for (int i = 0; i < n; ++i) F.best_x[i] = (float)F.x_eval[i];
        F.has_best = true;
    }
    switch (F.phase) {
        case WAIT_FIRST: {
            F.loss = loss;
            F.g.assign(grad, grad + n);
            F.current_evals = 1;
            F.func_evals += 1;
            F.n_iter = 0;
            if (amax(F.g) <= B.tol_grad) return finish_step(F);
            if (B.max_iter < 1) return finish_step(F);
            return iterate(B, F);
        }
        case WAIT_BRACKET:
        case WAIT_ZOOM: {
            F.f_new = loss;
            F.g_new.assign(grad, grad + n);
            F.gtd_new = dot(F.g_new, F.d);
            if (F.phase == WAIT_BRACKET) return bracket_result(B, F);
            return zoom_result(B, F);
        }
        case WAIT_PLAIN: {
            F.loss = loss;
            F.g.assign(grad, grad + n);
            F.opt_cond = amax(F.g) <= B.tol_grad;
            return after_iteration(B, F, 1);
        }
        default:
            return;
    }
}

inline Batch* as_batch(void* h) { return reinterpret_cast<Batch*>(h); }

}  // namespace

extern "C" {

void* dicp_lbfgs_create(int K, const int64_t* n, int64_t stride, int max_iter, int max_eval, int history,
                        double tolerance_grad, double tolerance_change) {
    if (K < 1 || !n || max_iter < 0 || max_eval < 1 || history < 1) return nullptr;
    for (int k = 0; k < K; ++k)
        if (n[k] < 1 || n[k] > stride) return nullptr;
    Batch* B = new Batch();
    B->K = K; B->stride = (long)stride;
    B->max_iter = max_iter; B->max_eval = max_eval; B->history = history;
    B->tol_grad = tolerance_grad; B->tol_change = tolerance_change;
    B->f.resize(K);
    for (int k = 0; k < K; ++k) {
        B->f[k].n = (int)n[k];
        B->f[k].x.assign((size_t)n[k], 0.0);
    }
    return B;
}

void dicp_lbfgs_destroy(void* h) { delete as_batch(h); }

int dicp_lbfgs_set_x(void* h, int k, const float* x) {
    Batch* B = as_batch(h);
    if (!B || k < 0 || k >= B->K || !x) return DICP_EBADARG;
    Frame& F = B->f[k];
    for (int i = 0; i < F.n; ++i) F.x[i] = (double)x[i];
    return DICP_OK;
}

int dicp_lbfgs_get_x(void* h, int k, float* x, int best) {
    Batch* B = as_batch(h);
    if (!B || k < 0 || k >= B->K || !x) return DICP_EBADARG;
    Frame& F = B->f[k];
    if (best) {
        if (!F.has_best) return DICP_EBADARG;
        std::memcpy(x, F.best_x.data(), sizeof(float) * (size_t)F.n);
    } else {
        for (int i = 0; i < F.n; ++i) x[i] = (float)F.x[i];
    }
    return DICP_OK;
}

int dicp_lbfgs_reset(void* h, int k, int line_search) {
    Batch* B = as_batch(h);
    if (!B || k < 0 || k >= B->K) return DICP_EBADARG;
    Frame& F = B->f[k];
    F.line_search = line_search != 0;
    F.old_dirs.clear(); F.old_stps.clear(); F.ro.clear();
    F.func_evals = 0; F.n_iter_total = 0; F.have_prev = false; F.H_diag = 1; F.t = 0;
    F.phase = IDLE;
    return DICP_OK;
}

int dicp_lbfgs_begin_step(void* h, const uint8_t* mask) {
    Batch* B = as_batch(h);
    if (!B) return DICP_EBADARG;
    int cnt = 0;
    for (int k = 0; k < B->K; ++k) {
        if (mask && !mask[k]) continue;
        Frame& F = B->f[k];
        if (F.phase != IDLE) return DICP_EBADARG;
        F.x_eval = F.x;
        F.phase = WAIT_FIRST;
        ++cnt;
    }
    return cnt;
}

int dicp_lbfgs_pending(void* h, float* X, uint8_t* active) {
    Batch* B = as_batch(h);
    if (!B || !active) return DICP_EBADARG;
    int cnt = 0;
    for (int k = 0; k < B->K; ++k) {
        Frame& F = B->f[k];
        active[k] = F.phase != IDLE;
        if (F.phase == IDLE) continue;
        ++cnt;
        if (X) {
            float* dst = X + (size_t)k * (size_t)B->stride;
            for (int i = 0; i < F.n; ++i) dst[i] = (float)F.x_eval[i];
        }
    }
    return cnt;
}

int dicp_lbfgs_feed(void* h, const float* losses, const float* grads) {
    Batch* B = as_batch(h);
    if (!B || !losses || !grads) return DICP_EBADARG;
    int cnt = 0;
    for (int k = 0; k < B->K; ++k) {
        Frame& F = B->f[k];
        if (F.phase == IDLE) continue;
        feed_one(*B, F, (double)losses[k], grads + (size_t)k * (size_t)B->stride);
        if (F.phase != IDLE) ++cnt;
    }
    return cnt;
}

int dicp_lbfgs_get_all(void* h, float* X, int best) {
    Batch* B = as_batch(h);
    if (!B || !X) return DICP_EBADARG;
    for (int k = 0; k < B->K; ++k) {
        const Frame& F = B->f[k];
        float* dst = X + (size_t)k * (size_t)B->stride;
        if (best) {
            if (!F.has_best) return DICP_EBADARG;
            std::memcpy(dst, F.best_x.data(), sizeof(float) * (size_t)F.n);
        } else {
            for (int i = 0; i < F.n; ++i) dst[i] = (float)F.x[i];
        }
    }
    return DICP_OK;
}

int dicp_lbfgs_stats_all(void* h, double* out) {
    Batch* B = as_batch(h);
    if (!B || !out) return DICP_EBADARG;
    for (int k = 0; k < B->K; ++k) {
        const Frame& F = B->f[k];
        out[4 * k] = F.last_eval;
        out[4 * k + 1] = F.best_loss;
        out[4 * k + 2] = (double)F.func_evals;
        out[4 * k + 3] = (double)F.n_iter_total;
    }
    return DICP_OK;
}

int dicp_lbfgs_stats(void* h, int k, double* out4) {
    Batch* B = as_batch(h);
    if (!B || k < 0 || k >= B->K || !out4) return DICP_EBADARG;
    const Frame& F = B->f[k];
    out4[0] = F.last_eval;
    out4[1] = F.best_loss;
    out4[2] = (double)F.func_evals;
    out4[3] = (double)F.n_iter_total;
    return DICP_OK;
}

}  // extern "C"
