// extern "C" entry points of libdicp_b200.so (see include/dicp_b200.h).
#include <atomic>
#include "../../include/dicp_b200.h"
#include "dispatch.cuh"
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "small_step.cuh"
#include "sym_engine.cuh"
#include "batch_closure.cuh"
#include "cluster_closure.cuh"
#include "lbfgs_device.cuh"
#include "pointset.cuh"
#include "em_col_small.cuh"

using namespace dicp;

namespace {

__global__ void axpy_kernel(long long n, float* __restrict__ out, const float* __restrict__ a, float alpha,
                            const float* __restrict__ f1, float beta, const float* __restrict__ f2) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float v = fmaf(alpha, f1[i], a[i]);
        if (f2 != nullptr) v = fmaf(beta, f2[i], v);
        out[i] = v;
    }
}

// dcost contribution of the (q,q) pass:  scal[0] = withdiv ? B + eta*C : 0   (scal = {_, A, B, C})
__global__ void rhs_scal_fix_kernel(float* scal, float eta, int withdiv) {
    if (threadIdx.x == 0) scal[0] = withdiv ? fmaf(eta, scal[3], scal[2]) : 0.f;
}

// Quadratic data loss of DiffPSR.QuadLossFunctor (core/PSR.py:498-516): loss = sum_n inv_n |x_n - y_n|^2 and its gradient
// g_n = 2 inv_n (x_n - y_n); block partial sums in a fixed order (deterministic), summed by scalar_reduce_kernel.
__global__ void quad_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ inv,
                                 long long n, int D, float* __restrict__ g, float* __restrict__ blocksum) {
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float w = inv[i];
        for (int k = 0; k < D; ++k) {
            const float r = x[i * D + k] - y[i * D + k];
            g[i * D + k] = 2.f * w * r;
            acc = fmaf(w * r, r, acc);
        }
    }
    const float v = block_sum(acc, red);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = v;
}

// symmetric engine switch: default on; DICP_SYM=0 in the environment or dicp_sym_mode(0) disable it (A/B measurements, tests)
// (process-wide DEFAULT only: dicp_rhs_forward / dicp_rhs_adjoint take the engine as a per-call argument)
inline std::atomic<int>& sym_mode_ref() {
    static std::atomic<int> mode([] {
        const char* e = getenv("DICP_SYM");
        return e ? atoi(e) : 1;
    }());
    return mode;
}
inline int sym_mode() { return sym_mode_ref().load(std::memory_order_relaxed); }

struct DeviceExec {
    void* ws;
    size_t wsb;
    cudaStream_t st;
    int engine = -1;          // DICP_ENGINE_*: -1 = the process-wide default
    int sym_mode() const { return engine >= 0 ? engine : ::sym_mode(); }
    template <class Op>
    int run(const typename Op::Params& prm, int M, int N, float* scal_out, int accumulate) {
        return run_pair<Op>(prm, M, N, scal_out, accumulate, ws, wsb, st);
    }
    // (q,q) adjoint passes through the symmetric engine (sym_engine.cuh): every unordered pair once
    bool use_sym(int M) const {
        if (sym_mode() == 0) return false;
        if (sym_applicable(M)) return true;
        // beyond 65536 points: super-blocks (run_pair_sym_blocked), if the workspace holds one tile's partials
        return M > kSymMaxBlocks * kSymRows && ws != nullptr && wsb >= sym_blocked_workspace_bytes(device_info().sms);
    }
    bool use_sym_forward(int M) const { return sym_mode() == 2 && sym_applicable(M); }
    // (x,q) adjoint: both sides from ONE ring pass (rect_pair_kernel) when the sets are large enough and the workspace fits
    bool use_rect(int Nx, int M) const {
        if (sym_mode() == 0 || M < 256 || Nx < 2048) return false;
        const RectPlan p = rect_make_plan(Nx, M, 4, device_info().sms);
        return ws != nullptr && rect_workspace_bytes(p, 8, 4, 8) <= wsb;
    }
    template <class Op>
    int run_rect(const typename Op::Params& prm, int Nx, int M) { return run_pair_rect<Op>(prm, Nx, M, ws, wsb, st); }
    template <class Op>
    int run_sym(const typename Op::Params& prm, int M, float* scal_out) {
        if constexpr (Op::NSCAL == 0) {
            if (!sym_applicable(M)) return run_pair_sym_blocked<Op>(prm, M, ws, wsb, st);
        }
        return run_pair_sym<Op>(prm, M, scal_out, ws, wsb, st);
    }
    void scal_fix(float* scal, float eta, int withdiv) {
        rhs_scal_fix_kernel<<<1, 32, 0, st>>>(scal, eta, withdiv);
        launch_counter() += 1;
    }
};

inline int last_error(int rc) {
    if (rc != DICP_OK) return rc;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DICP_OK : (int)e;
}

// log-responsibilities lgamma_nc = log_softmax_c(w_c - |x_n - mu_c|^2 / (2 sigma^2)) and their arg max
// (GaussianMixtureUnif.log_responsibilities, core/GMM.py:221-232; plot_bis :677-680).  One thread per point; the
// distance is evaluated un-fused in the reference's operation order so that arg max decisions agree on tie-free rows.
template <int D>
__global__ void log_resp_kernel(const float* __restrict__ X, int N, const float* __restrict__ mu,
                                const float* __restrict__ w, int C, float den, float* __restrict__ lgam,
                                long long* __restrict__ amax) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float x[D];
#pragma unroll
    for (int k = 0; k < D; ++k) x[k] = X[(size_t)n * D + k];
    float best = -INFINITY, m = -INFINITY, S = 0.f;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
        float d2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float dk = __fsub_rn(x[k], mu[(size_t)c * D + k]);
            d2 = __fadd_rn(d2, __fmul_rn(dk, dk));
        }
        const float t = __fsub_rn(w[c], __fdiv_rn(d2, den));
        if (t > best) { best = t; bi = c; }
        if (t > m) { S = S * __expf(m - t) + 1.f; m = t; }
        else S += __expf(t - m);
    }
    if (amax) amax[n] = bi;
    if (lgam) {
        const float T = m + __logf(S);
        for (int c = 0; c < C; ++c) {
            float d2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float dk = __fsub_rn(x[k], mu[(size_t)c * D + k]);
                d2 = __fadd_rn(d2, __fmul_rn(dk, dk));
            }
            lgam[(size_t)n * C + c] = __fsub_rn(w[c], __fdiv_rn(d2, den)) - T;
        }
    }
}

// ---- pipe probes ---------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) probe_kernel(int iters, float* out) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 1.0f + 1e-3f * (threadIdx.x + k);
    const float m = 0.9999f, c = 1e-4f;
    if (WHICH == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], m, c);
        }
    } else if (WHICH == 1) {
        unsigned long long w[4], mm, cc;
        float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
        mm = *reinterpret_cast<unsigned long long*>(&m2);
        cc = *reinterpret_cast<unsigned long long*>(&c2);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 t = make_float2(v[2 * k], v[2 * k + 1]);
            w[k] = *reinterpret_cast<unsigned long long*>(&t);
        }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(w[k]) : "l"(w[k]), "l"(mm), "l"(cc));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 t = *reinterpret_cast<float2*>(&w[k]);
            v[2 * k] = t.x; v[2 * k + 1] = t.y;
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ex2_neg(v[k]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// launch the instantiation selected by (D, withlogdet, eta != 0) with `smem` bytes of dynamic shared memory
template <int DD, bool W, bool E, bool BIG>
static void launch_small_rhs_v(const SmallStep& S, int xpass, dim3 grid, size_t smem, cudaStream_t st) {
    static const bool optin = [] {          // static + dynamic shared memory may exceed the 48 KB default: opt in once
        return cudaFuncSetAttribute(small_rhs_step_kernel<DD, W, E, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)small_fwd_smem_bytes(kSmallMaxQ, DD)) == cudaSuccess;
    }();
    (void)optin;
    small_rhs_step_kernel<DD, W, E, BIG><<<grid, kSmallThreads, smem, st>>>(S, xpass);
}
// supports beyond the ring form's 64 points: the instantiation with 128 registers per thread (small_step.cuh: kSmallMinbBig)
static bool small_big(long long maxM) {
    static const long long minM = [] {       // DICP_SMALL_BIG_MIN: tuning sweeps only
        const char* e = getenv("DICP_SMALL_BIG_MIN");
        return e ? atoll(e) : (long long)kRingMaxQ + 1;
    }();
    return maxM >= minM;
}
template <int DD, bool W, bool E>
static void launch_small_rhs(const SmallStep& S, int xpass, dim3 grid, size_t smem, cudaStream_t st) {
    if (small_big(S.M)) launch_small_rhs_v<DD, W, E, true>(S, xpass, grid, smem, st);
    else launch_small_rhs_v<DD, W, E, false>(S, xpass, grid, smem, st);
}
template <int DD, bool W, bool E>
static void launch_small_adj(const SmallStep& S, int nsplit, int xpass, dim3 grid, size_t smem, cudaStream_t st) {
    static const bool optin = [] {
        return cudaFuncSetAttribute(small_adj_step_kernel<DD, W, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)small_adj_smem_bytes(kSmallMaxQ, DD)) == cudaSuccess;
    }();
    (void)optin;
    small_adj_step_kernel<DD, W, E><<<grid, kSmallThreads, smem, st>>>(S, nsplit, xpass);
}
static void dispatch_small_rhs(int D, int withlogdet, float eta, const SmallStep& S, int xpass, dim3 grid, long long maxM,
                               cudaStream_t st) {
    const size_t smem = small_fwd_smem_bytes(maxM, D);
    if (D == 2) {
        if (eta != 0.f) launch_small_rhs<2, true, true>(S, xpass, grid, smem, st);
        else if (withlogdet) launch_small_rhs<2, true, false>(S, xpass, grid, smem, st);
        else launch_small_rhs<2, false, false>(S, xpass, grid, smem, st);
    } else {
        if (eta != 0.f) launch_small_rhs<3, true, true>(S, xpass, grid, smem, st);
        else if (withlogdet) launch_small_rhs<3, true, false>(S, xpass, grid, smem, st);
        else launch_small_rhs<3, false, false>(S, xpass, grid, smem, st);
    }
    launch_counter() += 1;
}
// ring form of the adjoint stage (small_adj_ring_kernel): data points present, at most kRingMaxQ support points
static bool small_ring_applicable(float, long long maxM, long long maxNx) {
    return sym_mode() != 0 && maxNx > 0 && maxM <= kRingMaxQ;
}
template <int DD, bool W, bool E>
static void launch_small_ring(const SmallStep& S, long long maxNx, unsigned frames, cudaStream_t st) {
    // One block of kRingRows rows per x CTA.  (Measured on B200, 64 frames x 10k points: giving every CTA two blocks so that
    // all CTAs are resident in one wave is 6 % SLOWER -- the block scheduler back-fills the short CTAs well -- so the kernel's
    // `xpass` stays 1.)
    const int xpass = 1;
    const long long nx1 = (maxNx + kRingRows - 1) / kRingRows;
    const dim3 grid((unsigned)((nx1 + xpass - 1) / xpass) + 1u, frames);
    small_adj_ring_kernel<DD, W, E><<<grid, kSmallThreads, 0, st>>>(S, xpass);
}
static void dispatch_small_ring(int D, int withlogdet, float eta, const SmallStep& S, long long maxNx, unsigned frames,
                                cudaStream_t st) {
#define DICP_LAUNCH(DD, W, E) launch_small_ring<DD, W, E>(S, maxNx, frames, st)
    if (D == 2) { if (eta != 0.f) DICP_LAUNCH(2, true, true); else if (withlogdet) DICP_LAUNCH(2, true, false); else DICP_LAUNCH(2, false, false); }
    else { if (eta != 0.f) DICP_LAUNCH(3, true, true); else if (withlogdet) DICP_LAUNCH(3, true, false); else DICP_LAUNCH(3, false, false); }
#undef DICP_LAUNCH
    launch_counter() += 1;
}
// mid form of the adjoint stage (small_adj_mid_kernel + small_mid_finish_kernel): more than kRingMaxQ support points and
// enough data points that the 512-row x CTAs of all frames fill the SMs at least once
static bool small_mid_applicable(long long maxM, long long maxNx, long long frames) {
    // DICP_SMALL_MID = 0 / 1: never / whenever it can run (tests, tuning sweeps; read at every call so that a test can switch it)
    const char* e = getenv("DICP_SMALL_MID");
    const int forced = e ? atoi(e) : -1;
    if (sym_mode() == 0 || maxM <= kRingMaxQ || maxNx <= 0 || forced == 0) return false;
    return forced == 1 || frames * ((maxNx + kRingRows - 1) / kRingRows) >= device_info().sms;
}
// data points per lane of an x CTA: 4 (fewest shuffles per pair) or 2, by the fill of the grid's last wave (small_rows_by_waves)
static int small_mid_rows_per_lane(long long maxM, long long maxNx, long long frames) {
    const char* e = getenv("DICP_SMALL_MID_R");      // 2 / 4: tests and tuning sweeps (read at every call)
    const int forced = e ? atoi(e) : 0;
    if (forced == 2 || forced == 4) return forced;
    return small_rows_by_waves(frames, maxNx, frames * ((maxM + kSmallThreads - 1) / kSmallThreads), device_info().sms, 2);
}
template <int DD, bool W, bool E>
static void launch_small_mid(const SmallStep& S, long long maxM, long long maxNx, unsigned frames, cudaStream_t st) {
    const int R = small_mid_rows_per_lane(maxM, maxNx, frames), rows = kSmallThreads * R;
    const dim3 grid((unsigned)((maxNx + rows - 1) / rows + (maxM + kSmallThreads - 1) / kSmallThreads), frames);
    if (R == 4) small_adj_mid_kernel<DD, W, E, 4><<<grid, kSmallThreads, 0, st>>>(S);
    else small_adj_mid_kernel<DD, W, E, 2><<<grid, kSmallThreads, 0, st>>>(S);
    const dim3 gfin((unsigned)((maxM + 31) / 32), frames);
    small_mid_finish_kernel<DD, W, E><<<gfin, kSmallThreads, 0, st>>>(S, rows);
}
static void dispatch_small_mid(int D, int withlogdet, float eta, const SmallStep& S, long long maxM, long long maxNx,
                               unsigned frames, cudaStream_t st) {
#define DICP_LAUNCH(DD, W, E) launch_small_mid<DD, W, E>(S, maxM, maxNx, frames, st)
    if (D == 2) { if (eta != 0.f) DICP_LAUNCH(2, true, true); else if (withlogdet) DICP_LAUNCH(2, true, false); else DICP_LAUNCH(2, false, false); }
    else { if (eta != 0.f) DICP_LAUNCH(3, true, true); else if (withlogdet) DICP_LAUNCH(3, true, false); else DICP_LAUNCH(3, false, false); }
#undef DICP_LAUNCH
    launch_counter() += 2;
}
static void dispatch_small_adj(int D, int withlogdet, float eta, const SmallStep& S, int nsplit, int xpass, dim3 grid,
                               long long maxM, cudaStream_t st) {
    const size_t smem = small_adj_smem_bytes(maxM, D);
    if (D == 2) {
        if (eta != 0.f) launch_small_adj<2, true, true>(S, nsplit, xpass, grid, smem, st);
        else if (withlogdet) launch_small_adj<2, true, false>(S, nsplit, xpass, grid, smem, st);
        else launch_small_adj<2, false, false>(S, nsplit, xpass, grid, smem, st);
    } else {
        if (eta != 0.f) launch_small_adj<3, true, true>(S, nsplit, xpass, grid, smem, st);
        else if (withlogdet) launch_small_adj<3, true, false>(S, nsplit, xpass, grid, smem, st);
        else launch_small_adj<3, false, false>(S, nsplit, xpass, grid, smem, st);
    }
    launch_counter() += 1;
}


}  // namespace

extern "C" {

int dicp_version(void) { return 100; }

int dicp_sm_count(void) { return device_info().sms; }

int dicp_sym_mode(int mode) {
    const int prev = sym_mode_ref().load();
    if (mode >= 0) sym_mode_ref().store(mode);
    return prev;
}

unsigned long long dicp_launch_count(void) { return launch_counter(); }

size_t dicp_pair_workspace_bytes(int64_t rows, int64_t cols) {
    size_t b = pair_workspace_bound(rows, cols);
    if (rows == cols && rows > (int64_t)kSymMaxBlocks * kSymRows) {      // blocked symmetric adjoint: one super-block tile
        const size_t s = sym_blocked_workspace_bytes(device_info().sms);
        if (s > b) b = s;
    }
    if (rows == cols && sym_applicable(rows)) {          // symmetric (q,q) adjoint: packed columns + row / column partials
        const size_t col = align_up((size_t)(rows + 256) * kMaxColF4 * 16, 256);
        const size_t s = col + sym_workspace_bound(rows, device_info().sms) + 1024;
        if (s > b) b = s;
    }
    return b;
}

int dicp_ksum(int D, unsigned mask, float sigma, const float* x, int64_t M, const float* y, int64_t N,
              const float* b, const float* c, const float* d,
              float* o_base, float* o_redscal, float* o_red, float* o_grad, float* o_dd, float* o_gend,
              float* o_hess, float* o_lap, float* o_gradlap, float* o_minsq, float* o_dot,
              void* workspace, size_t workspace_bytes, void* stream) {
    float* outs[11] = {o_base, o_redscal, o_red, o_grad, o_dd, o_gend, o_hess, o_lap, o_gradlap, o_minsq, o_dot};
    DeviceExec ex{workspace, workspace_bytes, (cudaStream_t)stream};
    return last_error(ksum_entry(ex, D, mask, sigma, x, M, y, N, b, c, d, outs));
}

int dicp_rhs_forward(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                     const float* x, int64_t Nx, float* vq, float* dp, float* vx, float* scal,
                     void* workspace, size_t workspace_bytes, void* stream, int engine) {
    if (engine < DICP_ENGINE_DEFAULT || engine > DICP_ENGINE_SYMMETRIC_ALL) return DICP_EBADARG;
    DeviceExec ex{workspace, workspace_bytes, (cudaStream_t)stream, engine};
    return last_error(rhs_forward_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, vq, dp, vx, scal));
}

int dicp_rhs_adjoint(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                     const float* x, int64_t Nx, const float* a, const float* u, const float* wx, const float* gc,
                     float* gq, float* gp, float* gx, void* workspace, size_t workspace_bytes, void* stream, int engine) {
    if (engine < DICP_ENGINE_DEFAULT || engine > DICP_ENGINE_SYMMETRIC_ALL) return DICP_EBADARG;
    DeviceExec ex{workspace, workspace_bytes, (cudaStream_t)stream, engine};
    return last_error(rhs_adjoint_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, a, u, wx, gc, gq, gp, gx));
}

int dicp_em_rowpass(int D, int lite, float sigma_old, const float* X, int64_t N, const float* mu_old,
                    const float* wl2, int64_t C, const float* mu_new, const float* lpi_new, float* T2, float* Y,
                    float* rowP, float* rowQ, float* sq, float* scal4, void* workspace, size_t workspace_bytes,
                    void* stream) {
    if (C >= 1 && C <= kEmColMaxC && (D == 2 || D == 3) && sigma_old > 0.f && N >= 1 && N <= INT32_MAX && X &&
        mu_old && wl2 && T2 && (lite || (mu_new && lpi_new && Y && scal4)) && workspace &&
        workspace_bytes >= em_row_small_workspace(N)) {
        // few components: one launch, components resident in shared memory (em_col_small.cuh)
        EmParams prm{};
        prm.X = X; prm.mu_old = mu_old; prm.wl2 = wl2; prm.mu_new = mu_new; prm.lpi_new = lpi_new; prm.origin = mu_old;
        prm.kappa = gauss_const(sigma_old).kappa;
        prm.o_T2 = T2; prm.o_Y = Y; prm.o_rowP = rowP; prm.o_rowQ = rowQ; prm.o_sq = sq;
        cudaStream_t st = (cudaStream_t)stream;
        float* blockscal = (float*)((char*)workspace + 256);
        const long long groups = (N + kEmRowRows - 1) / kEmRowRows;
        long long passes = groups / ((long long)device_info().sms * 8);          // about 8 CTAs per SM, at most 16 row groups each
        if (passes < 1) passes = 1;
        if (passes > 16) passes = 16;
        const unsigned blocks = (unsigned)((groups + passes - 1) / passes);
        if (D == 2) {
            if (lite) em_row_small_kernel<2, true><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, (int)passes, blockscal);
            else em_row_small_kernel<2, false><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, (int)passes, blockscal);
        } else {
            if (lite) em_row_small_kernel<3, true><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, (int)passes, blockscal);
            else em_row_small_kernel<3, false><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, (int)passes, blockscal);
        }
        if (!lite) {
            scalar_reduce_kernel<<<1, 256, 0, st>>>(blockscal, (int)blocks, 4, scal4, 0);
            launch_counter() += 1;
        }
        launch_counter() += 1;
        return last_error(DICP_OK);
    }
    DeviceExec ex{workspace, workspace_bytes, (cudaStream_t)stream};
    return last_error(em_rowpass_entry(ex, D, lite, sigma_old, X, N, mu_old, wl2, C, mu_new, lpi_new, T2, Y, rowP, rowQ,
                                       sq, scal4));
}

int dicp_em_colstats(int D, float sigma_old, const float* X, int64_t N, const float* T2, const float* mu_old,
                     const float* wl2, int64_t C, float* stats, void* workspace, size_t workspace_bytes, void* stream) {
    if (C >= 1 && C <= kEmColMaxC && (D == 2 || D == 3) && sigma_old > 0.f && N >= 1 && N <= INT32_MAX && X && T2 && mu_old &&
        wl2 && stats && workspace && workspace_bytes >= em_col_small_workspace(N, (int)C, device_info().sms)) {
        // few components: dedicated one-launch kernel with all lanes busy (em_col_small.cuh)
        EmParams prm{};
        prm.X = X; prm.T2 = T2; prm.mu_old = mu_old; prm.wl2 = wl2; prm.origin = mu_old;
        prm.kappa = gauss_const(sigma_old).kappa;
        prm.o_stats = stats;
        cudaStream_t st = (cudaStream_t)stream;
        unsigned* counter = (unsigned*)workspace;
        const size_t cbytes = em_col_small_counter_bytes(N, device_info().sms);
        float* part = (float*)((char*)workspace + cbytes);
        cudaMemsetAsync(counter, 0, cbytes, st);
        const int nsplit = em_col_small_splits(N, device_info().sms);
        if (D == 2) em_col_small_kernel<2><<<nsplit, 128, 0, st>>>(prm, (int)N, (int)C, part, counter);
        else em_col_small_kernel<3><<<nsplit, 128, 0, st>>>(prm, (int)N, (int)C, part, counter);
        launch_counter() += 1;
        return last_error(DICP_OK);
    }
    DeviceExec ex{workspace, workspace_bytes, (cudaStream_t)stream};
    return last_error(em_colstats_entry(ex, D, sigma_old, X, N, T2, mu_old, wl2, C, stats));
}

int dicp_em_lse_colstats(int D, float sigma_old, const float* X, int64_t N, const float* mu_old, const float* wl2, int64_t C,
                         float* T2_scratch, float* stats, void* workspace, size_t workspace_bytes, void* stream) {
    if (C >= 1 && C <= kEmColMaxC && (D == 2 || D == 3) && sigma_old > 0.f && N >= 1 && N <= INT32_MAX && X && mu_old &&
        wl2 && stats && workspace && workspace_bytes >= em_col_small_workspace(N, (int)C, device_info().sms)) {
        // few components: row log-sum-exp and column statistics from ONE read of X (em_lse_col_small_kernel)
        EmParams prm{};
        prm.X = X; prm.mu_old = mu_old; prm.wl2 = wl2; prm.origin = mu_old;
        prm.kappa = gauss_const(sigma_old).kappa;
        prm.o_stats = stats;
        cudaStream_t st = (cudaStream_t)stream;
        unsigned* counter = (unsigned*)workspace;
        const size_t cbytes = em_col_small_counter_bytes(N, device_info().sms);
        float* part = (float*)((char*)workspace + cbytes);
        cudaMemsetAsync(counter, 0, cbytes, st);
        int blocks = 1, passes = 1;
        em_lse_col_small_grid(N, device_info().sms, &blocks, &passes);
        if (D == 2) em_lse_col_small_kernel<2><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, passes, part, counter);
        else em_lse_col_small_kernel<3><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, passes, part, counter);
        launch_counter() += 1;
        return last_error(DICP_OK);
    }
    // many components: the two sweeps of the general engine, T2 through the caller's scratch array
    if (!T2_scratch) return DICP_EBADARG;
    const int rc = dicp_em_rowpass(D, 1, sigma_old, X, N, mu_old, wl2, C, nullptr, nullptr, T2_scratch, nullptr, nullptr,
                                   nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
    if (rc != DICP_OK) return rc;
    return dicp_em_colstats(D, sigma_old, X, N, T2_scratch, mu_old, wl2, C, stats, workspace, workspace_bytes, stream);
}

int dicp_em_mstep(int D, const float* stats, const float* mu_old, const float* w_old, int64_t C, int do_mu, int do_w,
                  int sig_mode, float* mu_new, float* w_new, float* lpi_new, float* out_scal, void* stream) {
    if ((D != 2 && D != 3) || C < 1 || C > INT32_MAX || sig_mode < 0 || sig_mode > 2 || !mu_old || !w_old || !mu_new ||
        !w_new || !lpi_new || !out_scal || ((do_mu || do_w || sig_mode) && !stats))
        return DICP_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) em_mstep_kernel<2><<<1, 256, 0, st>>>(stats, mu_old, w_old, (int)C, do_mu, do_w, sig_mode, mu_new, w_new, lpi_new, out_scal);
    else em_mstep_kernel<3><<<1, 256, 0, st>>>(stats, mu_old, w_old, (int)C, do_mu, do_w, sig_mode, mu_new, w_new, lpi_new, out_scal);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

size_t dicp_em_state_workspace_bytes(int64_t N, int64_t C) {
    if (N < 1 || C < 1 || C > kEmColMaxC) return 0;
    const size_t a = (em_col_small_workspace(N, (int)C, device_info().sms) + 255) / 256 * 256;
    return a + em_row_small_workspace(N);
}

static int em_state_step_launch(int D, const float* X, int64_t N, int64_t C, float* mu, float* w, float* lpi, float* wl2,
                                float* mu_new, float* w_new, float* lpi_new, float* stats, float* Y, float* T2, float* scal4,
                                double* state, int do_mu, int do_w, int sig_mode, int keops_sem, void* workspace,
                                size_t workspace_bytes, cudaStream_t st, int use_cond, cudaGraphConditionalHandle cond) {
    if ((D != 2 && D != 3) || N < 1 || N > INT32_MAX || C < 1 || C > kEmColMaxC || sig_mode < 0 || sig_mode > 2 || !X || !mu ||
        !w || !lpi || !wl2 || !mu_new || !w_new || !lpi_new || !stats || !Y || !T2 || !scal4 || !state || !workspace)
        return DICP_EBADARG;
    if (workspace_bytes < dicp_em_state_workspace_bytes(N, C)) return DICP_EWORKSPACE;
    const int sms = device_info().sms;
    // (1) row log-sum-exp + column statistics, old parameters
    EmParams prm{};
    prm.X = X; prm.mu_old = mu; prm.wl2 = wl2; prm.origin = mu; prm.o_stats = stats; prm.state = state;
    unsigned* counter = (unsigned*)workspace;
    const size_t cbytes = em_col_small_counter_bytes(N, sms);
    float* part = (float*)((char*)workspace + cbytes);
    cudaMemsetAsync(counter, 0, cbytes, st);
    int blocks = 1, passes = 1;
    em_lse_col_small_grid(N, sms, &blocks, &passes);
    if (D == 2) em_lse_col_small_kernel<2><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, passes, part, counter);
    else em_lse_col_small_kernel<3><<<blocks, 128, 0, st>>>(prm, (int)N, (int)C, passes, part, counter);
    // (2) M step; sigma' into the state
    if (D == 2) em_mstep_state_kernel<2><<<1, 256, 0, st>>>(stats, mu, w, (int)C, do_mu, do_w, sig_mode, mu_new, w_new, lpi_new, state);
    else em_mstep_state_kernel<3><<<1, 256, 0, st>>>(stats, mu, w, (int)C, do_mu, do_w, sig_mode, mu_new, w_new, lpi_new, state);
    // (3) full row pass: old responsibilities, new centroids / weights
    EmParams row{};
    row.X = X; row.mu_old = mu; row.wl2 = wl2; row.mu_new = mu_new; row.lpi_new = lpi_new; row.origin = mu;
    row.o_T2 = T2; row.o_Y = Y; row.state = state;
    float* blockscal = (float*)((char*)workspace + (em_col_small_workspace(N, (int)C, sms) + 255) / 256 * 256 + 256);
    const long long groups = (N + kEmRowRows - 1) / kEmRowRows;
    long long rp = groups / ((long long)sms * 8);
    if (rp < 1) rp = 1;
    if (rp > 16) rp = 16;
    const unsigned rblocks = (unsigned)((groups + rp - 1) / rp);
    if (D == 2) em_row_small_kernel<2, false><<<rblocks, 128, 0, st>>>(row, (int)N, (int)C, (int)rp, blockscal);
    else em_row_small_kernel<3, false><<<rblocks, 128, 0, st>>>(row, (int)N, (int)C, (int)rp, blockscal);
    scalar_reduce_kernel<<<1, 256, 0, st>>>(blockscal, (int)rblocks, 4, scal4, 0);
    // (4) free energy, stop test, commit of the new parameters (and, inside a WHILE node, the loop condition)
    if (D == 2) em_state_finalize_kernel<2><<<1, 256, 0, st>>>(scal4, (int)C, keops_sem, mu_new, w_new, lpi_new, mu, w, lpi, wl2, state, use_cond, cond);
    else em_state_finalize_kernel<3><<<1, 256, 0, st>>>(scal4, (int)C, keops_sem, mu_new, w_new, lpi_new, mu, w, lpi, wl2, state, use_cond, cond);
    return DICP_OK;
}

int dicp_em_state_step(int D, const float* X, int64_t N, int64_t C, float* mu, float* w, float* lpi, float* wl2, float* mu_new,
                       float* w_new, float* lpi_new, float* stats, float* Y, float* T2, float* scal4, double* state, int do_mu,
                       int do_w, int sig_mode, int keops_sem, void* workspace, size_t workspace_bytes, void* stream) {
    const int rc = em_state_step_launch(D, X, N, C, mu, w, lpi, wl2, mu_new, w_new, lpi_new, stats, Y, T2, scal4, state, do_mu,
                                        do_w, sig_mode, keops_sem, workspace, workspace_bytes, (cudaStream_t)stream, 0,
                                        cudaGraphConditionalHandle{});
    if (rc != DICP_OK) return rc;
    launch_counter() += 5;
    return last_error(DICP_OK);
}

// ---- the EM loop as ONE graph launch: a WHILE conditional node whose body is one step -------------------------------------
struct EmLoopGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

void* dicp_em_loop_create(int D, const float* X, int64_t N, int64_t C, float* mu, float* w, float* lpi, float* wl2, float* mu_new,
                          float* w_new, float* lpi_new, float* stats, float* Y, float* T2, float* scal4, double* state, int do_mu,
                          int do_w, int sig_mode, int keops_sem, void* workspace, size_t workspace_bytes, void* stream) {
    EmLoopGraph* L = new EmLoopGraph();
    cudaGraphConditionalHandle cond;
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    cudaGraphNode_t node;
    (void)stream;
    cudaStream_t st = nullptr;             // capture needs a stream of its own (the caller's may be the legacy default stream)
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { delete L; return nullptr; }
    bool ok = cudaGraphCreate(&L->graph, 0) == cudaSuccess &&
              cudaGraphConditionalHandleCreate(&cond, L->graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) {
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        ok = cudaGraphAddNode(&node, L->graph, nullptr, 0, &np) == cudaSuccess;
    }
    if (ok) ok = cudaStreamBeginCaptureToGraph(st, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                               cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int rc = em_state_step_launch(D, X, N, C, mu, w, lpi, wl2, mu_new, w_new, lpi_new, stats, Y, T2, scal4, state,
                                            do_mu, do_w, sig_mode, keops_sem, workspace, workspace_bytes, st, 1, cond);
        cudaGraph_t captured = nullptr;
        ok = cudaStreamEndCapture(st, &captured) == cudaSuccess && rc == DICP_OK;
    }
    if (ok) ok = cudaGraphInstantiate(&L->exec, L->graph, 0) == cudaSuccess;
    cudaStreamDestroy(st);
    if (!ok) {
        fprintf(stderr, "dicp_em_loop_create: %s\n", cudaGetErrorString(cudaPeekAtLastError()));
        cudaGetLastError();
        if (L->graph) cudaGraphDestroy(L->graph);
        delete L;
        return nullptr;
    }
    return L;
}

int dicp_em_loop_launch(void* loop, void* stream) {
    EmLoopGraph* L = (EmLoopGraph*)loop;
    if (!L || !L->exec) return DICP_EBADARG;
    const cudaError_t e = cudaGraphLaunch(L->exec, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    launch_counter() += 1;
    return DICP_OK;
}

void dicp_em_loop_destroy(void* loop) {
    EmLoopGraph* L = (EmLoopGraph*)loop;
    if (!L) return;
    if (L->exec) cudaGraphExecDestroy(L->exec);
    if (L->graph) cudaGraphDestroy(L->graph);
    delete L;
}

int dicp_em_reduce_pack(int D, const float* stats, const float* m_ref, int64_t C, const float* extra, int n_extra, float* buf,
                        void* stream) {
    if ((D != 2 && D != 3) || C < 1 || C > INT32_MAX || !stats || !m_ref || !buf || n_extra < 0 || (n_extra > 0 && !extra))
        return DICP_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) em_reduce_pack_kernel<2><<<1, 256, 0, st>>>(stats, m_ref, (int)C, extra, n_extra, buf);
    else em_reduce_pack_kernel<3><<<1, 256, 0, st>>>(stats, m_ref, (int)C, extra, n_extra, buf);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

int dicp_em_mstep_merged(int D, const float* buf, const float* m_ref, const float* mu_old, const float* w_old, int64_t C,
                         int do_mu, int do_w, int sig_mode, int n_extra, float* mu_new, float* w_new, float* lpi_new,
                         float* m_next, float* host, void* stream) {
    if ((D != 2 && D != 3) || C < 1 || C > INT32_MAX || sig_mode < 0 || sig_mode > 2 || n_extra < 0 || !buf || !m_ref ||
        !mu_old || !w_old || !mu_new || !w_new || !lpi_new || !m_next || !host)
        return DICP_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) em_mstep_merged_kernel<2><<<1, 256, 0, st>>>(buf, m_ref, mu_old, w_old, (int)C, do_mu, do_w, sig_mode, n_extra,
                                                            mu_new, w_new, lpi_new, m_next, host);
    else em_mstep_merged_kernel<3><<<1, 256, 0, st>>>(buf, m_ref, mu_old, w_old, (int)C, do_mu, do_w, sig_mode, n_extra,
                                                     mu_new, w_new, lpi_new, m_next, host);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

int dicp_log_resp(int D, float sigma, const float* X, int64_t N, const float* mu, const float* w, int64_t C,
                  float* lgam, long long* argmax, void* stream) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || N < 0 || C < 1 || N > INT32_MAX || C > INT32_MAX) return DICP_EBADARG;
    if (N == 0) return DICP_OK;
    if (!X || !mu || !w || (!lgam && !argmax)) return DICP_EBADARG;
    const float den = 2.f * sigma * sigma;
    const int blocks = (int)((N + 127) / 128);
    if (D == 2) log_resp_kernel<2><<<blocks, 128, 0, (cudaStream_t)stream>>>(X, (int)N, mu, w, (int)C, den, lgam, argmax);
    else log_resp_kernel<3><<<blocks, 128, 0, (cudaStream_t)stream>>>(X, (int)N, mu, w, (int)C, den, lgam, argmax);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

int dicp_small_max_support(void) { return kSmallMaxQ; }

size_t dicp_small_workspace_bytes(int64_t M, int64_t Nx) { return small_workspace_bytes(M, Nx); }

static int small_fill(SmallStep& S, int D, float sigma, float eta, int64_t M, int64_t Nx, void* ws, size_t wsb) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || M < 1 || M > kSmallMaxQ || Nx < 0 || Nx > INT32_MAX) return DICP_EBADARG;
    if (ws == nullptr || wsb < small_workspace_bytes(M, Nx)) return DICP_EWORKSPACE;
    GaussConst g = gauss_const(sigma);
    S.M = (int)M; S.Nx = (int)Nx;
    S.kappa = g.kappa; S.s = g.s; S.alpha = g.alpha; S.beta = g.beta; S.eta = eta;
    S.counters = (unsigned*)ws;
    S.ws = (float*)((char*)ws + kSmallCounters * 4);
    return DICP_OK;
}

int dicp_small_rhs_step(int D, int withlogdet, float sigma, float eta, int64_t M, int64_t Nx, const float* s_eval,
                        const float* base, const float* other, float c_this, float c_other, float* out, float* F,
                        void* workspace, size_t workspace_bytes, void* stream) {
    SmallStep S{};
    int rc = small_fill(S, D, sigma, eta, M, Nx, workspace, workspace_bytes);
    if (rc != DICP_OK) return rc;
    if (!s_eval || !F || (out && !base) || (eta != 0.f && !withlogdet)) return DICP_EBADARG;
    S.s_eval = s_eval; S.base = base; S.other = other; S.out = out; S.This = F; S.c_this = c_this; S.c_other = c_other;
    const int xpass = small_big(M) ? small_xpass_big(1, Nx, M, device_info().sms) : small_xpass(1, Nx, device_info().sms);
    const unsigned grid = (unsigned)((Nx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass) + (M + kSmallThreads - 1) / kSmallThreads);
    cudaStream_t st = (cudaStream_t)stream;
    dispatch_small_rhs(D, withlogdet, eta, S, xpass, dim3(grid), M, st);
    return last_error(DICP_OK);
}

int dicp_small_adj_step(int D, int withlogdet, float sigma, float eta, int64_t M, int64_t Nx, const float* s_eval,
                        const float* lam, const float* base, const float* other, const float* add, float c_this,
                        float c_other, float* out, float* G, void* workspace, size_t workspace_bytes, void* stream) {
    SmallStep S{};
    int rc = small_fill(S, D, sigma, eta, M, Nx, workspace, workspace_bytes);
    if (rc != DICP_OK) return rc;
    if (!s_eval || !lam || !G || (out && !base) || out == lam || (eta != 0.f && !withlogdet)) return DICP_EBADARG;
    S.s_eval = s_eval; S.lam = lam; S.base = base; S.other = other; S.add = add; S.out = out; S.This = G;
    S.c_this = c_this; S.c_other = c_other;
    const int nsplit = small_adj_nsplit((int)Nx, small_adj_groups((int)M));
    const unsigned nQB = (unsigned)((M + kSmallThreads - 1) / kSmallThreads);
    if (1 + nQB > (unsigned)kSmallCounters) return DICP_EBADARG;
    const int xpass = small_xpass(1, Nx, device_info().sms);
    const unsigned grid = (unsigned)((Nx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass)) + nQB * (unsigned)nsplit;
    cudaStream_t st = (cudaStream_t)stream;
    if (small_ring_applicable(eta, M, Nx)) dispatch_small_ring(D, withlogdet, eta, S, Nx, 1u, st);
    else if (small_mid_applicable(M, Nx, 1)) dispatch_small_mid(D, withlogdet, eta, S, M, Nx, 1u, st);
    else dispatch_small_adj(D, withlogdet, eta, S, nsplit, xpass, dim3(grid), M, st);
    return last_error(DICP_OK);
}

// ---- batched (multi-frame) closure for small supports ----------------------------------------------------------------
static int batch_fill(SmallStep& S, int D, float sigma, float eta, int K, const int* dims, const int* active,
                      int64_t maxM, int64_t maxNx, int64_t fstride, void* ws, size_t ws_frame_bytes) {
    if ((D != 2 && D != 3) || !(sigma > 0.f) || K < 1 || K > 65535 || !dims || maxM < 1 || maxM > kSmallMaxQ ||
        maxNx < 0 || maxNx > INT32_MAX || fstride < 2 * maxM * D + maxNx * D + 4)
        return DICP_EBADARG;
    if (ws == nullptr || ws_frame_bytes < small_workspace_bytes(maxM, maxNx) || (ws_frame_bytes & 15)) return DICP_EWORKSPACE;
    GaussConst g = gauss_const(sigma);
    S.M = (int)maxM; S.Nx = (int)maxNx;
    S.kappa = g.kappa; S.s = g.s; S.alpha = g.alpha; S.beta = g.beta; S.eta = eta;
    S.counters = (unsigned*)ws;
    S.ws = (float*)((char*)ws + kSmallCounters * 4);
    S.dims = dims; S.active = active; S.fstride = fstride; S.ws_fstride = (long long)ws_frame_bytes;
    return DICP_OK;
}

int dicp_batch_rhs_step(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                        int64_t maxM, int64_t maxNx, int64_t fstride, const float* s_eval, const float* base,
                        const float* other, float c_this, float c_other, float* out, float* F, void* workspace,
                        size_t ws_frame_bytes, void* stream) {
    SmallStep S{};
    int rc = batch_fill(S, D, sigma, eta, K, dims, active, maxM, maxNx, fstride, workspace, ws_frame_bytes);
    if (rc != DICP_OK) return rc;
    if (!s_eval || !F || (out && !base) || (eta != 0.f && !withlogdet)) return DICP_EBADARG;
    S.s_eval = s_eval; S.base = base; S.other = other; S.out = out; S.This = F; S.c_this = c_this; S.c_other = c_other;
    const int xpass = small_big(maxM) ? small_xpass_big(K, maxNx, maxM, device_info().sms) : small_xpass(K, maxNx, device_info().sms);
    const dim3 grid((unsigned)((maxNx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass) +
                               (maxM + kSmallThreads - 1) / kSmallThreads), (unsigned)K);
    cudaStream_t st = (cudaStream_t)stream;
    dispatch_small_rhs(D, withlogdet, eta, S, xpass, grid, maxM, st);
    return last_error(DICP_OK);
}

int dicp_batch_adj_step(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                        int64_t maxM, int64_t maxNx, int64_t fstride, const float* s_eval, const float* lam,
                        const float* base, const float* other, const float* add, float c_this, float c_other, float* out,
                        float* G, void* workspace, size_t ws_frame_bytes, void* stream) {
    SmallStep S{};
    int rc = batch_fill(S, D, sigma, eta, K, dims, active, maxM, maxNx, fstride, workspace, ws_frame_bytes);
    if (rc != DICP_OK) return rc;
    if (!s_eval || !lam || !G || (out && !base) || out == lam || (eta != 0.f && !withlogdet)) return DICP_EBADARG;
    S.s_eval = s_eval; S.lam = lam; S.base = base; S.other = other; S.add = add; S.out = out; S.This = G;
    S.c_this = c_this; S.c_other = c_other;
    // grid.x bounds every frame's CTA count: x-row CTAs + q-row blocks x splits are both monotone in Nx and M
    const int nsplit = small_adj_nsplit((int)maxNx, small_adj_groups((int)maxM));
    const unsigned nQB = (unsigned)((maxM + kSmallThreads - 1) / kSmallThreads);
    if (1 + nQB > (unsigned)kSmallCounters) return DICP_EBADARG;
    const int xpass = small_xpass(K, maxNx, device_info().sms);
    const dim3 grid((unsigned)((maxNx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass)) + nQB * (unsigned)nsplit,
                    (unsigned)K);
    cudaStream_t st = (cudaStream_t)stream;
    if (small_ring_applicable(eta, maxM, maxNx)) dispatch_small_ring(D, withlogdet, eta, S, maxNx, (unsigned)K, st);
    else if (small_mid_applicable(maxM, maxNx, K)) dispatch_small_mid(D, withlogdet, eta, S, maxM, maxNx, (unsigned)K, st);
    else dispatch_small_adj(D, withlogdet, eta, S, nsplit, xpass, grid, maxM, st);
    return last_error(DICP_OK);
}

static inline bool batch_dims_ok(int D, int K, const int* dims, int64_t fstride) {
    return (D == 2 || D == 3) && K >= 1 && K <= 65535 && dims != nullptr && fstride >= 1;
}

int dicp_batch_set_p(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t fstride, const float* X,
                     int64_t xstride, float* state0, void* stream) {
    if (!batch_dims_ok(D, K, dims, fstride) || !X || !state0 || maxM < 1 || xstride < maxM * D) return DICP_EBADARG;
    BatchDims B{dims, active, fstride};
    const dim3 grid((unsigned)((maxM * D + 127) / 128), (unsigned)K);
    batch_set_p_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(B, D, X, xstride, state0);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

size_t dicp_batch_quad_workspace_bytes(int K) { return (size_t)K * (64 + 1) * 4 + 256; }

int dicp_batch_quad_loss(int D, int K, const int* dims, const int* active, int64_t max_points, int64_t fstride,
                         const float* state_end, const float* y, const float* inv, int64_t ystride, float* g_end,
                         float* loss, int64_t lstride, void* workspace, size_t workspace_bytes, void* stream) {
    if (!batch_dims_ok(D, K, dims, fstride) || !state_end || !y || !inv || !g_end || !loss || max_points < 1 ||
        ystride < max_points || lstride < 1)
        return DICP_EBADARG;
    if (!workspace || workspace_bytes < dicp_batch_quad_workspace_bytes(K)) return DICP_EWORKSPACE;
    BatchDims B{dims, active, fstride};
    long long blocks = (max_points + 255) / 256;
    if (blocks > 64) blocks = 64;
    unsigned* counters = (unsigned*)workspace;                      // K words, zero before the first use
    float* partials = (float*)workspace + ((K + 63) / 64) * 64;   // K x 64 floats
    if ((size_t)(((K + 63) / 64) * 64 + (size_t)K * 64) * 4 > workspace_bytes) return DICP_EWORKSPACE;
    const dim3 grid((unsigned)blocks, (unsigned)K);
    batch_quad_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, D, state_end, y, inv, ystride, g_end, loss, lstride,
                                                                   partials, 64, counters);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

int dicp_batch_closure_out(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t fstride,
                           float lam_reg, const float* lam, const float* F0, const float* state_end, float* out,
                           int64_t ostride, int nscal, void* stream) {
    if (!batch_dims_ok(D, K, dims, fstride) || !F0 || !state_end || !out || nscal < 6 || maxM < 1 ||
        ostride < nscal + (lam ? maxM * D : 0))
        return DICP_EBADARG;
    BatchDims B{dims, active, fstride};
    const dim3 grid((unsigned)((maxM * D + 127) / 128), (unsigned)K);
    batch_closure_out_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(B, D, lam_reg, lam, F0, state_end, out, ostride, nscal);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

// ---- whole closure of every frame in ONE launch (cluster_closure.cuh) ------------------------------------------------------

int dicp_batch_closure_cluster_rows(int D, float eta, int scheme_euler, int64_t maxM, int64_t maxNx, int nt, int K) {
    if ((D != 2 && D != 3) || eta != 0.f || !scheme_euler || maxM < 1 || maxNx < 1 || nt < 1 || K < 1) return 0;
    return cc_pick_shape(D, maxM, maxNx, nt, K).cap;
}

int dicp_batch_closure_cluster(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                               int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                               const float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride,
                               float lam_reg, float* out, int64_t ostride, int nscal, void* stream) {
    if (!batch_dims_ok(D, K, dims, fstride) || !traj || !X || !y || !inv || !out || nscal < 6 || !(sigma > 0.f) ||
        ostride < nscal + maxM * D || xstride < maxM * D || ystride < maxNx)
        return DICP_EBADARG;
    if (dicp_batch_closure_cluster_rows(D, eta, 1, maxM, maxNx, nt, K) == 0) return DICP_EUNSUPPORTED;
    const CcShape sh = cc_pick_shape(D, maxM, maxNx, nt, K);
    const GaussConst gcst = gauss_const(sigma);
    ClusterClosure C{};
    C.dims = dims; C.active = active; C.traj = traj; C.fstride = fstride; C.tstride = tstride;
    C.X = X; C.xstride = xstride; C.y = y; C.inv = inv; C.ystride = ystride;
    C.out = out; C.ostride = ostride; C.ns = nscal; C.nt = nt; C.rows_cap = sh.cap;
    C.h = 1.f / (float)nt; C.kappa = gcst.kappa; C.s = gcst.s; C.alpha = gcst.alpha; C.beta = gcst.beta; C.lam_reg = lam_reg;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (sh.T == 128) {
        if (D == 2) rc = withlogdet ? cc_launch<2, true, 128>(C, sh, K, st) : cc_launch<2, false, 128>(C, sh, K, st);
        else rc = withlogdet ? cc_launch<3, true, 128>(C, sh, K, st) : cc_launch<3, false, 128>(C, sh, K, st);
    } else {
        if (D == 2) rc = withlogdet ? cc_launch<2, true, 64>(C, sh, K, st) : cc_launch<2, false, 64>(C, sh, K, st);
        else rc = withlogdet ? cc_launch<3, true, 64>(C, sh, K, st) : cc_launch<3, false, 64>(C, sh, K, st);
    }
    if (rc != DICP_OK) return rc;
    launch_counter() += 1;
    return last_error(DICP_OK);
}

// ---- lock-step L-BFGS on the device (lbfgs_device.cuh) ----------------------------------------------------------------------
int dicp_lbfgs_dev_begin(const dicp_lbfgs_dev* L, const unsigned char* mask, float* X, int64_t xstride, int* active, void* stream) {
    if (!L || L->K < 1 || !L->ints || !L->dbl || !L->vec || !L->counters || !X || !active || xstride < L->stride) return DICP_EBADARG;
    lbfgs_dev_begin_kernel<<<(L->K + kLdWarps - 1) / kLdWarps, kLdWarps * 32, 0, (cudaStream_t)stream>>>(*L, mask, X, xstride, active);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

static int lbfgs_dev_round_launch(const dicp_lbfgs_dev* L, int D, int withlogdet, float sigma, float eta, int K, const int* dims,
                                  int* active, int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                                  float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride, float lam_reg,
                                  float* out, int64_t ostride, int nscal, cudaStream_t st, int max_rounds, int use_cond,
                                  cudaGraphConditionalHandle cond) {
    if (!L || L->K != K || !L->ints || !L->dbl || !L->vec || !L->best_x || !L->dirs || !L->stps || !L->ro || !L->al ||
        !L->counters || L->history < 1 || L->stride < maxM * D || maxM * D > 32 * kLdNpl || !active)
        return DICP_EBADARG;
    const int rc = dicp_batch_closure_cluster(D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, nt, traj, tstride,
                                              X, xstride, y, inv, ystride, lam_reg, out, ostride, nscal, st);
    if (rc != DICP_OK) return rc;
    lbfgs_dev_feed_kernel<<<(K + kLdWarps - 1) / kLdWarps, kLdWarps * 32, 0, st>>>(*L, out, ostride, nscal, (double)lam_reg,
                                                                                 (double)eta, X, xstride, active, max_rounds,
                                                                                 use_cond, cond);
    launch_counter() += 1;
    const cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? DICP_OK : (int)e;
}

int dicp_lbfgs_dev_round(const dicp_lbfgs_dev* L, int D, int withlogdet, float sigma, float eta, int K, const int* dims,
                         int* active, int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                         float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride, float lam_reg, float* out,
                         int64_t ostride, int nscal, void* stream) {
    return last_error(lbfgs_dev_round_launch(L, D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, nt, traj, tstride,
                                             X, xstride, y, inv, ystride, lam_reg, out, ostride, nscal, (cudaStream_t)stream,
                                             1 << 30, 0, cudaGraphConditionalHandle{}));
}

void* dicp_lbfgs_dev_loop_create(const dicp_lbfgs_dev* L, int D, int withlogdet, float sigma, float eta, int K, const int* dims,
                                 int* active, int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                                 float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride, float lam_reg,
                                 float* out, int64_t ostride, int nscal, int max_rounds) {
    EmLoopGraph* G = new EmLoopGraph();
    cudaGraphConditionalHandle cond;
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    cudaGraphNode_t node;
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { delete G; return nullptr; }
    bool ok = cudaGraphCreate(&G->graph, 0) == cudaSuccess &&
              cudaGraphConditionalHandleCreate(&cond, G->graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    if (ok) {
        np.conditional.handle = cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        ok = cudaGraphAddNode(&node, G->graph, nullptr, 0, &np) == cudaSuccess;
    }
    if (ok) ok = cudaStreamBeginCaptureToGraph(st, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                               cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int rc = lbfgs_dev_round_launch(L, D, withlogdet, sigma, eta, K, dims, active, maxM, maxNx, fstride, nt, traj, tstride,
                                              X, xstride, y, inv, ystride, lam_reg, out, ostride, nscal, st, max_rounds, 1, cond);
        cudaGraph_t captured = nullptr;
        ok = cudaStreamEndCapture(st, &captured) == cudaSuccess && rc == DICP_OK;
    }
    if (ok) ok = cudaGraphInstantiate(&G->exec, G->graph, 0) == cudaSuccess;
    cudaStreamDestroy(st);
    if (!ok) {
        fprintf(stderr, "dicp_lbfgs_dev_loop_create: %s\n", cudaGetErrorString(cudaPeekAtLastError()));
        cudaGetLastError();
        if (G->graph) cudaGraphDestroy(G->graph);
        delete G;
        return nullptr;
    }
    return G;
}

int dicp_lbfgs_dev_loop_launch(void* loop, void* stream) { return dicp_em_loop_launch(loop, stream); }
void dicp_lbfgs_dev_loop_destroy(void* loop) { dicp_em_loop_destroy(loop); }

int dicp_batch_coverage(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t maxNx, int64_t fstride,
                        const float* traj, int64_t tstride, int ntimes, float radius, int* counts, void* stream) {
    if (!batch_dims_ok(D, K, dims, fstride) || !traj || !counts || ntimes < 1 || ntimes > 65535 || maxM < 1 ||
        maxM > kSmallMaxQ || maxNx < 0 || !(radius >= 0.f))
        return DICP_EBADARG;
    if (maxNx == 0) return DICP_OK;
    BatchDims B{dims, active, fstride};
    const dim3 grid((unsigned)((maxNx + 127) / 128), (unsigned)ntimes, (unsigned)K);
    const float thr2 = radius * radius;
    if (D == 2) batch_coverage_kernel<2><<<grid, 128, 0, (cudaStream_t)stream>>>(B, traj, tstride, thr2, counts, ntimes);
    else batch_coverage_kernel<3><<<grid, 128, 0, (cudaStream_t)stream>>>(B, traj, tstride, thr2, counts, ntimes);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

// ---- set-up helpers on point sets ----------------------------------------------------------------------------------------
int dicp_min2_sqdist(int D, const float* x, int64_t N, float* out, void* stream) {
    if ((D != 2 && D != 3) || N < 0 || N > INT32_MAX) return DICP_EBADARG;
    if (N == 0) return DICP_OK;
    if (!x || !out) return DICP_EBADARG;
    const unsigned blocks = (unsigned)((N + kPsThreads - 1) / kPsThreads);
    if (D == 2) min2_sqdist_kernel<2><<<blocks, kPsThreads, 0, (cudaStream_t)stream>>>(x, (int)N, out);
    else min2_sqdist_kernel<3><<<blocks, kPsThreads, 0, (cudaStream_t)stream>>>(x, (int)N, out);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

size_t dicp_decimate_workspace_bytes(int64_t N) {
    if (N < 1) N = 1;
    const size_t blocks = (size_t)((N + kPsThreads - 1) / kPsThreads);
    return align_up((size_t)N, 256) + 256 + align_up(blocks * 8, 256);          // flags | ctrl | cand
}

int dicp_decimate_steps(int D, const float* x, int64_t N, float radius, int restart, int nsteps, int* kept, void* workspace,
                        size_t workspace_bytes, void* stream) {
    if ((D != 2 && D != 3) || N < 1 || N > INT32_MAX || !(radius >= 0.f) || nsteps < 0 || !x || !kept) return DICP_EBADARG;
    if (!workspace || workspace_bytes < dicp_decimate_workspace_bytes(N)) return DICP_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    DecimState S{};
    S.x = x; S.N = (int)N;
    S.thr2 = (float)((double)radius * (double)radius);       // R**2 in double, compared in fp32 like the reference
    S.flags = (unsigned char*)workspace;
    S.ctrl = (int*)((char*)workspace + align_up((size_t)N, 256));
    S.cand = S.ctrl + 64;
    S.kept = kept;
    if (restart) {
        static const int init[4] = {0, 0, -1, 0};
        cudaMemsetAsync(S.flags, 1, (size_t)N, st);
        cudaMemcpyAsync(S.ctrl, init, sizeof(init), cudaMemcpyHostToDevice, st);
    }
    const unsigned blocks = (unsigned)((N + kPsThreads - 1) / kPsThreads);
    for (int it = 0; it < nsteps; ++it) {
        if (D == 2) decim_step_kernel<2><<<blocks, kPsThreads, 0, st>>>(S);
        else decim_step_kernel<3><<<blocks, kPsThreads, 0, st>>>(S);
    }
    launch_counter() += (unsigned long long)nsteps;
    return last_error(DICP_OK);
}

int dicp_decimate_status(const void* workspace, int64_t N, int* nkept_done, void* stream) {
    if (!workspace || !nkept_done || N < 1) return DICP_EBADARG;
    const int* ctrl = (const int*)((const char*)workspace + align_up((size_t)N, 256));
    cudaError_t e = cudaMemcpyAsync(nkept_done, ctrl, 2 * sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? DICP_OK : (int)e;
}

int dicp_quad_loss(int D, const float* x, const float* y, const float* inv, int64_t n, float* g, float* loss,
                   void* workspace, size_t workspace_bytes, void* stream) {
    if ((D != 2 && D != 3) || n < 1 || !x || !y || !inv || !g || !loss || !workspace) return DICP_EBADARG;
    long long blocks = (n + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (workspace_bytes < (size_t)blocks * 4) return DICP_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    quad_loss_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, inv, n, D, g, (float*)workspace);
    scalar_reduce_kernel<<<1, 256, 0, st>>>((const float*)workspace, (int)blocks, 1, loss, 0);
    launch_counter() += 2;
    return last_error(DICP_OK);
}

int dicp_axpy(int64_t n, float* out, const float* a, float alpha, const float* f1, float beta, const float* f2,
              void* stream) {
    if (n < 0 || (n > 0 && (!out || !a || !f1))) return DICP_EBADARG;
    if (n == 0) return DICP_OK;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)device_info().sms * 16;
    if (blocks > cap) blocks = cap;
    axpy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, out, a, alpha, f1, beta, f2);
    launch_counter() += 1;
    return last_error(DICP_OK);
}

int dicp_pipe_probe(int which, int blocks, int iters, float* out, void* stream) {
    if (!out || blocks < 1 || iters < 1) return DICP_EBADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (which == 0) probe_kernel<0><<<blocks, 256, 0, st>>>(iters, out);
    else if (which == 1) probe_kernel<1><<<blocks, 256, 0, st>>>(iters, out);
    else if (which == 2) probe_kernel<2><<<blocks, 256, 0, st>>>(iters, out);
    else return DICP_EBADARG;
    return last_error(DICP_OK);
}


}  // extern "C"
