// Batched (multi-frame) pieces of the registration closure for the SMALL-SUPPORT regime.
//
// DiffPSR.Reg_opt (/root/reference/diffICP/core/PSR.py:521-569) optimises the K frames independently; with a grid /
// decimated support each frame's closure is a few microseconds of arithmetic, so all frames are evaluated by the SAME
// launches (blockIdx.y = frame): the stage kernels of small_step.cuh in their batched form, plus the four small kernels
// below.  Frame k owns the floats [k * fstride, (k+1) * fstride) of every state-like buffer; its state layout is the usual
// flat [ q (M_k,D) | p (M_k,D) | x (Nx_k,D) | cost ], S_k = 2 M_k D + Nx_k D + 1.  dims = (K,2) int32 {M_k, Nx_k};
// active (K) int32: frames with active[k] == 0 are skipped by every kernel (their buffers keep their old contents).
#pragma once
#include "small_step.cuh"

namespace dicp {

struct BatchDims {
    const int* dims;
    const int* active;      // nullable
    long long fstride;
};

// p part of the initial state <- X[k, :M_k D]; cost <- 0      (LDDMMModel.Shoot's (q0, p0, cost0 = 0), core/LDDMM.py:293-297)
__global__ void batch_set_p_kernel(BatchDims B, int D, const float* __restrict__ X, long long xstride,
                                   float* __restrict__ state0) {
    const int k = blockIdx.y;
    if (B.active != nullptr && B.active[k] == 0) return;
    const int M = B.dims[2 * k], Nx = B.dims[2 * k + 1];
    const int MD = M * D;
    float* st = state0 + (long long)k * B.fstride;
    const float* x = X + (long long)k * xstride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < MD; i += gridDim.x * blockDim.x) st[MD + i] = x[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) st[2 * (long long)MD + (long long)Nx * D] = 0.f;
}

// Quadratic data loss of every frame (DiffPSR.QuadLossFunctor, core/PSR.py:498-516) on the arrival state:
//   loss_k = sum_n inv[k,n] |z_n - y[k,n]|^2,   cotangent of the data part of the arrival state  g_n = 2 inv (z_n - y_n)
// where z = x(1) (Nx_k > 0) or q(1) (dense support).  Block partials are summed by the frame's last CTA in block order
// (deterministic).  counters: one uint32 per frame, zero before the first launch, reset by the kernel.
__global__ void __launch_bounds__(256) batch_quad_loss_kernel(BatchDims B, int D, const float* __restrict__ state_end,
                                                              const float* __restrict__ Y, const float* __restrict__ INV,
                                                              long long ystride, float* __restrict__ g_end,
                                                              float* __restrict__ loss, long long lstride,
                                                              float* __restrict__ partials, int pstride,
                                                              unsigned* __restrict__ counters) {
    __shared__ float red[32];
    const int k = blockIdx.y;
    if (B.active != nullptr && B.active[k] == 0) return;
    const int M = B.dims[2 * k], Nx = B.dims[2 * k + 1];
    const int n = Nx > 0 ? Nx : M;
    const int nblk = (n + 255) / 256 < (int)gridDim.x ? (n + 255) / 256 : (int)gridDim.x;
    if ((int)blockIdx.x >= nblk) return;
    const long long off = (long long)k * B.fstride + (Nx > 0 ? 2LL * M * D : 0LL);
    const float* z = state_end + off;
    float* g = g_end + off;
    const float* y = Y + (long long)k * ystride * D;
    const float* inv = INV + (long long)k * ystride;
    float acc = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += nblk * 256) {
        const float w = inv[i];
        for (int c = 0; c < D; ++c) {
            const float r = z[(long long)i * D + c] - y[(long long)i * D + c];
            g[(long long)i * D + c] = 2.f * w * r;
            acc = fmaf(w * r, r, acc);
        }
    }
    const float v = block_sum(acc, red);
    float* part = partials + (long long)k * pstride;
    if (threadIdx.x == 0) part[blockIdx.x] = v;
    if (last_cta(&counters[k], (unsigned)nblk)) {
        float t = 0.f;
        for (int b = threadIdx.x; b < nblk; b += 256) t += __ldcg(&part[b]);
        t = block_sum(t, red);
        if (threadIdx.x == 0) {
            loss[(long long)k * lstride] = t;
            counters[k] = 0u;
        }
    }
}

// Closure output of frame k:  out[k] = [ dcost(0), A, B, C, cost(1), (data loss: written by the kernel above), -, - |
//                                        d loss / d p0 = lam_p + lambda * vq(0)  (M_k D floats) ]
// (Hamilton's equations: d/dp0 [lambda H(q0,p0)] = lambda vq(0), core/LDDMM.py:156-158).  lam == nullptr: scalars only.
__global__ void batch_closure_out_kernel(BatchDims B, int D, float lam_reg, const float* __restrict__ lam,
                                         const float* __restrict__ F0, const float* __restrict__ state_end,
                                         float* __restrict__ out, long long ostride, int ns) {
    const int k = blockIdx.y;
    if (B.active != nullptr && B.active[k] == 0) return;
    const int M = B.dims[2 * k], Nx = B.dims[2 * k + 1];
    const int MD = M * D;
    const long long S = 2LL * MD + (long long)Nx * D + 1;
    const long long o = (long long)k * B.fstride;
    float* ok = out + (long long)k * ostride;
    if (lam != nullptr)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < MD; i += gridDim.x * blockDim.x)
            ok[ns + i] = fmaf(lam_reg, F0[o + i], lam[o + MD + i]);
    if (blockIdx.x == 0 && threadIdx.x < 4) ok[threadIdx.x] = F0[o + S - 1 + threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 4) ok[4] = state_end[o + S - 1];
}

// Coverage of the (moving) data points by the (moving) support points at every stored time point
// (GaussKernel.check_coverage, tools/kernel.py:324-329, called from core/PSR.py:559-566):
//   counts[k, t] = #{ i : min_j |x_i(t) - q_j(t)|^2 > thr2 },  thr2 = (R sigma)^2.
// Distances are evaluated un-fused, in the reference's operation order.  Integer atomics: deterministic.
template <int D>
__global__ void __launch_bounds__(128) batch_coverage_kernel(BatchDims B, const float* __restrict__ traj, long long tstride,
                                                             float thr2, int* __restrict__ counts, int nt1) {
    __shared__ float sq[kSmallMaxQ * D];
    __shared__ int cnt;
    const int k = blockIdx.z, t = blockIdx.y;
    if (B.active != nullptr && B.active[k] == 0) return;
    const int M = B.dims[2 * k], Nx = B.dims[2 * k + 1];
    if ((int)blockIdx.x * 128 >= Nx) return;
    const float* st = traj + (long long)t * tstride + (long long)k * B.fstride;
    for (int i = threadIdx.x; i < M * D; i += 128) sq[i] = st[i];
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int i = blockIdx.x * 128 + threadIdx.x;
    bool uncovered = false;
    if (i < Nx) {
        float x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = st[2LL * M * D + (long long)i * D + c];
        float best = INFINITY;
        for (int j = 0; j < M; ++j) {
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float dc = __fsub_rn(x[c], sq[j * D + c]);
                d2 = __fadd_rn(d2, __fmul_rn(dc, dc));
            }
            best = fminf(best, d2);
        }
        uncovered = best > thr2;
    }
    const unsigned m = __ballot_sync(0xffffffffu, uncovered);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && cnt) atomicAdd(&counts[k * nt1 + t], cnt);
}

}  // namespace dicp
