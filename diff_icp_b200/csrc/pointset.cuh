// Set-up kernels on point sets: nearest-neighbour scale and greedy decimation.
//
//  * min2_sqdist_kernel: second smallest squared distance from every point to the points of its own set -- the
//    `Kmin(2, dim=1)[:, 1]` reduction of intrinsic_scale (/root/reference/diffICP/tools/point_sets.py:13-26); the smallest
//    is the point itself (0).
//  * decim_step_kernel: ONE pick of the greedy decimation (/root/reference/diffICP/tools/point_sets.py:102-133): among the
//    not-yet-covered points take the one with most not-yet-covered neighbours within R (itself included; smallest index
//    on ties, like argmax over the ascending `notcovered` list), keep it, and cover its neighbours.  The reference builds
//    the dense N x N boolean matrix on the host and loops in Python; here every pick is one launch over (uncovered rows) x
//    (all columns) with the points staged tile by tile in shared memory, and nothing N x N is ever stored.
// Distances are evaluated un-fused in the reference's operation order ((x_i - x_j)**2).sum(-1) <= R**2, so kept / rejected
// INDEX lists are bit-exact.
#pragma once
#include "common.cuh"

namespace dicp {

static constexpr int kPsThreads = 128;
static constexpr int kPsTile = 512;

template <int D>
DICP_D float sqdist_ref(const float* a, const float* b) {
    float d2 = 0.f;
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float dc = __fsub_rn(a[c], b[c]);
        d2 = c == 0 ? __fmul_rn(dc, dc) : __fadd_rn(d2, __fmul_rn(dc, dc));
    }
    return d2;
}

template <int D>
__global__ void __launch_bounds__(kPsThreads) min2_sqdist_kernel(const float* __restrict__ x, int N, float* __restrict__ out) {
    __shared__ float tile[kPsTile * D];
    const int i = blockIdx.x * kPsThreads + threadIdx.x;
    float xi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) xi[c] = i < N ? x[(size_t)i * D + c] : 0.f;
    float m1 = INFINITY, m2 = INFINITY;                       // smallest, second smallest
    for (int j0 = 0; j0 < N; j0 += kPsTile) {
        const int n = N - j0 < kPsTile ? N - j0 : kPsTile;
        __syncthreads();
        for (int t = threadIdx.x; t < n * D; t += kPsThreads) tile[t] = x[(size_t)j0 * D + t];
        __syncthreads();
        for (int j = 0; j < n; ++j) {
            const float d2 = sqdist_ref<D>(xi, &tile[j * D]);
            const float hi = fmaxf(m1, d2);
            m1 = fminf(m1, d2);
            m2 = fminf(m2, hi);
        }
    }
    if (i < N) out[i] = m2;
}

// ctrl: { nkept, done, last_pick, ticket }
struct DecimState {
    const float* x;
    int N;
    float thr2;
    unsigned char* flags;     // (N) 1 = not yet covered
    int* kept;                // (N) kept indices, in pick order
    int* ctrl;                // (4)
    int* cand;                // (gridDim.x, 2): best (count, index) of every row block
};

template <int D>
__global__ void __launch_bounds__(kPsThreads) decim_step_kernel(DecimState S) {
    __shared__ float tile[kPsTile * D];
    __shared__ unsigned char tflag[kPsTile];
    __shared__ int sc[kPsThreads / 32], si[kPsThreads / 32];
    __shared__ int s_last;
    volatile int* ctrl = S.ctrl;
    if (ctrl[1]) return;                                        // all points are covered
    const int N = S.N, tid = threadIdx.x;
    const int pick = ctrl[2];
    float xp[D];
#pragma unroll
    for (int c = 0; c < D; ++c) xp[c] = pick >= 0 ? S.x[(size_t)pick * D + c] : 0.f;
    const int ntiles = (N + kPsTile - 1) / kPsTile;
    // (1) cover the neighbours of the previous pick: the tiles assigned to this CTA get their flags written back
    for (int T = blockIdx.x; T < ntiles; T += gridDim.x) {
        for (int t = tid; t < kPsTile; t += kPsThreads) {
            const int j = T * kPsTile + t;
            if (j < N && pick >= 0 && S.flags[j]) {
                float xj[D];
#pragma unroll
                for (int c = 0; c < D; ++c) xj[c] = S.x[(size_t)j * D + c];
                if (sqdist_ref<D>(xp, xj) <= S.thr2) S.flags[j] = 0;
            }
        }
    }
    // (2) number of uncovered neighbours of every uncovered row of this block.  Other CTAs may not have written their
    // tiles' flags yet: every reader applies the previous pick's test itself (idempotent).
    const int i = blockIdx.x * kPsThreads + tid;
    float xi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) xi[c] = i < N ? S.x[(size_t)i * D + c] : 0.f;
    bool unc = i < N && S.flags[i] != 0;
    if (unc && pick >= 0 && sqdist_ref<D>(xp, xi) <= S.thr2) unc = false;
    int count = -1;
    if (__syncthreads_or(unc)) {
        count = unc ? 0 : -1;
        for (int j0 = 0; j0 < N; j0 += kPsTile) {
            const int n = N - j0 < kPsTile ? N - j0 : kPsTile;
            __syncthreads();
            for (int t = tid; t < n; t += kPsThreads) {
                float xj[D];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    xj[c] = S.x[(size_t)(j0 + t) * D + c];
                    tile[t * D + c] = xj[c];
                }
                unsigned char f = S.flags[j0 + t];
                if (f && pick >= 0 && sqdist_ref<D>(xp, xj) <= S.thr2) f = 0;
                tflag[t] = f;
            }
            __syncthreads();
            if (unc) {
                for (int j = 0; j < n; ++j)
                    if (tflag[j] && sqdist_ref<D>(xi, &tile[j * D]) <= S.thr2) ++count;
            }
        }
    }
    // (3) block arg max (largest count, smallest index), then the last CTA picks the global one
    int bc = count, bi = i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_down_sync(0xffffffffu, bc, o), oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
    if ((tid & 31) == 0) { sc[tid >> 5] = bc; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kPsThreads / 32; ++w)
            if (sc[w] > bc || (sc[w] == bc && si[w] < bi)) { bc = sc[w]; bi = si[w]; }
        S.cand[2 * blockIdx.x] = bc;
        S.cand[2 * blockIdx.x + 1] = bi;
        __threadfence();
        s_last = atomicAdd(&S.ctrl[3], 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    bc = -1; bi = 0x7fffffff;
    for (int b = tid; b < (int)gridDim.x; b += kPsThreads) {
        const int oc = __ldcg(&S.cand[2 * b]), oi = __ldcg(&S.cand[2 * b + 1]);
        if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_down_sync(0xffffffffu, bc, o), oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (oc > bc || (oc == bc && oi < bi)) { bc = oc; bi = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { sc[tid >> 5] = bc; si[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kPsThreads / 32; ++w)
            if (sc[w] > bc || (sc[w] == bc && si[w] < bi)) { bc = sc[w]; bi = si[w]; }
        if (bc <= 0) {
            ctrl[1] = 1;                                        // nothing left to cover
        } else {
            const int n = ctrl[0];
            S.kept[n] = bi;
            ctrl[0] = n + 1;
            ctrl[2] = bi;
        }
        ctrl[3] = 0;
    }
}

}  // namespace dicp
