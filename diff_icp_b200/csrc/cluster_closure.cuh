// The WHOLE registration closure of a frame -- Euler shoot, lambda*H(q0,p0) + cost(1), quadratic data loss, adjoint sweep,
// gradient with respect to p0 -- in ONE kernel launch: one thread-block CLUSTER per frame (blockIdx.y = frame).
//
// Why: with a grid / decimated support (tens of support points, 10^4 data points per frame) DiffPSR.Reg_opt
// (/root/reference/diffICP/core/PSR.py:521-569 -> core/LDDMM.py:338-398 -> tools/optim.py:32-50) evaluates ~25 rounds of
// closures per outer iteration, and a round through the stage kernels of small_step.cuh is 1 + nt + 1 + nt + 1 DEPENDENT
// launches of a few microseconds of arithmetic each: the round is bound by launch / tail latency (0.58 ms for 64 frames
// where the arithmetic is ~0.2 ms).  Here the stages of a frame follow each other inside one kernel:
//   * the cluster's CTAs split the frame's data points (contiguous ranges of `rows_cap` rows); a CTA keeps its rows' positions
//     and cotangents in shared memory for the whole closure and writes x(t) to the trajectory buffer for its own later use;
//   * the support state (q_t, p_t) is tiny, so EVERY CTA integrates it redundantly (M x M pairs per stage) and keeps the
//     whole support trajectory in shared memory: the forward sweep needs no communication at all;
//   * an adjoint stage needs, for every support point, sums over ALL data points: each CTA reduces its own rows
//     (thread groups per support point, like small_adj_step_kernel), publishes M x (2D+1) partial sums in its shared memory,
//     and after ONE cluster barrier per stage (arrive ... x-row pass ... wait) every CTA adds the partials of all ranks in
//     rank order through distributed shared memory -- identical bits everywhere, no atomics, deterministic;
//   * cost(1) = h sum_t dcost_t and the data loss are summed per thread over all stages and reduced once at the end.
// The per-pair arithmetic is the SAME Op code as everywhere else (ops_rhs.cuh: RhsXQ, RhsQQ, AdjXQx, AdjXQq, AdjQQ), fed from
// shared-memory copies of the state instead of global memory.
//
// Scope: eta = 0 (classic / hybrid model), data points present, Euler scheme, M <= T / 2 support points (T = 64 or 128 threads
// per CTA), Nx <= cluster size * kCcMaxRowsPerCta data points per frame.  Everything else keeps the stage kernels.
#pragma once
#include "small_step.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>

namespace dicp {
namespace cg = cooperative_groups;

// Launch shapes: T threads per CTA (template), `cluster` CTAs per frame (run time; 16 is the non-portable size, opt-in).  More,
// smaller CTAs per frame even out the load over the SMs (64 frames: 8 x 128 threads = 512 CTAs on 148 SMs leave every SM
// with 3 or 4 of them, 16 x 64 threads = 1024 CTAs with 6 or 7) and use more SMs when a rank holds few frames.
static constexpr int kCcRegCap = 128;          // registers per thread asked of the compiler (65536 / (T * resident CTAs))
static constexpr int kCcMaxRowsPerCta = 2048;  // data points per CTA at most
DICP_HD int cc_max_support(int T) { return T / 2; }   // one q row per thread, >= 2 column groups in the reduction

struct ClusterClosure {
    const int* dims;          // (K,2): M_k, Nx_k
    const int* active;        // (K) nullable
    float* traj;              // (nt+1, K, fstride): state layout [q | p | x | cost]; traj[0] holds q0 and x0; x(t) is written here
    long long fstride, tstride;
    const float* X;           // (K, xstride): trial momenta p0
    long long xstride;
    const float* y;           // (K, ystride, D) targets
    const float* inv;         // (K, ystride)   weights 1 / (2 sigma_s^2)
    long long ystride;
    float* out;               // (K, ostride): [0, A, 0, 0, cost(1), data loss, 0, 0 | d loss / d p0 (M D)]
    long long ostride;
    int ns;
    int nt;
    int rows_cap;             // rows per CTA (multiple of the CTA size)
    float h, kappa, s, alpha, beta, lam_reg;
};

DICP_HD size_t cc_al4(size_t n) { return (n + 3) & ~(size_t)3; }      // segments start on 16-byte boundaries (float4 reads)
DICP_HD size_t cluster_closure_smem_floats(int M, int D, int nt, int rows_cap, int T) {
    const size_t Mp = (size_t)(M + 1) / 2 * 2, cap = (size_t)rows_cap, MD = (size_t)M * D;
    return cc_al4((size_t)(nt + 1) * 2 * MD)   // support trajectory (q_t | p_t)
           + 3 * cc_al4(2 * MD)                 // cotangents (a | u), F (vq | dp), G (gq | gp)
           + cc_al4(MD)                         // vq(0)
           + cc_al4(Mp * 4 * D)                 // packed support columns (largest record: q', p, a, u)
           + cc_al4(cap * 2 * D)                // packed data-point columns (x', wx)
           + 2 * cc_al4(cap * D)                // x, lambda_x of the CTA's rows
           + cc_al4(2 * (size_t)(2 * D + 1) * M)    // published partial sums, double buffered
           + cc_al4((size_t)T * (2 * D + 1))            // group reduction scratch
           + 64;                                // origin, gc, block scalars
}

// R rows of one thread (j, j + T, ...) of the forward (x,q) pass swept together: x <- x + h v(x), trajectory write, dcost sum
template <class OpXQ, int D, bool WLD, int T, int R>
DICP_D void cc_fwd_rows(const RhsParams& P, const ClusterClosure& C, const float* cols, float* sx, const float* misc, int M,
                        int j, int t, long long xoff, float h, float& dcsum) {
    typename OpXQ::Row row[R];
    F2 acc[R][OpXQ::NACC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int c = 0; c < D; ++c) row[r].x[c] = (sx[(size_t)(j + r * T) * D + c] - misc[c]) * C.kappa;
#pragma unroll
        for (int a = 0; a < OpXQ::NACC; ++a) acc[r][a] = f2(0.f, 0.f);
    }
    sweep_cols_multi<OpXQ, R>(P, row, cols, M, acc);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t o = (size_t)(j + r * T) * D;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float xn = fmaf(h, f2_sum(acc[r][OpXQ::A_V + c]), sx[o + c]);
            sx[o + c] = xn;
            C.traj[(long long)(t + 1) * C.tstride + xoff + (long long)o + c] = xn;
        }
        if (WLD) dcsum = fmaf(C.alpha, f2_sum(acc[r][OpXQ::A_DS]), dcsum);
    }
}

// R rows of one thread of the adjoint (x,q) pass w.r.t. x: lambda_x <- lambda_x + h gx
template <class OpX, int D, bool WLD, int T, int R>
DICP_D void cc_adj_rows(const RhsParams& P, const ClusterClosure& C, const float* cols, const float* sx, float* slx,
                        const float* misc, int M, int j, float h) {
    typename OpX::Row row[R];
    F2 acc[R][OpX::NACC];
#pragma unroll
    for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            row[q].x[c] = (sx[(size_t)(j + q * T) * D + c] - misc[c]) * C.kappa;
            row[q].w[c] = slx[(size_t)(j + q * T) * D + c];
        }
        row[q].gc = WLD ? 1.f : 0.f;
#pragma unroll
        for (int a = 0; a < OpX::NACC; ++a) acc[q][a] = f2(0.f, 0.f);
    }
    sweep_cols_multi<OpX, R>(P, row, cols, M, acc);
#pragma unroll
    for (int q = 0; q < R; ++q) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const size_t o = (size_t)(j + q * T) * D + c;
            slx[o] = fmaf(h, f2_sum(acc[q][c]), slx[o]);
        }
    }
}

template <int D, bool WLD, int T>
__global__ void __launch_bounds__(T, 65536 / (T * kCcRegCap)) cluster_closure_kernel(ClusterClosure C) {
    using OpXQ = RhsXQ<D, WLD, false, 1>;        // forward, rows x
    using OpQQf = RhsQQ<D, false, false, 1>;     // forward, rows q (x present: the divergence cost comes from the x rows)
    using OpX = AdjXQx<D, WLD, 1>;               // adjoint, rows x, cols (q,p)
    using OpQx = AdjXQq<D, WLD, 1>;              // adjoint, rows q, cols (x, wx)
    using OpQQa = AdjQQ<D, false, 1>;            // adjoint, rows q, cols (q,p,a,u)
    constexpr int NAX = OpQx::NACC;
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[32];

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), NC = (int)cluster.num_blocks();
    const int k = blockIdx.y, tid = threadIdx.x;
    if (C.active != nullptr && C.active[k] == 0) return;            // uniform over the cluster
    const int M = C.dims[2 * k], Nx = C.dims[2 * k + 1];
    const int MD = M * D, Mp = (M + 1) / 2 * 2, cap = C.rows_cap, nt = C.nt;
    const float h = C.h;

    // ---- shared memory carve-up (must match cluster_closure_smem_floats) ----------------------------------------------
    float* qtr = sm;                                          // (nt+1) x [q (MD) | p (MD)]
    float* lam = qtr + cc_al4((size_t)(nt + 1) * 2 * MD);     // a (MD) | u (MD)
    float* Fq = lam + cc_al4(2 * (size_t)MD);                 // vq | dp
    float* Gq = Fq + cc_al4(2 * (size_t)MD);                  // gq | gp
    float* vq0 = Gq + cc_al4(2 * (size_t)MD);                 // vq at t = 0
    float* cols = vq0 + cc_al4((size_t)MD);                   // Mp * 4D
    float* xcols = cols + cc_al4((size_t)Mp * 4 * D);         // cap * 2D
    float* sx = xcols + cc_al4((size_t)cap * 2 * D);          // cap * D
    float* slx = sx + cc_al4((size_t)cap * D);                // cap * D
    float* part = slx + cc_al4((size_t)cap * D);              // 2 x NAX x M
    float* xch = part + cc_al4(2 * (size_t)(2 * D + 1) * M);  // T x NAX
    float* misc = xch + cc_al4((size_t)T * (2 * D + 1));      // [0..D): origin, [4]: gc, [8..16): scalars

    const long long fo = (long long)k * C.fstride;
    const int r0 = rank * cap;                                        // first global row of this CTA
    const int n_cta = Nx - r0 < 0 ? 0 : (Nx - r0 < cap ? Nx - r0 : cap);
    const long long xoff = fo + 2LL * MD + (long long)r0 * D;          // x rows of this CTA inside a state

    // ---- initial state --------------------------------------------------------------------------------------------------
    for (int i = tid; i < MD; i += T) {
        qtr[i] = C.traj[fo + i];
        qtr[MD + i] = C.X[(long long)k * C.xstride + i];
        lam[i] = 0.f;
        lam[MD + i] = 0.f;
    }
    for (int i = tid; i < n_cta * D; i += T) sx[i] = C.traj[xoff + i];
    if (tid < D) misc[tid] = C.traj[fo + tid];        // origin = first support point at t = 0, for the whole closure
    if (tid == 0) misc[4] = 1.f;                      // cotangent of the cost channel
    __syncthreads();

    RhsParams P{};
    P.origin = misc;
    P.kappa = C.kappa; P.s = C.s; P.alpha = C.alpha; P.beta = C.beta; P.eta = 0.f;
    P.gc = misc + 4;
    P.vq = Fq; P.dp = Fq + MD;
    P.gq = Gq; P.gp = Gq + MD;
    P.a = lam; P.u = lam + MD;
    P.x = sx; P.wx = slx;

    float dcsum = 0.f, Asum = 0.f;
    // ---- forward sweep: no communication ----------------------------------------------------------------------------------
    for (int t = 0; t < nt; ++t) {
        P.q = qtr + (size_t)t * 2 * MD;
        P.p = P.q + MD;
        stage_cols<OpXQ>(P, 0, M, M, cols);
        __syncthreads();
        // x rows of this CTA: 4, then 2, then 1 rows of a thread swept together (independent chains hide the pair latency)
        {
            int j = tid;
            for (; j + 3 * T < n_cta; j += 4 * T) cc_fwd_rows<OpXQ, D, WLD, T, 4>(P, C, cols, sx, misc, M, j, t, xoff, h, dcsum);
            if (j + T < n_cta) { cc_fwd_rows<OpXQ, D, WLD, T, 2>(P, C, cols, sx, misc, M, j, t, xoff, h, dcsum); j += 2 * T; }
            if (j < n_cta) cc_fwd_rows<OpXQ, D, WLD, T, 1>(P, C, cols, sx, misc, M, j, t, xoff, h, dcsum);
        }
        // q rows (every CTA, redundantly): vq, dp, A_i; then the Euler update into the next slot of the support trajectory
        if (tid < M) {
            typename OpQQf::Row row;
            OpQQf::load_row(P, tid, row);
            F2 acc[OpQQf::NACC];
#pragma unroll
            for (int a = 0; a < OpQQf::NACC; ++a) acc[a] = f2(0.f, 0.f);
            sweep_cols<OpQQf>(P, row, cols, M, acc);
            float a[OpQQf::NACC], rs[3];
#pragma unroll
            for (int c = 0; c < OpQQf::NACC; ++c) a[c] = f2_sum(acc[c]);
            OpQQf::finish(P, tid, row, a, rs);
            if (t == 0) Asum = rs[0];
            float* nq = qtr + (size_t)(t + 1) * 2 * MD;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const int o = tid * D + c;
                nq[o] = fmaf(h, Fq[o], P.q[o]);
                nq[MD + o] = fmaf(h, Fq[MD + o], P.p[o]);
                if (t == 0) vq0[o] = Fq[o];
            }
        }
        __syncthreads();
    }

    // ---- data loss on the arrival points and its cotangent (DiffPSR.QuadLossFunctor, core/PSR.py:498-516) -------------------
    float dlsum = 0.f;
    for (int j = tid; j < n_cta; j += T) {
        const long long g = (long long)k * C.ystride + r0 + j;
        const float w = C.inv[g];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float r = sx[(size_t)j * D + c] - C.y[g * D + c];
            slx[(size_t)j * D + c] = 2.f * w * r;
            dlsum = fmaf(w * r, r, dlsum);
        }
    }
    __syncthreads();                           // the adjoint sweep reloads sx (other threads' rows) right away

    // ---- adjoint sweep: ONE cluster barrier per stage ----------------------------------------------------------------------
    const int G = T / M;                       // column groups of the q-side reduction (M <= T / 2)
    const int g = tid / M, r = tid - g * M;
    const bool work = g < G;
    int buf = 0;
    for (int t = nt - 1; t >= 0; --t) {
        P.q = qtr + (size_t)t * 2 * MD;
        P.p = P.q + MD;
        // this CTA's x(t) (its own earlier writes; t = 0: the frame's data points)
        for (int i = tid; i < n_cta * D; i += T) sx[i] = C.traj[(long long)t * C.tstride + xoff + i];
        stage_cols<OpX>(P, 0, M, M, cols);
        __syncthreads();
        stage_cols<OpQx>(P, 0, n_cta, n_cta, xcols);
        __syncthreads();
        // (1) support-point sums over this CTA's rows: thread (g, r) sweeps group g's share for support point r
        float ax[NAX];
        {
            typename OpQx::Row row;
            F2 acc[NAX];
#pragma unroll
            for (int a = 0; a < NAX; ++a) acc[a] = f2(0.f, 0.f);
            if (work) {
                OpQx::load_row(P, r, row);
                sweep_share<OpQx>(P, row, xcols, n_cta, g, G, acc);
            }
#pragma unroll
            for (int a = 0; a < NAX; ++a) ax[a] = f2_sum(acc[a]);
        }
        if (work && g > 0) {
#pragma unroll
            for (int a = 0; a < NAX; ++a) xch[(size_t)(g * NAX + a) * M + r] = ax[a];
        }
        __syncthreads();
        float* mypart = part + (size_t)buf * NAX * M;
        if (work && g == 0) {
            for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
                for (int a = 0; a < NAX; ++a) ax[a] += xch[(size_t)(g2 * NAX + a) * M + r];
            }
#pragma unroll
            for (int a = 0; a < NAX; ++a) mypart[(size_t)a * M + r] = ax[a];
        }
        cluster.barrier_arrive();                                  // release: my partials are published
        // (2) x rows: gx from the OLD cotangents, lambda_x <- lambda_x + h gx (overlaps the other ranks' arrival)
        {
            int j = tid;
            for (; j + 3 * T < n_cta; j += 4 * T) cc_adj_rows<OpX, D, WLD, T, 4>(P, C, cols, sx, slx, misc, M, j, h);
            if (j + T < n_cta) { cc_adj_rows<OpX, D, WLD, T, 2>(P, C, cols, sx, slx, misc, M, j, h); j += 2 * T; }
            if (j < n_cta) cc_adj_rows<OpX, D, WLD, T, 1>(P, C, cols, sx, slx, misc, M, j, h);
        }
        __syncthreads();                                           // every thread is done with `cols` (2D records)
        // (3) (q,q) interaction columns (q', p, a, u)
        stage_cols<OpQQa>(P, 0, M, M, cols);
        cluster.barrier_wait();                                    // acquire: every rank's partials are visible
        __syncthreads();
        if (tid < M) {
            float tot[NAX];
#pragma unroll
            for (int a = 0; a < NAX; ++a) tot[a] = 0.f;
            for (int rk = 0; rk < NC; ++rk) {                      // rank order: identical sums in every CTA
                const float* pp = cluster.map_shared_rank(part, rk) + (size_t)buf * NAX * M;
#pragma unroll
                for (int a = 0; a < NAX; ++a) tot[a] += pp[(size_t)a * M + tid];
            }
            typename OpQQa::Row row;
            OpQQa::load_row(P, tid, row);
            F2 acc[OpQQa::NACC];
#pragma unroll
            for (int a = 0; a < OpQQa::NACC; ++a) acc[a] = f2(0.f, 0.f);
            sweep_cols<OpQQa>(P, row, cols, M, acc);
            float aq[OpQQa::NACC];
#pragma unroll
            for (int a = 0; a < OpQQa::NACC; ++a) aq[a] = f2_sum(acc[a]);
            P.accumulate = 0;
            OpQQa::finish(P, tid, row, aq, nullptr);
            typename OpQx::Row rowx;
            OpQx::load_row(P, tid, rowx);
            P.accumulate = 1;
            OpQx::finish(P, tid, rowx, tot, nullptr);
        }
        __syncthreads();                                           // all rows have read lam (their own row) and `cols`
        if (tid < M) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const int o = tid * D + c;
                lam[o] = fmaf(h, Gq[o], lam[o]);
                lam[MD + o] = fmaf(h, Gq[MD + o], lam[MD + o]);
            }
        }
        __syncthreads();
        buf ^= 1;
    }

    // ---- outputs ----------------------------------------------------------------------------------------------------------------
    {
        const float a = block_sum(tid < M ? Asum : 0.f, red);
        const float dc = block_sum(dcsum, red);
        const float dl = block_sum(dlsum, red);
        if (tid == 0) { misc[8] = a; misc[9] = h * dc; misc[10] = dl; }
    }
    cluster.sync();
    if (rank == 0) {
        float* ok = C.out + (long long)k * C.ostride;
        for (int i = tid; i < MD; i += T) ok[C.ns + i] = fmaf(C.lam_reg, vq0[i], lam[MD + i]);
        if (tid == 0) {
            float cost = 0.f, dl = 0.f;
            for (int rk = 0; rk < NC; ++rk) {
                const float* m2 = cluster.map_shared_rank(misc, rk);
                cost += m2[9];
                dl += m2[10];
            }
            ok[0] = 0.f; ok[1] = misc[8]; ok[2] = 0.f; ok[3] = 0.f; ok[4] = cost; ok[5] = dl;
        }
    }
    cluster.sync();                                                // nobody leaves while rank 0 still reads its shared memory
}

// ---- host side: launch shape and launch ------------------------------------------------------------------------------------------
struct CcShape { int T = 0, cluster = 0, cap = 0; size_t smem = 0; };

// Launch shape for K frames of at most maxM support / maxNx data points: 128 threads per CTA when the support allows 64
// (MEASURED on B200, 64 frames x 10k points, 25 support points, 2-D: 0.43 ms with 8 x 128 threads per frame, 0.47 with 8 x 64,
// 0.49 with 16 x 128 (two waves), 0.55 with 16 x 64 -- the per-CTA fixed work of a stage, redundant support integration, staging,
// barriers, outweighs the better SM balance of many small CTAs); 16 CTAs per frame (the non-portable cluster size) when 8 per
// frame would leave SMs without a CTA (8 frames: 0.172 vs 0.184 ms) or when a frame has more than 8 x 2048 data points.
// DICP_CC_SHAPE="T,cluster" forces a shape (experiments, tests).
inline CcShape cc_pick_shape(int D, int64_t maxM, int64_t maxNx, int nt, int K) {
    int forcedT = 0, forcedC = 0;
    if (const char* e = getenv("DICP_CC_SHAPE")) {
        forcedT = atoi(e);
        const char* c = strchr(e, ',');
        forcedC = c ? atoi(c + 1) : 0;
    }
    const int sms = device_info().sms;
    auto shape = [&](int T, int cl) {
        CcShape sh;
        if ((T != 64 && T != 128) || cl < 1 || cl > 16 || maxM > cc_max_support(T)) return sh;
        const long long per = (long long)cl * T;
        const long long cap = (maxNx + per - 1) / per * T;
        if (cap > kCcMaxRowsPerCta) return sh;
        const size_t smem = cluster_closure_smem_floats((int)maxM, D, nt, (int)cap, T) * 4;
        if (smem > 200 * 1024) return sh;
        sh.T = T; sh.cluster = cl; sh.cap = (int)cap; sh.smem = smem;
        return sh;
    };
    if (forcedT) return shape(forcedT, forcedC);
    const int T = maxM <= cc_max_support(64) && maxNx <= 8 * 64 ? 64 : 128;      // tiny frames: one row per thread at most
    CcShape sh = shape(T, (long long)K * 8 < sms ? 16 : 8);
    if (sh.T == 0) sh = shape(T, 16);
    if (sh.T == 0) sh = shape(128, 16);
    return sh;
}

template <int D, bool WLD, int T>
int cc_launch(const ClusterClosure& C, const CcShape& sh, int K, cudaStream_t st) {
    auto kern = cluster_closure_kernel<D, WLD, T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem);
    if (e != cudaSuccess) return (int)e;
    if (sh.cluster > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)sh.cluster, (unsigned)K, 1);
    cfg.blockDim = dim3(T, 1, 1);
    cfg.dynamicSmemBytes = sh.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)sh.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, C);
    return e == cudaSuccess ? DICP_OK : (int)e;
}

}  // namespace dicp
