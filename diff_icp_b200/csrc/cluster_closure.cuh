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
// Scope: eta = 0 (classic / hybrid model), data points present, Euler scheme, M <= kCcMaxM support points,
// Nx <= kCcCluster * kCcThreads * kCcMaxRows data points per frame.  Everything else keeps the stage kernels.
#pragma once
#include "small_step.cuh"
#include <cooperative_groups.h>

namespace dicp {
namespace cg = cooperative_groups;

#ifndef DICP_CC_THREADS
#define DICP_CC_THREADS 128                    // threads per CTA (swept on B200)
#endif
#ifndef DICP_CC_MINB
#define DICP_CC_MINB 4                         // resident CTAs per SM asked of the compiler (register cap 65536 / (T * MINB))
#endif
static constexpr int kCcThreads = DICP_CC_THREADS;
static constexpr int kCcCluster = 8;           // portable cluster size
static constexpr int kCcMaxRows = 2048 / kCcThreads;   // rows per thread at most: up to 8 * 2048 = 16384 data points per frame
static constexpr int kCcMaxM = kCcThreads / 2; // support points (one q row per thread, >= 2 column groups in the reduction)

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
    int rows_cap;             // rows per CTA (multiple of kCcThreads)
    float h, kappa, s, alpha, beta, lam_reg;
};

DICP_HD size_t cc_al4(size_t n) { return (n + 3) & ~(size_t)3; }      // segments start on 16-byte boundaries (float4 reads)
DICP_HD size_t cluster_closure_smem_floats(int M, int D, int nt, int rows_cap) {
    const size_t Mp = (size_t)(M + 1) / 2 * 2, cap = (size_t)rows_cap, MD = (size_t)M * D;
    return cc_al4((size_t)(nt + 1) * 2 * MD)   // support trajectory (q_t | p_t)
           + 3 * cc_al4(2 * MD)                 // cotangents (a | u), F (vq | dp), G (gq | gp)
           + cc_al4(MD)                         // vq(0)
           + cc_al4(Mp * 4 * D)                 // packed support columns (largest record: q', p, a, u)
           + cc_al4(cap * 2 * D)                // packed data-point columns (x', wx)
           + 2 * cc_al4(cap * D)                // x, lambda_x of the CTA's rows
           + cc_al4(2 * (size_t)(2 * D + 1) * M)    // published partial sums, double buffered
           + cc_al4((size_t)kCcThreads * (2 * D + 1))   // group reduction scratch
           + 64;                                // origin, gc, block scalars
}

template <int D, bool WLD>
__global__ void __launch_bounds__(kCcThreads, DICP_CC_MINB) cluster_closure_kernel(ClusterClosure C) {
    using OpXQ = RhsXQ<D, WLD, false, 1>;        // forward, rows x
    using OpQQf = RhsQQ<D, false, false, 1>;     // forward, rows q (x present: the divergence cost comes from the x rows)
    using OpX = AdjXQx<D, WLD, 1>;               // adjoint, rows x, cols (q,p)
    using OpQx = AdjXQq<D, WLD, 1>;              // adjoint, rows q, cols (x, wx)
    using OpQQa = AdjQQ<D, false, 1>;            // adjoint, rows q, cols (q,p,a,u)
    constexpr int NAX = OpQx::NACC, T = kCcThreads;
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
        // x rows of this CTA, four at a time
        int j = tid;
        for (; j + 3 * T < n_cta; j += 4 * T) {
            typename OpXQ::Row row[4];
            F2 acc[4][OpXQ::NACC];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int c = 0; c < D; ++c) row[r].x[c] = (sx[(size_t)(j + r * T) * D + c] - misc[c]) * C.kappa;
#pragma unroll
                for (int a = 0; a < OpXQ::NACC; ++a) acc[r][a] = f2(0.f, 0.f);
            }
            sweep_cols_multi<OpXQ, 4>(P, row, cols, M, acc);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
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
        for (; j < n_cta; j += T) {
            typename OpXQ::Row row;
            F2 acc[OpXQ::NACC];
#pragma unroll
            for (int c = 0; c < D; ++c) row.x[c] = (sx[(size_t)j * D + c] - misc[c]) * C.kappa;
#pragma unroll
            for (int a = 0; a < OpXQ::NACC; ++a) acc[a] = f2(0.f, 0.f);
            sweep_cols<OpXQ>(P, row, cols, M, acc);
            const size_t o = (size_t)j * D;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float xn = fmaf(h, f2_sum(acc[OpXQ::A_V + c]), sx[o + c]);
                sx[o + c] = xn;
                C.traj[(long long)(t + 1) * C.tstride + xoff + (long long)o + c] = xn;
            }
            if (WLD) dcsum = fmaf(C.alpha, f2_sum(acc[OpXQ::A_DS]), dcsum);
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
        int j = tid;
        for (; j + 3 * T < n_cta; j += 4 * T) {
            typename OpX::Row row[4];
            F2 acc[4][OpX::NACC];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    row[q].x[c] = (sx[(size_t)(j + q * T) * D + c] - misc[c]) * C.kappa;
                    row[q].w[c] = slx[(size_t)(j + q * T) * D + c];
                }
                row[q].gc = WLD ? 1.f : 0.f;
#pragma unroll
                for (int a = 0; a < OpX::NACC; ++a) acc[q][a] = f2(0.f, 0.f);
            }
            sweep_cols_multi<OpX, 4>(P, row, cols, M, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const size_t o = (size_t)(j + q * T) * D + c;
                    slx[o] = fmaf(h, f2_sum(acc[q][c]), slx[o]);
                }
            }
        }
        for (; j < n_cta; j += T) {
            typename OpX::Row row;
            F2 acc[OpX::NACC];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                row.x[c] = (sx[(size_t)j * D + c] - misc[c]) * C.kappa;
                row.w[c] = slx[(size_t)j * D + c];
            }
            row.gc = WLD ? 1.f : 0.f;
#pragma unroll
            for (int a = 0; a < OpX::NACC; ++a) acc[a] = f2(0.f, 0.f);
            sweep_cols<OpX>(P, row, cols, M, acc);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const size_t o = (size_t)j * D + c;
                slx[o] = fmaf(h, f2_sum(acc[c]), slx[o]);
            }
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

}  // namespace dicp
