// Fused integrator-step kernels for the SMALL-SUPPORT regime (M <= kSmallMaxQ support points, any number of data points).
//
// With a grid / decimated support (DiffPSR.set_support_scheme, /root/reference/diffICP/core/PSR.py:430-493) the support
// set has tens to hundreds of points while every frame carries 10^4..10^5 data points.  One right-hand-side evaluation is
// then ~10^5..10^6 pairs -- about a microsecond of arithmetic -- and the general engine's launch sequence (pack, pair,
// finish, scalar reduction, axpy: ~20 launches per time step forward + adjoint) is pure latency.  Here ONE launch does a
// whole forward stage (right-hand side for the q rows and the x rows, dcost / Hamiltonian scalars, and the Euler / Ralston
// state update) and ONE launch does a whole adjoint stage (VJP for the x rows and the q rows, merge of the column-split
// partials, and the cotangent update).  The per-pair arithmetic is the SAME Op code as the general engine
// (ops_rhs.cuh, packed-fp32 form); only the scaffolding differs:
//   * every CTA stages the (tiny) support set straight from the state vector into shared memory (no pack kernel);
//   * cross-CTA reductions use the "last CTA finishes" pattern (threadfence + ticket counter, partials summed in a fixed
//     order -> still bit-reproducible, no floating-point atomics); counters live in the first bytes of the workspace, must be
//     zero before the first launch and are reset by the kernels themselves.
//
// State / cotangent layout: flat [ q (M,D) | p (M,D) | x (Nx,D) | cost ];  F layout: [ vq | dp | vx | dcost, A, B, C ].
#pragma once
#include "ops_rhs.cuh"

namespace dicp {

static constexpr int kSmallMaxQ = 512;        // support points staged entirely in shared memory
static constexpr int kSmallThreads = 128;     // rows per CTA (one row per thread)
static constexpr int kSmallChunk = 512;       // data-point columns per q-row CTA in the adjoint
static constexpr int kSmallCounters = 64;     // uint32 counters at the head of the workspace

struct SmallStep {
    const float* s_eval;     // state where the right-hand side / Jacobian is evaluated
    const float* lam;        // adjoint only: cotangent [a | u | wx | gc]
    const float* base;       // update: out = base + c_this * This + c_other * other (+ add)
    const float* other;      // nullable
    const float* add;        // nullable (adjoint: d loss / d state at this time point)
    float* out;              // nullable
    float* This;             // F (S+3 floats) for the forward stage, G (S floats) for the adjoint stage; never null
    float c_this, c_other;
    int M, Nx;
    float kappa, s, alpha, beta, eta;
    float* ws;               // workspace after the counters: block scalars / partials
    unsigned* counters;
};

// ---- shared helpers ------------------------------------------------------------------------------------------------
template <int D>
DICP_D RhsParams small_params(const SmallStep& S) {
    RhsParams P{};
    const size_t MD = (size_t)S.M * D;
    P.q = S.s_eval; P.p = S.s_eval + MD; P.x = S.Nx > 0 ? S.s_eval + 2 * MD : nullptr;
    P.origin = S.s_eval;
    P.kappa = S.kappa; P.s = S.s; P.alpha = S.alpha; P.beta = S.beta; P.eta = S.eta;
    return P;
}

// Pack columns [j0, j0+n) of Op's column set into the pair-interleaved layout at `smem` (pair index relative to j0).
template <class Op>
DICP_D void stage_cols(const RhsParams& P, int j0, int n, int Ntotal, float* smem) {
    for (int jj = threadIdx.x; jj < n; jj += blockDim.x) {
        float c[Op::COLF4 * 4];
        Op::pack_col(P, j0 + jj, Ntotal, c);
        float* dst = smem + (size_t)(jj >> 1) * (2 * Op::NF) + (jj & 1);
#pragma unroll
        for (int k = 0; k < Op::NF; ++k) dst[2 * k] = c[k];
    }
}

// acc (packed, NACC) += sum over the n staged columns of Op::pair(row, col)
template <class Op>
DICP_D void sweep_cols(const RhsParams& P, const typename Op::Row& row, const float* smem, int n, F2* acc) {
    constexpr int NF = Op::NF, PF4 = NF / 2;
    const float4* sp = reinterpret_cast<const float4*>(smem);
    const int npair = n >> 1;
#pragma unroll 2
    for (int Pp = 0; Pp < npair; ++Pp) {
        F2 c[NF];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = sp[Pp * PF4 + k];
            c[2 * k] = f2(v.x, v.y);
            c[2 * k + 1] = f2(v.z, v.w);
        }
        Op::template pair<F2>(P, row, c, acc);
    }
    if (n & 1) {
        float c[NF], tmp[Op::NACC];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = sp[npair * PF4 + k];
            c[2 * k] = v.x;
            c[2 * k + 1] = v.z;
        }
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) tmp[k] = 0.f;
        Op::template pair<float>(P, row, c, tmp);
#pragma unroll
        for (int k = 0; k < Op::NACC; ++k) acc[k] = vadd(acc[k], f2(tmp[k], 0.f));
    }
}

DICP_D void small_update(const SmallStep& S, size_t idx) {
    if (S.out == nullptr) return;
    float v = fmaf(S.c_this, S.This[idx], S.base[idx]);
    if (S.other != nullptr) v = fmaf(S.c_other, S.other[idx], v);
    if (S.add != nullptr) v += S.add[idx];
    S.out[idx] = v;
}

// Returns true in every thread of the CTA that took the last ticket of `counter` (out of `total`).
DICP_D bool last_cta(unsigned* counter, unsigned total) {
    __shared__ unsigned ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ticket = atomicAdd(counter, 1u);
    __syncthreads();
    const bool last = ticket == total - 1;
    if (last) __threadfence();
    return last;
}

// ---- forward stage -------------------------------------------------------------------------------------------------
template <int D, bool WLD, bool ETA>
__global__ void __launch_bounds__(kSmallThreads) small_rhs_step_kernel(SmallStep S) {
    using OpQQx = RhsQQ<D, false, ETA, 1>;       // x present: the divergence cost comes from the (x,q) pass
    using OpQQn = RhsQQ<D, WLD, ETA, 1>;         // x absent
    using OpXQ = RhsXQ<D, WLD, ETA, 1>;
    __shared__ __align__(16) float cols[kSmallMaxQ * 2 * D];
    __shared__ float red[32];
    const int tid = threadIdx.x, M = S.M, Nx = S.Nx;
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.vq = S.This; P.dp = S.This + MD; P.vx = S.This + 2 * MD;
    stage_cols<OpXQ>(P, 0, M, M, cols);
    __syncthreads();

    const int nXB = (Nx + kSmallThreads - 1) / kSmallThreads;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    if ((int)blockIdx.x < nXB) {
        const int i = blockIdx.x * kSmallThreads + tid;
        if (i < Nx) {
            typename OpXQ::Row row;
            OpXQ::load_row(P, i, row);
            F2 acc[OpXQ::NACC];
#pragma unroll
            for (int k = 0; k < OpXQ::NACC; ++k) acc[k] = f2(0.f, 0.f);
            sweep_cols<OpXQ>(P, row, cols, M, acc);
            float a[OpXQ::NACC];
#pragma unroll
            for (int k = 0; k < OpXQ::NACC; ++k) a[k] = f2_sum(acc[k]);
            OpXQ::finish(P, i, row, a, &rs[0]);
#pragma unroll
            for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)i * D + k);
        }
    } else {
        const int i = ((int)blockIdx.x - nXB) * kSmallThreads + tid;
        if (i < M) {
            float a[OpQQn::NACC > OpQQx::NACC ? OpQQn::NACC : OpQQx::NACC];
            if (Nx > 0) {
                typename OpQQx::Row row;
                OpQQx::load_row(P, i, row);
                F2 acc[OpQQx::NACC];
#pragma unroll
                for (int k = 0; k < OpQQx::NACC; ++k) acc[k] = f2(0.f, 0.f);
                sweep_cols<OpQQx>(P, row, cols, M, acc);
#pragma unroll
                for (int k = 0; k < OpQQx::NACC; ++k) a[k] = f2_sum(acc[k]);
                OpQQx::finish(P, i, row, a, &rs[1]);
            } else {
                typename OpQQn::Row row;
                OpQQn::load_row(P, i, row);
                F2 acc[OpQQn::NACC];
#pragma unroll
                for (int k = 0; k < OpQQn::NACC; ++k) acc[k] = f2(0.f, 0.f);
                sweep_cols<OpQQn>(P, row, cols, M, acc);
#pragma unroll
                for (int k = 0; k < OpQQn::NACC; ++k) a[k] = f2_sum(acc[k]);
                OpQQn::finish(P, i, row, a, &rs[1]);
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                small_update(S, (size_t)i * D + k);
                small_update(S, MD + (size_t)i * D + k);
            }
        }
    }
    // block partials of the four scalars, then the last CTA sums them in CTA order
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float v = block_sum(rs[k], red);
        if (tid == 0) S.ws[(size_t)blockIdx.x * 4 + k] = v;
    }
    if (last_cta(&S.counters[0], gridDim.x)) {
        // fixed-order parallel sum over CTAs (strided per thread, then the fixed block tree): deterministic
        float tot[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = 0.f;
            for (unsigned b = tid; b < gridDim.x; b += kSmallThreads) v += __ldcg(&S.ws[(size_t)b * 4 + k]);
            tot[k] = block_sum(v, red);
        }
        if (tid == 0) {
            const float A = tot[1], B = tot[2], C = tot[3];
            const float dcost = Nx > 0 ? tot[0] : (WLD ? fmaf(S.eta, C, B) : 0.f);
            S.This[Ssz - 1] = dcost; S.This[Ssz] = A; S.This[Ssz + 1] = B; S.This[Ssz + 2] = C;
            small_update(S, Ssz - 1);
            S.counters[0] = 0u;
        }
    }
}

// ---- adjoint stage ---------------------------------------------------------------------------------------------------
// CTA kinds: [0, nXB) x rows (columns = support set);  then nQB * nsplit q-row CTAs: row block rb, column split sp over
// the data points (plus, for sp == 0, the support-set columns of the (q,q) interaction).
template <int D, bool WLD, bool ETA>
__global__ void __launch_bounds__(kSmallThreads) small_adj_step_kernel(SmallStep S, int nsplit) {
    using OpX = typename std::conditional<ETA, AdjXQxEta<D, 1>, AdjXQx<D, WLD, 1>>::type;       // rows x, cols (q,p)
    using OpQx = typename std::conditional<ETA, AdjXQqEta<D, 1>, AdjXQq<D, WLD, 1>>::type;      // rows q, cols (x,wx)
    using OpQQx = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, false, 1>>::type;     // rows q, cols q; x present
    using OpQQn = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, WLD, 1>>::type;       // x absent
    constexpr int NAQ = OpQQn::NACC, NAX = OpQx::NACC, NPART = NAQ + NAX;
    __shared__ __align__(16) float cols[kSmallMaxQ * 4 * D];      // >= kSmallChunk * 2 * D as well
    const int tid = threadIdx.x, M = S.M, Nx = S.Nx;
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.a = S.lam; P.u = S.lam + MD; P.wx = S.lam + 2 * MD; P.gc = S.lam + (Ssz - 1);
    P.gq = S.This; P.gp = S.This + MD; P.gx = S.This + 2 * MD;
    const int nXB = (Nx + kSmallThreads - 1) / kSmallThreads;

    if ((int)blockIdx.x < nXB) {
        stage_cols<OpX>(P, 0, M, M, cols);
        __syncthreads();
        const int i = blockIdx.x * kSmallThreads + tid;
        if (i < Nx) {
            typename OpX::Row row;
            OpX::load_row(P, i, row);
            F2 acc[OpX::NACC];
#pragma unroll
            for (int k = 0; k < OpX::NACC; ++k) acc[k] = f2(0.f, 0.f);
            sweep_cols<OpX>(P, row, cols, M, acc);
            float a[OpX::NACC];
#pragma unroll
            for (int k = 0; k < OpX::NACC; ++k) a[k] = f2_sum(acc[k]);
            P.accumulate = 0;
            OpX::finish(P, i, row, a, nullptr);
#pragma unroll
            for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)i * D + k);
        }
        if (blockIdx.x == 0 && tid == 0) {          // cost entry: the right-hand side does not depend on cost
            S.This[Ssz - 1] = 0.f;
            small_update(S, Ssz - 1);
        }
        return;
    }

    const int qb = (int)blockIdx.x - nXB;
    const int rb = qb / nsplit, sp = qb % nsplit;
    const int i = rb * kSmallThreads + tid;
    const bool valid = i < M;
    float aq[NAQ], ax[NAX];
#pragma unroll
    for (int k = 0; k < NAQ; ++k) aq[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NAX; ++k) ax[k] = 0.f;

    if (sp == 0) {                                   // (q,q) interaction
        if (Nx > 0) stage_cols<OpQQx>(P, 0, M, M, cols); else stage_cols<OpQQn>(P, 0, M, M, cols);
        __syncthreads();
        if (valid) {
            F2 acc[NAQ];
#pragma unroll
            for (int k = 0; k < NAQ; ++k) acc[k] = f2(0.f, 0.f);
            if (Nx > 0) {
                typename OpQQx::Row row;
                OpQQx::load_row(P, i, row);
                sweep_cols<OpQQx>(P, row, cols, M, acc);
            } else {
                typename OpQQn::Row row;
                OpQQn::load_row(P, i, row);
                sweep_cols<OpQQn>(P, row, cols, M, acc);
            }
#pragma unroll
            for (int k = 0; k < NAQ; ++k) aq[k] = f2_sum(acc[k]);
        }
        __syncthreads();
    }
    if (Nx > 0) {                                    // data-point columns [c0, c1) of this split
        const int per = (Nx + nsplit - 1) / nsplit;
        const int c0 = sp * per, c1 = (c0 + per < Nx) ? c0 + per : Nx;
        typename OpQx::Row row;
        if (valid) OpQx::load_row(P, i, row);
        F2 acc[NAX];
#pragma unroll
        for (int k = 0; k < NAX; ++k) acc[k] = f2(0.f, 0.f);
        for (int j0 = c0; j0 < c1; j0 += kSmallChunk) {
            const int n = (c1 - j0 < kSmallChunk) ? c1 - j0 : kSmallChunk;
            stage_cols<OpQx>(P, j0, n, Nx, cols);
            __syncthreads();
            if (valid) sweep_cols<OpQx>(P, row, cols, n, acc);
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < NAX; ++k) ax[k] = f2_sum(acc[k]);
    }

    // partials -> workspace [split][k][row]; the last CTA of this row block merges them in split order
    float* part = S.ws;
    if (valid) {
#pragma unroll
        for (int k = 0; k < NAQ; ++k) part[((size_t)sp * NPART + k) * M + i] = aq[k];
#pragma unroll
        for (int k = 0; k < NAX; ++k) part[((size_t)sp * NPART + NAQ + k) * M + i] = ax[k];
    }
    if (!last_cta(&S.counters[1 + rb], (unsigned)nsplit)) return;
    // merge: every (row, accumulator) item of this row block is summed over the splits, in split order, by one of the 128
    // threads (independent loads in flight), staged in shared memory, then each row thread finishes its row
    {
        float* merged = cols;                                     // reuse: kSmallThreads * NPART floats
        const int rows_here = (M - rb * kSmallThreads < kSmallThreads) ? M - rb * kSmallThreads : kSmallThreads;
        for (int item = tid; item < rows_here * NPART; item += kSmallThreads) {
            const int rr = item / NPART, k = item - rr * NPART;
            const size_t col = (size_t)k * M + (size_t)(rb * kSmallThreads + rr);
            float v;
            if (k < NAQ) {
                v = __ldcg(&part[col]);                           // the (q,q) part lives in split 0
            } else {
                v = 0.f;
                if (Nx > 0)
                    for (int s2 = 0; s2 < nsplit; ++s2) v += __ldcg(&part[(size_t)s2 * NPART * M + col]);
            }
            merged[item] = v;
        }
        __syncthreads();
        if (valid) {
#pragma unroll
            for (int k = 0; k < NAQ; ++k) aq[k] = merged[tid * NPART + k];
#pragma unroll
            for (int k = 0; k < NAX; ++k) ax[k] = merged[tid * NPART + NAQ + k];
        }
    }
    if (valid) {
        P.accumulate = 0;
        if (Nx > 0) {
            typename OpQQx::Row row;
            OpQQx::load_row(P, i, row);
            OpQQx::finish(P, i, row, aq, nullptr);
            typename OpQx::Row rowx;
            OpQx::load_row(P, i, rowx);
            P.accumulate = 1;
            OpQx::finish(P, i, rowx, ax, nullptr);
        } else {
            typename OpQQn::Row row;
            OpQQn::load_row(P, i, row);
            OpQQn::finish(P, i, row, aq, nullptr);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            small_update(S, (size_t)i * D + k);
            small_update(S, MD + (size_t)i * D + k);
        }
    }
    if (Nx == 0 && rb == 0 && tid == 0) {            // no x CTA exists: handle the cost entry here
        S.This[Ssz - 1] = 0.f;
        small_update(S, Ssz - 1);
    }
    if (tid == 0) S.counters[1 + rb] = 0u;
}

// Column splits of the q-row CTAs: chunks of about sqrt(1.6 Nx) data points (clamped to [64, kSmallChunk]) balance the
// serial sweep of a chunk against the serial merge of the splits.
inline int small_adj_nsplit(int Nx) {
    if (Nx <= 0) return 1;
    int chunk = (int)sqrtf(1.6f * (float)Nx);
    if (chunk < 64) chunk = 64;
    if (chunk > kSmallChunk) chunk = kSmallChunk;
    int n = (Nx + chunk - 1) / chunk;
    return n < 1 ? 1 : n;
}

// workspace bytes: counters + max(forward block scalars, adjoint partials)
inline size_t small_workspace_bytes(long long M, long long Nx) {
    const long long nXB = (Nx + kSmallThreads - 1) / kSmallThreads, nQB = (M + kSmallThreads - 1) / kSmallThreads;
    const size_t fwd = (size_t)(nXB + nQB) * 4 * 4;
    const size_t adj = (size_t)small_adj_nsplit((int)Nx) * 16 * (size_t)M * 4;
    return kSmallCounters * 4 + (fwd > adj ? fwd : adj) + 256;
}

}  // namespace dicp
