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
#include "sym_engine.cuh"
#include <cstdlib>

namespace dicp {

static constexpr int kSmallMaxQ = 2048;       // support points staged entirely in (dynamic) shared memory (forward stage: 48 KB in 3-D)
static constexpr int kSmallThreads = 128;     // rows per CTA (one row per thread)
#ifndef DICP_SMALL_MINB_FWD
#define DICP_SMALL_MINB_FWD 12                // minimum resident CTAs per SM asked of the compiler (register cap); swept on B200
#endif
#ifndef DICP_SMALL_MINB_ADJ
#define DICP_SMALL_MINB_ADJ 8
#endif
// Mid-size supports (more than kRingMaxQ = 64 points): a thread sweeps hundreds of columns for 4 rows at a time, which needs
// ~100 registers -- under the caps above the 4-row sweep spills (224-688 bytes of stack per thread) and the forward stage ran at
// a quarter of the FP32 roof.  The BIG instantiation asks for 4 resident CTAs (128 registers), like the tiled engine.
static constexpr int kSmallMinbBig = 4;
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
    // Batched form (dims != nullptr; see batch_closure.cuh): blockIdx.y = frame k.  Every pointer above addresses frame 0
    // and advances by `fstride` floats per frame, the workspace (counters included) by `ws_fstride` bytes; M / Nx above
    // are replaced by the frame's own sizes.  Frames with active[k] == 0 are skipped.
    const int* dims;         // (K,2): M_k, Nx_k
    const int* active;       // (K), nullable
    long long fstride;
    long long ws_fstride;
};

// Selects the frame of this CTA in the batched form; false => nothing to do for this CTA's frame.
DICP_D bool small_select_frame(SmallStep& S) {
    if (S.dims == nullptr) return true;
    const int k = blockIdx.y;
    if (S.active != nullptr && S.active[k] == 0) return false;
    S.M = S.dims[2 * k];
    S.Nx = S.dims[2 * k + 1];
    const long long o = (long long)k * S.fstride;
    S.s_eval += o;
    S.This += o;
    if (S.lam) S.lam += o;
    if (S.base) S.base += o;
    if (S.other) S.other += o;
    if (S.add) S.add += o;
    if (S.out) S.out += o;
    S.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(S.ws) + (long long)k * S.ws_fstride);
    S.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(S.counters) + (long long)k * S.ws_fstride);
    return true;
}

// Column splits of the q-row CTAs: chunks of about sqrt(1.6 Nx G) data points (clamped to [64, kSmallChunk]) balance the
// serial sweep of a chunk (by G column groups) against the serial merge of the splits.  Integer arithmetic only: host (grid size, workspace)
// and device (batched form) must agree exactly.
DICP_HD int small_adj_nsplit(int Nx, int G = 1) {
    if (Nx <= 0) return 1;
    const long long v = ((long long)Nx * 16 * G) / 10;
    long long r = (long long)sqrtf((float)v);      // floor(sqrt(v)) made exact by the two integer corrections
    while (r * r > v) --r;
    while ((r + 1) * (r + 1) <= v) ++r;
    int chunk = (int)r;
    if (chunk < 64) chunk = 64;
    if (chunk > kSmallChunk) chunk = kSmallChunk;
    const int n = (Nx + chunk - 1) / chunk;
    return n < 1 ? 1 : n;
}
// column groups of a q-row CTA: with M <= 64 support points the 128 threads form 128 / M groups, each sweeping its own share
// of the staged columns (G times faster sweeps => G times longer chunks in small_adj_nsplit)
DICP_HD int small_adj_groups(int M) { return M <= kSmallThreads / 2 ? kSmallThreads / (M < 1 ? 1 : M) : 1; }

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

// acc (packed, NACC) += sum over the staged column pairs [P0, P1) of Op::pair(row, col); with `tail`, also the odd last
// column (stored alone in pair slot `tailpair`).
template <class Op>
DICP_D void sweep_range(const RhsParams& P, const typename Op::Row& row, const float* smem, int P0, int P1, bool tail,
                        int tailpair, F2* acc) {
    constexpr int NF = Op::NF, PF4 = NF / 2;
    const float4* sp = reinterpret_cast<const float4*>(smem);
#pragma unroll 2
    for (int Pp = P0; Pp < P1; ++Pp) {
        F2 c[NF];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = sp[Pp * PF4 + k];
            c[2 * k] = f2(v.x, v.y);
            c[2 * k + 1] = f2(v.z, v.w);
        }
        Op::template pair<F2>(P, row, c, acc);
    }
    if (tail) {
        float c[NF], tmp[Op::NACC];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = sp[tailpair * PF4 + k];
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

// R rows at once against all n staged columns: one broadcast LDS.128 set per column pair serves the R rows, and the R
// independent accumulator chains give the scheduler something to overlap
template <class Op, int R>
DICP_D void sweep_cols_multi(const RhsParams& P, const typename Op::Row (&row)[R], const float* smem, int n,
                             F2 (&acc)[R][Op::NACC]) {
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
#pragma unroll
        for (int r = 0; r < R; ++r) Op::template pair<F2>(P, row[r], c, acc[r]);
    }
    if (n & 1) {
        float c[NF];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = sp[npair * PF4 + k];
            c[2 * k] = v.x;
            c[2 * k + 1] = v.z;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float tmp[Op::NACC];
#pragma unroll
            for (int k = 0; k < Op::NACC; ++k) tmp[k] = 0.f;
            Op::template pair<float>(P, row[r], c, tmp);
#pragma unroll
            for (int k = 0; k < Op::NACC; ++k) acc[r][k] = vadd(acc[r][k], f2(tmp[k], 0.f));
        }
    }
}

// all n staged columns
template <class Op>
DICP_D void sweep_cols(const RhsParams& P, const typename Op::Row& row, const float* smem, int n, F2* acc) {
    sweep_range<Op>(P, row, smem, 0, n >> 1, (n & 1) != 0, n >> 1, acc);
}

// group g of G: an even share of the n staged columns (whole pairs; the odd last column goes to the last group)
template <class Op>
DICP_D void sweep_share(const RhsParams& P, const typename Op::Row& row, const float* smem, int n, int g, int G, F2* acc) {
    const int npair = n >> 1;
    sweep_range<Op>(P, row, smem, (int)(((long long)npair * g) / G), (int)(((long long)npair * (g + 1)) / G),
                    (n & 1) != 0 && g == G - 1, npair, acc);
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
// CTAs [0, nXB): x rows, `xpass` consecutive blocks of 128 rows each (more rows per CTA amortise the staging of the support
// set, the block reductions and the ticket when many frames / data points make the grid large); then the q-row CTAs.
// R rows (i0, i0 + 128, ...) of one thread of the forward x-row pass: vx, the fused state update, and the rows' dcost sum
template <class OpXQ, int D, int R>
DICP_D float small_fwd_x_rows(const SmallStep& S, const RhsParams& P, const float* cols, int M, int Nx, size_t MD, int i0) {
    typename OpXQ::Row row[R];
    F2 acc[R][OpXQ::NACC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r * kSmallThreads;
        OpXQ::load_row(P, i < Nx ? i : Nx - 1, row[r]);
#pragma unroll
        for (int k = 0; k < OpXQ::NACC; ++k) acc[r][k] = f2(0.f, 0.f);
    }
    sweep_cols_multi<OpXQ, R>(P, row, cols, M, acc);
    float dcs = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r * kSmallThreads;
        if (i < Nx) {
            float a[OpXQ::NACC];
#pragma unroll
            for (int k = 0; k < OpXQ::NACC; ++k) a[k] = f2_sum(acc[r][k]);
            float dc = 0.f;                 // finish ASSIGNS the row's dcost contribution
            OpXQ::finish(P, i, row[r], a, &dc);
            dcs += dc;
#pragma unroll
            for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)i * D + k);
        }
    }
    return dcs;
}

template <int D, bool WLD, bool ETA, bool BIG = false>
__global__ void __launch_bounds__(kSmallThreads, BIG ? kSmallMinbBig : DICP_SMALL_MINB_FWD) small_rhs_step_kernel(SmallStep S, int xpass) {
    using OpQQx = RhsQQ<D, false, ETA, 1>;       // x present: the divergence cost comes from the (x,q) pass
    using OpQQn = RhsQQ<D, WLD, ETA, 1>;         // x absent
    using OpXQ = RhsXQ<D, WLD, ETA, 1>;
    extern __shared__ __align__(16) float cols[];                // small_fwd_smem_bytes(max M, D)
    __shared__ float red[32];
    if (!small_select_frame(S)) return;
    const int tid = threadIdx.x, M = S.M, Nx = S.Nx;
    const int nXB = (Nx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass);
    const unsigned nblk = (unsigned)(nXB + (M + kSmallThreads - 1) / kSmallThreads);    // CTAs of this frame
    if (blockIdx.x >= nblk) return;
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.vq = S.This; P.dp = S.This + MD; P.vx = S.This + 2 * MD;
    stage_cols<OpXQ>(P, 0, M, M, cols);
    __syncthreads();

    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    if ((int)blockIdx.x < nXB) {
        // the CTA's xpass blocks of 128 rows: 4, then 2, then 1 rows of a thread swept together
        int ps = 0;
        for (; ps + 4 <= xpass; ps += 4)
            rs[0] += small_fwd_x_rows<OpXQ, D, 4>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid);
        if (ps + 2 <= xpass) {
            rs[0] += small_fwd_x_rows<OpXQ, D, 2>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid);
            ps += 2;
        }
        if (ps < xpass)
            rs[0] += small_fwd_x_rows<OpXQ, D, 1>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid);
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
    if (last_cta(&S.counters[0], nblk)) {
        // fixed-order parallel sum over CTAs (strided per thread, then the fixed block tree): deterministic
        float tot[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = 0.f;
            for (unsigned b = tid; b < nblk; b += kSmallThreads) v += __ldcg(&S.ws[(size_t)b * 4 + k]);
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
// CTA kinds: [0, nXB) x rows (columns = support set), `xpass` blocks of 128 rows each;  then nQB * nsplit q-row CTAs: row
// block rb, column split sp over the data points (plus, for sp == 0, the support-set columns of the (q,q) interaction).
// In a q-row CTA with Mr <= 64 rows the 128 threads form G = 128 / Mr groups; every group sweeps its own share of the
// staged columns for all Mr rows and the group results are added in group order through shared memory, so that all lanes
// work even when the support set is tiny (25 points: G = 5).
// R rows (i0, i0 + 128, ...) of one thread of the adjoint x-row pass: gx and the fused cotangent update of those rows
template <class OpX, int D, int R>
DICP_D void small_adj_x_rows(const SmallStep& S, const RhsParams& P, const float* cols, int M, int Nx, size_t MD, int i0) {
    typename OpX::Row row[R];
    F2 acc[R][OpX::NACC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r * kSmallThreads;
        OpX::load_row(P, i < Nx ? i : Nx - 1, row[r]);
#pragma unroll
        for (int k = 0; k < OpX::NACC; ++k) acc[r][k] = f2(0.f, 0.f);
    }
    sweep_cols_multi<OpX, R>(P, row, cols, M, acc);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = i0 + r * kSmallThreads;
        if (i < Nx) {
            float a[OpX::NACC];
#pragma unroll
            for (int k = 0; k < OpX::NACC; ++k) a[k] = f2_sum(acc[r][k]);
            OpX::finish(P, i, row[r], a, nullptr);
#pragma unroll
            for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)i * D + k);
        }
    }
}

template <int D, bool WLD, bool ETA>
__global__ void __launch_bounds__(kSmallThreads, DICP_SMALL_MINB_ADJ) small_adj_step_kernel(SmallStep S, int nsplit, int xpass) {
    using OpX = typename std::conditional<ETA, AdjXQxEta<D, 1>, AdjXQx<D, WLD, 1>>::type;       // rows x, cols (q,p)
    using OpQx = typename std::conditional<ETA, AdjXQqEta<D, 1>, AdjXQq<D, WLD, 1>>::type;      // rows q, cols (x,wx)
    using OpQQx = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, false, 1>>::type;     // rows q, cols q; x present
    using OpQQn = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, WLD, 1>>::type;       // x absent
    constexpr int NAQ = OpQQn::NACC, NAX = OpQx::NACC, NPART = NAQ + NAX;
    extern __shared__ __align__(16) float cols[];                // small_adj_smem_bytes(max M, D)
    __shared__ float xch[kSmallThreads * NPART];
    const int maxM = S.M;                             // batched form: bound of the frames' support sizes
    if (!small_select_frame(S)) return;
    const int tid = threadIdx.x, M = S.M, Nx = S.Nx;
    const int nXB = (Nx + kSmallThreads * xpass - 1) / (kSmallThreads * xpass);
    if (S.dims != nullptr) {                          // batched form: this frame's own split count and CTA count
        nsplit = small_adj_nsplit(Nx, small_adj_groups(maxM));
        if (blockIdx.x >= (unsigned)(nXB + ((M + kSmallThreads - 1) / kSmallThreads) * nsplit)) return;
    }
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.a = S.lam; P.u = S.lam + MD; P.wx = S.lam + 2 * MD; P.gc = S.lam + (Ssz - 1);
    P.gq = S.This; P.gp = S.This + MD; P.gx = S.This + 2 * MD;

    if ((int)blockIdx.x < nXB) {
        stage_cols<OpX>(P, 0, M, M, cols);
        __syncthreads();
        P.accumulate = 0;
        // the CTA's xpass blocks of 128 rows: 4, then 2, then 1 rows of a thread swept together (one broadcast LDS per column
        // pair serves the rows, independent accumulator chains), like the forward stage
        int ps = 0;
        for (; ps + 4 <= xpass; ps += 4) small_adj_x_rows<OpX, D, 4>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid);
        if (ps + 2 <= xpass) { small_adj_x_rows<OpX, D, 2>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid); ps += 2; }
        if (ps < xpass) small_adj_x_rows<OpX, D, 1>(S, P, cols, M, Nx, MD, ((int)blockIdx.x * xpass + ps) * kSmallThreads + tid);
        if (blockIdx.x == 0 && tid == 0) {          // cost entry: the right-hand side does not depend on cost
            S.This[Ssz - 1] = 0.f;
            small_update(S, Ssz - 1);
        }
        return;
    }

    const int qb = (int)blockIdx.x - nXB;
    const int rb = qb / nsplit, sp = qb % nsplit;
    // thread -> (column group g, row r)
    const int Mr = (M <= kSmallThreads / 2) ? M : kSmallThreads;        // rows handled by this CTA's groups
    const int G = kSmallThreads / Mr;
    const int g = tid / Mr, r = tid - g * Mr;
    const int i = rb * kSmallThreads + r;
    const bool work = g < G && i < M;                // sweeps columns
    const bool valid = work && g == 0;               // owns the row's results
    float aq[NAQ], ax[NAX];
#pragma unroll
    for (int k = 0; k < NAQ; ++k) aq[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NAX; ++k) ax[k] = 0.f;

    if (sp == 0) {                                   // (q,q) interaction
        if (Nx > 0) stage_cols<OpQQx>(P, 0, M, M, cols); else stage_cols<OpQQn>(P, 0, M, M, cols);
        __syncthreads();
        if (work) {
            F2 acc[NAQ];
#pragma unroll
            for (int k = 0; k < NAQ; ++k) acc[k] = f2(0.f, 0.f);
            if (Nx > 0) {
                typename OpQQx::Row row;
                OpQQx::load_row(P, i, row);
                sweep_share<OpQQx>(P, row, cols, M, g, G, acc);
            } else {
                typename OpQQn::Row row;
                OpQQn::load_row(P, i, row);
                sweep_share<OpQQn>(P, row, cols, M, g, G, acc);
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
        if (work) OpQx::load_row(P, i, row);
        F2 acc[NAX];
#pragma unroll
        for (int k = 0; k < NAX; ++k) acc[k] = f2(0.f, 0.f);
        for (int j0 = c0; j0 < c1; j0 += kSmallChunk) {
            const int n = (c1 - j0 < kSmallChunk) ? c1 - j0 : kSmallChunk;
            stage_cols<OpQx>(P, j0, n, Nx, cols);
            __syncthreads();
            if (work) sweep_share<OpQx>(P, row, cols, n, g, G, acc);
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < NAX; ++k) ax[k] = f2_sum(acc[k]);
    }
    if (G > 1) {                                     // add the groups' results in group order
        if (work && g > 0) {
#pragma unroll
            for (int k = 0; k < NAQ; ++k) xch[(g * NPART + k) * Mr + r] = aq[k];
#pragma unroll
            for (int k = 0; k < NAX; ++k) xch[(g * NPART + NAQ + k) * Mr + r] = ax[k];
        }
        __syncthreads();
        if (valid) {
            for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
                for (int k = 0; k < NAQ; ++k) aq[k] += xch[(g2 * NPART + k) * Mr + r];
#pragma unroll
                for (int k = 0; k < NAX; ++k) ax[k] += xch[(g2 * NPART + NAQ + k) * Mr + r];
            }
        }
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
                if (Nx > 0) {
                    int s2 = 0;
                    for (; s2 + 8 <= nsplit; s2 += 8) {           // 8 independent loads in flight, added in split order
                        float t[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = __ldcg(&part[(size_t)(s2 + u) * NPART * M + col]);
#pragma unroll
                        for (int u = 0; u < 8; ++u) v += t[u];
                    }
                    for (; s2 < nsplit; ++s2) v += __ldcg(&part[(size_t)s2 * NPART * M + col]);
                }
            }
            merged[item] = v;
        }
        __syncthreads();
        if (valid) {
#pragma unroll
            for (int k = 0; k < NAQ; ++k) aq[k] = merged[r * NPART + k];
#pragma unroll
            for (int k = 0; k < NAX; ++k) ax[k] = merged[r * NPART + NAQ + k];
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

// ---- adjoint stage, ring form (data points present, M <= 64) -------------------------------------------------------
// The x rows need sum_j over the support points (gx_k) and the support points need sum_k over the x rows (gq_j, gp_j) of terms
// built from the SAME K, z', dot products: small_adj_step_kernel evaluates every (x_k, q_j) pair twice (x-row CTAs with
// AdjXQx, q-row CTAs with AdjXQq over column splits of the data).  Here every pair is evaluated ONCE (Op AdjXQ, both sides):
// a lane owns 4 x rows; the support points -- at most 32 column pairs -- travel round a ring of W = 16 or 32 lanes together
// with their accumulators (see sym_engine.cuh), so after W steps lane l holds column pair l summed over the ring's rows; the
// rings of a CTA are added in ring order through shared memory and the CTA's column sums go to the workspace.  One more CTA
// does the (q,q) interaction.  The last CTA of the frame (ticket) adds the CTAs' column sums in CTA order, finishes the
// support points' gradients and applies their cotangent update.  Deterministic, no atomics.
static constexpr int kRingR = 4;                                   // x rows per lane
static constexpr int kRingRows = kSmallThreads * kRingR;           // x rows per CTA
static constexpr int kRingMaxQ = 64;                               // support points (one ring round)

template <int D, bool WLD, bool ETA>
__global__ void __launch_bounds__(kSmallThreads) small_adj_ring_kernel(SmallStep S, int xpass) {
    using OpX = typename std::conditional<ETA, AdjXQE<D>, AdjXQ<D, WLD>>::type;
    // x present: no divergence term in the (q,q) pass (AdjQQEta reads gc = 0 itself when x is given)
    using OpQQ = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, false, 1>>::type;
    constexpr int NF = OpX::NF, REC = 2 * NF, STRIDE = (REC % 8 == 4) ? REC : REC + 4, NC = OpX::NACC_COL;
    constexpr int NAQ = OpQQ::NACC;
    __shared__ __align__(16) float cols[32 * STRIDE > kRingMaxQ * 4 * D ? 32 * STRIDE : kRingMaxQ * 4 * D];
    __shared__ float xch[kSmallThreads * (2 * NC > NAQ ? 2 * NC : NAQ)];
    if (!small_select_frame(S)) return;
    const int tid = threadIdx.x, lane = tid & 31, M = S.M, Nx = S.Nx;
    // an x CTA takes `xpass` consecutive blocks of kRingRows rows (host: chosen so that all CTAs of all frames are resident at once)
    const int nXB = (Nx + kRingRows * xpass - 1) / (kRingRows * xpass);
    const unsigned nblk = (unsigned)nXB + 1u;
    if (blockIdx.x >= nblk) return;
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.a = S.lam; P.u = S.lam + MD; P.wx = S.lam + 2 * MD; P.gc = S.lam + (Ssz - 1);
    P.gq = S.This; P.gp = S.This + MD; P.gx = S.This + 2 * MD;
    P.accumulate = 0;
    float* part = S.ws;                                            // [x CTA][k][2W columns]
    const int W = M <= 32 ? 16 : 32;

    if ((int)blockIdx.x < nXB) {
        // stage the support points as zero-padded pair records
        for (int t = tid; t < 32 * STRIDE; t += kSmallThreads) cols[t] = 0.f;
        __syncthreads();
        for (int j = tid; j < M; j += kSmallThreads) {
            float c[OpX::COLF4 * 4];
            OpX::pack_col(P, j, M, c);
            float* dst = cols + (j >> 1) * STRIDE + (j & 1);
#pragma unroll
            for (int k = 0; k < NF; ++k) dst[2 * k] = c[k];
        }
        __syncthreads();
        F2 cacc[NC];                                  // column sums: kept over the passes (W steps bring every pair home)
#pragma unroll
        for (int k = 0; k < NC; ++k) cacc[k] = f2(0.f, 0.f);
        const int lr = lane & (W - 1);
      for (int ps = 0; ps < xpass; ++ps) {
        typename OpX::Row row[kRingR];
        float rmask[kRingR];
        int ri[kRingR];
#pragma unroll
        for (int r = 0; r < kRingR; ++r) {
            ri[r] = (blockIdx.x * xpass + ps) * kRingRows + (tid >> 5) * (32 * kRingR) + r * 32 + lane;
            rmask[r] = ri[r] < Nx ? 1.f : 0.f;
            OpX::load_row(P, ri[r] < Nx ? ri[r] : Nx - 1, row[r]);
        }
        F2 acc[kRingR][OpX::NACC];
#pragma unroll
        for (int r = 0; r < kRingR; ++r)
#pragma unroll
            for (int k = 0; k < OpX::NACC; ++k) acc[r][k] = f2(0.f, 0.f);
        for (int s = 0; s < W; ++s) {
            const int pr = (lr + s) & (W - 1);
            const float4* rec = reinterpret_cast<const float4*>(cols + pr * STRIDE);
            F2 c[NF];
#pragma unroll
            for (int k = 0; k < NF / 2; ++k) {
                const float4 v = rec[k];
                c[2 * k] = f2(v.x, v.y);
                c[2 * k + 1] = f2(v.z, v.w);
            }
            const float m0 = 2 * pr < M ? 1.f : 0.f, m1 = 2 * pr + 1 < M ? 1.f : 0.f;
#pragma unroll
            for (int r = 0; r < kRingR; ++r)
                OpX::template pair_sym<F2, true>(P, row[r], c, acc[r], cacc, f2(m0 * rmask[r], m1 * rmask[r]));
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                unsigned lo = (unsigned)(cacc[k].v & 0xffffffffull), hi = (unsigned)(cacc[k].v >> 32);
                lo = __shfl_sync(0xffffffffu, lo, (lr + 1) & (W - 1), W);
                hi = __shfl_sync(0xffffffffu, hi, (lr + 1) & (W - 1), W);
                cacc[k].v = ((unsigned long long)hi << 32) | lo;
            }
        }
        // rows: gx and the cotangent update of the x entries
#pragma unroll
        for (int r = 0; r < kRingR; ++r) {
            if (ri[r] < Nx) {
                float a[OpX::NACC];
#pragma unroll
                for (int k = 0; k < OpX::NACC; ++k) a[k] = f2_sum(acc[r][k]);
                OpX::finish(P, ri[r], row[r], a, nullptr);
#pragma unroll
                for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)ri[r] * D + k);
            }
        }
      }
        // columns: lane lr of every ring holds pair lr; add the CTA's rings in ring order
        const int ring = tid / W, nring = kSmallThreads / W;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float a, b;
            f2_unpack(cacc[k], a, b);
            xch[((ring * NC + k) * W + lr) * 2] = a;
            xch[((ring * NC + k) * W + lr) * 2 + 1] = b;
        }
        __syncthreads();
        for (int t = tid; t < NC * 2 * W; t += kSmallThreads) {
            const int k = t / (2 * W), cc = t - k * 2 * W;
            float v = 0.f;
            for (int g2 = 0; g2 < nring; ++g2) v += xch[((g2 * NC + k) * W + (cc >> 1)) * 2 + (cc & 1)];
            part[((size_t)blockIdx.x * NC + k) * kRingMaxQ + cc] = v;
        }
        if (blockIdx.x == 0 && tid == 0) {            // cost entry: the right-hand side does not depend on cost
            S.This[Ssz - 1] = 0.f;
            small_update(S, Ssz - 1);
        }
    } else {
        // (q,q) interaction: rows = support points, column groups as in small_adj_step_kernel
        const int Mr = (M <= kSmallThreads / 2) ? M : kSmallThreads;
        const int G = kSmallThreads / Mr;
        const int g = tid / Mr, r = tid - g * Mr;
        const bool work = g < G && r < M;
        stage_cols<OpQQ>(P, 0, M, M, cols);
        __syncthreads();
        float aq[NAQ];
#pragma unroll
        for (int k = 0; k < NAQ; ++k) aq[k] = 0.f;
        if (work) {
            F2 acc[NAQ];
#pragma unroll
            for (int k = 0; k < NAQ; ++k) acc[k] = f2(0.f, 0.f);
            typename OpQQ::Row row;
            OpQQ::load_row(P, r, row);
            sweep_share<OpQQ>(P, row, cols, M, g, G, acc);
#pragma unroll
            for (int k = 0; k < NAQ; ++k) aq[k] = f2_sum(acc[k]);
        }
        if (work && g > 0) {
#pragma unroll
            for (int k = 0; k < NAQ; ++k) xch[(g * NAQ + k) * Mr + r] = aq[k];
        }
        __syncthreads();
        if (work && g == 0) {
            for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
                for (int k = 0; k < NAQ; ++k) aq[k] += xch[(g2 * NAQ + k) * Mr + r];
            }
            typename OpQQ::Row row;
            OpQQ::load_row(P, r, row);
            OpQQ::finish(P, r, row, aq, nullptr);
        }
    }
    if (!last_cta(&S.counters[0], nblk)) return;
    // last CTA of the frame: column sums of the x CTAs in CTA order, then the support points' gradients and update
    for (int j = tid; j < M; j += kSmallThreads) {
        float cs[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) cs[k] = 0.f;
        for (int b = 0; b < nXB; ++b) {
#pragma unroll
            for (int k = 0; k < NC; ++k) cs[k] += __ldcg(&part[((size_t)b * NC + k) * kRingMaxQ + j]);
        }
        OpX::finish_col(P, j, cs);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            small_update(S, (size_t)j * D + k);
            small_update(S, MD + (size_t)j * D + k);
        }
    }
    if (tid == 0) S.counters[0] = 0u;
}

// ---- adjoint stage, mid-size supports (kRingMaxQ < M <= kSmallMaxQ, data points present) ---------------------------------
// The ring form above generalised to any number of column groups: an x CTA owns 128 R data points (R = 4 per lane, or 2
// when the 512-row CTAs of all frames would be too few waves: launch_small_mid in capi.cu) and
// meets ALL support points, 64 at a time (one ring round of sym_ring_round per group of 32 column pairs: every (x_k, q_j) pair
// is evaluated once, Op AdjXQ / AdjXQE, row AND column side).  The group's records are packed straight from the state vector
// into a double-buffered shared-memory tile (the loads of group g + 1 are in flight during the ring round of group g; no pack
// kernel, no packed copy in HBM).  Rows finish in the kernel (gx, cotangent update of the x entries); the column sums of the
// CTA's four warps are added in warp order and go to the workspace [x CTA][k][column].  nQB further CTAs do the (q,q)
// interaction (one support point per thread, the columns staged in chunks) and write gq, gp.  A second, tiny launch
// (small_mid_finish_kernel) adds the x CTAs' column sums in CTA order, finishes the support points' gradients and applies their
// cotangent update.  Two launches per adjoint stage; before, mid-size supports paid every (x,q) pair twice
// (small_adj_step_kernel) or ~15 launches per stage (tiled engine).  Deterministic, no atomics.
#ifndef DICP_MID_MINB
#define DICP_MID_MINB kSmallMinbBig         // resident CTAs asked for the mid-size adjoint stage; measured on B200: 3 (168 registers)
                                            // instead of 4 (128) changes nothing (Reg_opt of 16 configs[3] frames 919 vs 914 ms)
#endif
static constexpr int kMidQChunk = 256;                             // support columns per staged chunk of a (q,q) CTA

template <int D, bool WLD, bool ETA, int R>
__global__ void __launch_bounds__(kSmallThreads, DICP_MID_MINB) small_adj_mid_kernel(SmallStep S) {
    using OpX = typename std::conditional<ETA, AdjXQE<D>, AdjXQ<D, WLD>>::type;
    using OpQQ = typename std::conditional<ETA, AdjQQEta<D, 1>, AdjQQ<D, false, 1>>::type;      // x present: see the ring form
    constexpr int NF = OpX::NF, REC = 2 * NF, STRIDE = SymStride<REC>::value, NC = OpX::NACC_COL, NAQ = OpQQ::NACC;
    constexpr int NW = kSmallThreads / 32, TILE = 32 * STRIDE, ROWS = kSmallThreads * R;      // ROWS data points per x CTA
    constexpr int SM_X = 2 * TILE + NW * NC * kSymGroup, SM_Q = kMidQChunk * OpQQ::NF;
    __shared__ __align__(16) float sm[SM_X > SM_Q ? SM_X : SM_Q];
    if (!small_select_frame(S)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, M = S.M, Nx = S.Nx;
    const int nXB = (Nx + ROWS - 1) / ROWS;
    if ((int)blockIdx.x >= nXB + (M + kSmallThreads - 1) / kSmallThreads) return;
    const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
    RhsParams P = small_params<D>(S);
    P.a = S.lam; P.u = S.lam + MD; P.wx = S.lam + 2 * MD; P.gc = S.lam + (Ssz - 1);
    P.gq = S.This; P.gp = S.This + MD; P.gx = S.This + 2 * MD;
    P.accumulate = 0;

    if ((int)blockIdx.x < nXB) {
        const int ngroups = (M + kSymGroup - 1) / kSymGroup, mpad = ngroups * kSymGroup;
        float* xch = sm + 2 * TILE;
        float* part = S.ws + (size_t)blockIdx.x * NC * mpad;      // [k][column]
        typename OpX::Row row[R];
        float rmask[R];
        int ri[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ri[r] = (int)blockIdx.x * ROWS + warp * (32 * R) + r * 32 + lane;
            rmask[r] = ri[r] < Nx ? 1.f : 0.f;
            OpX::load_row(P, ri[r] < Nx ? ri[r] : Nx - 1, row[r]);
        }
        F2 acc[R][OpX::NACC];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < OpX::NACC; ++k) acc[r][k] = f2(0.f, 0.f);
        const bool rows_ragged = ((int)blockIdx.x + 1) * ROWS > Nx;
        // threads 0..63 pack one column each (zero record beyond M)
        float creg[OpX::COLF4 * 4];
        if (tid < kSymGroup) {
            OpX::pack_col(P, tid, M, creg);
            float* dst = sm + (tid >> 1) * STRIDE + (tid & 1);
#pragma unroll
            for (int k = 0; k < NF; ++k) dst[2 * k] = creg[k];
        }
        for (int g = 0; g < ngroups; ++g) {
            __syncthreads();                         // tile g complete; the previous group's column sums have been read
            const bool more = g + 1 < ngroups;
            if (more && tid < kSymGroup) OpX::pack_col(P, (g + 1) * kSymGroup + tid, M, creg);       // in flight during the round
            const float* tile = sm + (g & 1) * TILE;
            const int g0 = g * kSymGroup;
            F2 cacc[NC];
#pragma unroll
            for (int k = 0; k < NC; ++k) cacc[k] = f2(0.f, 0.f);
            if (rows_ragged || g0 + kSymGroup > M) sym_ring_round<OpX, true, R>(P, row, tile, STRIDE, lane, rmask, g0, M, acc, cacc);
            else sym_ring_round<OpX, false, R>(P, row, tile, STRIDE, lane, rmask, g0, M, acc, cacc);
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                float a, b;
                f2_unpack(cacc[k], a, b);
                xch[((warp * NC + k) * 32 + lane) * 2] = a;
                xch[((warp * NC + k) * 32 + lane) * 2 + 1] = b;
            }
            if (more && tid < kSymGroup) {
                float* dst = sm + ((g + 1) & 1) * TILE + (tid >> 1) * STRIDE + (tid & 1);
#pragma unroll
                for (int k = 0; k < NF; ++k) dst[2 * k] = creg[k];
            }
            __syncthreads();
            for (int t = tid; t < NC * kSymGroup; t += kSmallThreads) {
                const int k = t / kSymGroup, cc = t - k * kSymGroup;
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) v += xch[((w * NC + k) * 32 + (cc >> 1)) * 2 + (cc & 1)];
                part[(size_t)k * mpad + g0 + cc] = v;
            }
        }
        // rows: gx and the cotangent update of the x entries
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (ri[r] < Nx) {
                float a[OpX::NACC];
#pragma unroll
                for (int k = 0; k < OpX::NACC; ++k) a[k] = f2_sum(acc[r][k]);
                OpX::finish(P, ri[r], row[r], a, nullptr);
#pragma unroll
                for (int k = 0; k < D; ++k) small_update(S, 2 * MD + (size_t)ri[r] * D + k);
            }
        }
        if (blockIdx.x == 0 && tid == 0) {            // cost entry: the right-hand side does not depend on cost
            S.This[Ssz - 1] = 0.f;
            small_update(S, Ssz - 1);
        }
    } else {
        // (q,q) interaction: one support point per thread, all M columns in chunks; gq, gp written (the finish kernel adds the
        // (x,q) column sums and applies the update)
        const int i = ((int)blockIdx.x - nXB) * kSmallThreads + tid;
        const bool valid = i < M;
        typename OpQQ::Row row;
        OpQQ::load_row(P, valid ? i : M - 1, row);
        F2 acc[NAQ];
#pragma unroll
        for (int k = 0; k < NAQ; ++k) acc[k] = f2(0.f, 0.f);
        for (int j0 = 0; j0 < M; j0 += kMidQChunk) {
            const int n = (M - j0 < kMidQChunk) ? M - j0 : kMidQChunk;
            __syncthreads();
            stage_cols<OpQQ>(P, j0, n, M, sm);
            __syncthreads();
            if (valid) sweep_cols<OpQQ>(P, row, sm, n, acc);
        }
        if (valid) {
            float a[NAQ];
#pragma unroll
            for (int k = 0; k < NAQ; ++k) a[k] = f2_sum(acc[k]);
            OpQQ::finish(P, i, row, a, nullptr);
        }
    }
}

// grid (ceil(maxM / 32), frames): 32 support points x 4 thread groups per CTA; group g adds the x CTAs g, g + 4, ... in that
// order, the groups are added in group order
template <int D, bool WLD, bool ETA>
__global__ void __launch_bounds__(kSmallThreads) small_mid_finish_kernel(SmallStep S, int rows_per_cta) {
    using OpX = typename std::conditional<ETA, AdjXQE<D>, AdjXQ<D, WLD>>::type;
    constexpr int NC = OpX::NACC_COL, FR = 32, G = kSmallThreads / FR;
    __shared__ float xch[G * NC * FR];
    if (!small_select_frame(S)) return;
    const int tid = threadIdx.x, r = tid & (FR - 1), g = tid / FR, M = S.M, Nx = S.Nx;
    if ((int)blockIdx.x * FR >= M) return;
    const int j = (int)blockIdx.x * FR + r;
    const int nXB = (Nx + rows_per_cta - 1) / rows_per_cta;
    const int mpad = (M + kSymGroup - 1) / kSymGroup * kSymGroup;
    const bool valid = j < M;
    float cs[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) cs[k] = 0.f;
    if (valid) {
        const float* part = S.ws + j;
        int b = g;
        for (; b + 3 * G < nXB; b += 4 * G) {         // 4 x NC independent loads in flight, added in CTA order
            float v[4][NC];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < NC; ++k) v[u][k] = __ldcg(&part[((size_t)(b + u * G) * NC + k) * mpad]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < NC; ++k) cs[k] += v[u][k];
        }
        for (; b < nXB; b += G) {
#pragma unroll
            for (int k = 0; k < NC; ++k) cs[k] += __ldcg(&part[((size_t)b * NC + k) * mpad]);
        }
    }
    if (g > 0) {
#pragma unroll
        for (int k = 0; k < NC; ++k) xch[(g * NC + k) * FR + r] = cs[k];
    }
    __syncthreads();
    if (g == 0 && valid) {
        for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
            for (int k = 0; k < NC; ++k) cs[k] += xch[(g2 * NC + k) * FR + r];
        }
        const size_t MD = (size_t)M * D, Ssz = 2 * MD + (size_t)Nx * D + 1;
        RhsParams P = small_params<D>(S);
        P.gc = S.lam + (Ssz - 1);
        P.gq = S.This; P.gp = S.This + MD;
        OpX::finish_col(P, j, cs);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            small_update(S, (size_t)j * D + k);
            small_update(S, MD + (size_t)j * D + k);
        }
    }
}

// Row passes of the x-row CTAs: one block of 128 rows per CTA until the x-row CTAs of all frames exceed ~8 resident CTAs
// per SM, then proportionally more (at most 8).  DICP_SMALL_XPASS overrides (tuning sweeps only).
inline int small_xpass(long long frames, long long maxNx, int sms) {
    static const int forced = [] {
        const char* e = getenv("DICP_SMALL_XPASS");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 64) ? v : 0;
    }();
    if (forced) return forced;
    const long long ctas = frames * ((maxNx + kSmallThreads - 1) / kSmallThreads);
    const long long r = ctas / ((long long)sms * 8);
    // a power of two: the kernels sweep 4 rows of a thread together (then 2, then 1), and a leftover single-row pass costs
    // about as much as a four-row one (latency-bound), so 5 passes were 4 + 1 = two sweeps where 4 passes are one
    return r >= 8 ? 8 : r >= 4 ? 4 : r >= 2 ? 2 : 1;
}

// Mid-size supports: CTAs of 128 threads x `rows` data points per thread (rows = 4, 2 or 1; more rows per thread = fewer
// shared-memory loads / shuffles per pair), 4 resident per SM.  A short grid of long CTAs ends in a badly filled last wave (8 frames
// x 50k points at 4 rows: 1.46 waves, i.e. the time of 2), so the largest `rows` whose grid fills its last wave to >= 85 % is
// taken, else the best-filled one.  `extra`: the other CTAs of the launch (support-point rows).
inline int small_rows_by_waves(long long frames, long long maxNx, long long extra, int sms, int min_rows) {
    int best = min_rows;
    double best_eff = -1.0;
    for (int rows = 4; rows >= min_rows; rows >>= 1) {
        const long long n = frames * ((maxNx + (long long)kSmallThreads * rows - 1) / ((long long)kSmallThreads * rows)) + extra;
        const double w = (double)n / ((double)kSmallMinbBig * sms);
        const double eff = w / (double)(long long)(w + 0.999999);
        if (eff >= 0.85) return rows;
        if (eff > best_eff) { best_eff = eff; best = rows; }
    }
    return best;
}
// BIG forward instantiation: rows of a thread swept together = row passes of an x CTA
inline int small_xpass_big(long long frames, long long maxNx, long long maxM, int sms) {
    static const int forced = [] {
        const char* e = getenv("DICP_SMALL_XPASS");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 64) ? v : 0;
    }();
    if (forced) return forced;
    return small_rows_by_waves(frames, maxNx, frames * ((maxM + kSmallThreads - 1) / kSmallThreads), sms, 1);
}

// dynamic shared memory of the two kernels: the staged column records (pair-interleaved, so an even number of columns)
inline size_t small_fwd_smem_bytes(long long M, int D) { return (size_t)((M + 1) / 2 * 2) * 2 * D * 4; }
inline size_t small_adj_smem_bytes(long long M, int D) {
    size_t a = (size_t)((M + 1) / 2 * 2) * 4 * D * 4;            // (q,q) records: 4D floats
    const size_t b = (size_t)kSmallChunk * 2 * D * 4;            // data-point chunk: 2D floats
    const size_t c = (size_t)kSmallThreads * 16 * 4;             // merge staging: 128 x NPART floats
    if (b > a) a = b;
    if (c > a) a = c;
    return a;
}

// workspace bytes: counters + max(forward block scalars, adjoint partials)
inline size_t small_workspace_bytes(long long M, long long Nx) {
    const long long nXB = (Nx + kSmallThreads - 1) / kSmallThreads, nQB = (M + kSmallThreads - 1) / kSmallThreads;
    const size_t fwd = (size_t)(nXB + nQB) * 4 * 4;
    size_t adj = (size_t)small_adj_nsplit((int)Nx) * 16 * (size_t)M * 4;
    const size_t ringb = (size_t)((Nx + kSmallThreads * 4 - 1) / (kSmallThreads * 4)) * 8 * 64 * 4;      // ring form: x CTAs (xpass = 1) x 8 x 64
    if (ringb > adj) adj = ringb;
    const size_t midb = (size_t)((Nx + kSmallThreads * 2 - 1) / (kSmallThreads * 2)) * 8 * (size_t)((M + kSymGroup - 1) / kSymGroup * kSymGroup) * 4;   // mid form
    if (M > kRingMaxQ && midb > adj) adj = midb;
    return kSmallCounters * 4 + (fwd > adj ? fwd : adj) + 256;
}

}  // namespace dicp
