// Symmetric pair engine for the (q,q) passes: every UNORDERED pair of support points is evaluated once.  Used for the
// ADJOINT pass (6 travelling accumulators, 41-83 operations per pair: 1.3-1.5x faster than the general engine on B200);
// the forward pass (up to 11 travelling accumulators, 16-30 operations per pair) is SHFL / shared-memory bound in this
// form and measured slower, so it stays on the general engine (dicp_sym_mode(2) routes it here for experiments).
//
// In the (q,q) passes rows and columns are the same point set and the pair term of (row n, column m) follows from the one
// of (row m, column n) by sign flips of the odd quantities (Op::pair_sym, ops_rhs.cuh): the exponential, the dot products
// and the scalar coefficients -- most of the 41 / 56 / 83 FP32 operations of a pair -- are shared.  The general engine
// (pair_engine.cuh) cannot exploit this: it keeps ROW accumulators in registers and streams columns, so a column-side
// contribution would need a cross-thread reduction per column.  Here a warp works as a ring:
//
//   * lane l owns R = 2 rows (register-resident data and accumulators, as before);
//   * the columns come in groups of 64 = 32 column PAIRS (packed fp32: one 64-bit register carries the same field of two
//     columns), staged in shared memory; at step s (0..31) lane l evaluates its rows against column pair (l + s) mod 32
//     and adds the column-side terms to a set of accumulators that TRAVELS with the column pair: after each step the
//     column accumulators move one lane down the ring (SHFL), so after 32 steps every column pair has met all 64 rows of
//     the warp and its accumulators are back in lane l = pair index;
//   * the four warps of a CTA (256 rows) then add their column sums in warp order through shared memory and write them
//     to the column-partial array [row block][k][column]; row accumulators are written once per work item to
//     [item][k][row].  A finish kernel adds, for row i, its items' row partials (item order) and the column partials of the
//     row blocks above it (block order), then runs Op::finish.  Fixed orders everywhere: deterministic, no atomics.
//
// Work items: for row block I (256 rows) one DIAGONAL item -- its own 256 columns, all ordered pairs, plain Op::pair --
// and the columns beyond it in chunks of Lc columns, symmetric.  Ragged tails (rows / columns >= M) are evaluated on
// zero-filled records with K multiplied by a 0/1 mask (a separate instantiation of the inner loop).
#pragma once
#include "ops_rhs.cuh"

namespace dicp {

static constexpr int kSymThreads = 128, kSymR = 2, kSymRows = (kSymThreads / 32) * 32 * kSymR;      // 256 rows per CTA
static constexpr int kSymGroup = 64;                                                                 // columns per ring round
static constexpr int kSymMaxBlocks = 256;                                                            // row blocks (M <= 65536)
static constexpr int kSymMinM = 4096;                // below this the general engine is as fast (measured)

// Record stride (floats) in shared memory such that the 8 lanes of an LDS.128 wavefront, reading consecutive records, hit 8
// different bank quads: the stride must be 4 mod 8 (REC is a multiple of 4).
template <int REC> struct SymStride { static constexpr int value = REC % 8 == 4 ? REC : REC + 4; };

struct SymPlan {
    int M, nrb, Lc, ngroups_total;      // points, row blocks, chunk length (columns), column groups (ceil(M/64))
    int items;
    int prefix[kSymMaxBlocks + 1];      // first item of every row block
};

inline SymPlan sym_make_plan(int M, int sms) {
    SymPlan p{};
    p.M = M;
    p.nrb = (M + kSymRows - 1) / kSymRows;
    p.ngroups_total = (M + kSymGroup - 1) / kSymGroup;
    const long long target = (long long)sms * 16;                   // symmetric items wanted: short items => short tail
    long long lc = ((long long)M * M / 2) / ((long long)kSymRows * target);
    lc = (lc + kSymGroup - 1) / kSymGroup * kSymGroup;
    if (lc < kSymGroup) lc = kSymGroup;
    p.Lc = (int)lc;
    const int mpad = p.ngroups_total * kSymGroup;
    int n = 0;
    for (int I = 0; I < p.nrb; ++I) {
        p.prefix[I] = n;
        const int beyond = mpad - (I + 1) * kSymRows;
        n += 1 + (beyond > 0 ? (beyond + p.Lc - 1) / p.Lc : 0);
    }
    p.prefix[p.nrb] = n;
    p.items = n;
    return p;
}
inline bool sym_applicable(long long M) { return M >= kSymMinM && M <= (long long)kSymMaxBlocks * kSymRows; }
inline size_t sym_rowpart_floats(const SymPlan& p, int nacc) { return (size_t)p.items * nacc * kSymRows; }
inline size_t sym_colpart_floats(const SymPlan& p, int nacc) {
    return (size_t)p.nrb * nacc * (size_t)p.ngroups_total * kSymGroup;
}
// upper bound of the extra workspace (beyond the packed columns) for any Op (NACC <= 16) at M points
inline size_t sym_workspace_bound(long long M, int sms) {
    if (!sym_applicable(M)) return 0;
    const SymPlan p = sym_make_plan((int)M, sms);
    return (sym_rowpart_floats(p, 16) + sym_colpart_floats(p, 16)) * 4 + (size_t)((M + 31) / 32) * 8 * 4 + 2048;
}

// ---- rectangular form: rows and columns are DIFFERENT point sets (data points x, support points q) -------------------------
// The adjoint of the (x,q) pass needs sums over the columns for every row (gx_k) AND sums over the rows for every column
// (gq_j, gp_j) of terms built from the same K, z', dot products: the general engine makes two passes over the same pairs
// (AdjXQx with rows = x, AdjXQq with rows = q).  Here ONE ring pass does both: rows = x (2 per lane), column pairs of q
// travelling round the warp with their accumulators (Op::pair_sym of AdjXQ).  Items = (256-row block) x (column chunk).
struct RectPlan {
    int Mrows, Ncols, rpc, nrb, ngroups_total, Lc, nchunks, items;       // rpc: rows per CTA (128 x rows per lane)
};
inline RectPlan rect_make_plan(int Mrows, int Ncols, int rows_per_lane, int sms) {
    RectPlan p{};
    p.Mrows = Mrows; p.Ncols = Ncols;
    p.rpc = kSymThreads * rows_per_lane;
    p.nrb = (Mrows + p.rpc - 1) / p.rpc;
    p.ngroups_total = (Ncols + kSymGroup - 1) / kSymGroup;
    long long want = ((long long)sms * 16 + p.nrb - 1) / p.nrb;            // column chunks per row block
    if (want < 1) want = 1;
    if (want > p.ngroups_total) want = p.ngroups_total;
    const int gpc = (int)((p.ngroups_total + want - 1) / want);              // groups per chunk
    p.Lc = gpc * kSymGroup;
    p.nchunks = (p.ngroups_total + gpc - 1) / gpc;
    p.items = p.nrb * p.nchunks;
    return p;
}
inline size_t rect_workspace_bytes(const RectPlan& p, int nf, int nacc_row, int nacc_col) {
    const size_t mpad = (size_t)p.ngroups_total * kSymGroup;
    const size_t npadcol = (mpad + 127) / 128 * 128;
    return align_up(npadcol * nf * 4, 256) + align_up((size_t)p.items * nacc_row * p.rpc * 4, 256) +
           align_up((size_t)p.nrb * nacc_col * mpad * 4, 256);
}

#if defined(__CUDACC__)

DICP_D F2 f2_shfl(F2 v, int src) {
    unsigned lo = (unsigned)(v.v & 0xffffffffull), hi = (unsigned)(v.v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    F2 r;
    r.v = ((unsigned long long)hi << 32) | lo;
    return r;
}

// one ring round: 32 steps over the 32 staged column pairs
template <class Op, bool MASKED, int R>
DICP_D void sym_ring_round(const typename Op::Params& prm, const typename Op::Row (&row)[R], const float* tile,
                           int stride, int lane, const float (&rmask)[R], int col0, int M, F2 (&acc)[R][Op::NACC],
                           F2 (&cacc)[Op::NACC_COL]) {
    constexpr int NF = Op::NF, PF4 = NF / 2;
#pragma unroll 2
    for (int s = 0; s < 32; ++s) {
        const int pr = (lane + s) & 31;
        const float4* rec = reinterpret_cast<const float4*>(tile + pr * stride);
        F2 c[NF];
#pragma unroll
        for (int k = 0; k < PF4; ++k) {
            const float4 v = rec[k];
            c[2 * k] = f2(v.x, v.y);
            c[2 * k + 1] = f2(v.z, v.w);
        }
        if (MASKED) {
            const int j = col0 + 2 * pr;
            const float m0 = j < M ? 1.f : 0.f, m1 = j + 1 < M ? 1.f : 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
                Op::template pair_sym<F2, true>(prm, row[r], c, acc[r], cacc, f2(m0 * rmask[r], m1 * rmask[r]));
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) Op::template pair_sym<F2, false>(prm, row[r], c, acc[r], cacc);
        }
        // the column accumulators follow their column pair: lane l next handles pair (l + s + 1) mod 32, held by lane l + 1
#pragma unroll
        for (int k = 0; k < Op::NACC_COL; ++k) cacc[k] = f2_shfl(cacc[k], (lane + 1) & 31);
    }
}

template <class Op>
__global__ void __launch_bounds__(kSymThreads) sym_pair_kernel(typename Op::Params prm, const float* __restrict__ colpack,
                                                               float* __restrict__ rowpart, float* __restrict__ colpart,
                                                               SymPlan plan) {
    constexpr int NF = Op::NF, NACC = Op::NACC, REC = 2 * NF, STRIDE = SymStride<REC>::value;   // conflict-free per-lane LDS.128
    __shared__ __align__(16) float tile[32 * STRIDE];
    __shared__ float xch[(kSymThreads / 32) * 32 * 2 * NACC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = plan.M;
    // item -> (row block I, chunk)
    int lo = 0, hi = plan.nrb;
    const int item = blockIdx.x;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (plan.prefix[mid] <= item) lo = mid; else hi = mid;
    }
    const int I = lo, chunk = item - plan.prefix[I];
    const int mpad = plan.ngroups_total * kSymGroup;

    // rows of this lane (zero-filled beyond M: masked wherever they could matter)
    typename Op::Row row[kSymR];
    float rmask[kSymR];
    int ri[kSymR];
#pragma unroll
    for (int r = 0; r < kSymR; ++r) {
        ri[r] = I * kSymRows + warp * (32 * kSymR) + r * 32 + lane;
        rmask[r] = ri[r] < M ? 1.f : 0.f;
        Op::load_row(prm, ri[r] < M ? ri[r] : M - 1, row[r]);
    }
    F2 acc[kSymR][NACC];
#pragma unroll
    for (int r = 0; r < kSymR; ++r)
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[r][k] = f2(0.f, 0.f);

    const bool rows_ragged = (I + 1) * kSymRows > M;
    int j0, j1;
    if (chunk == 0) { j0 = I * kSymRows; j1 = j0 + kSymRows; }
    else { j0 = (I + 1) * kSymRows + (chunk - 1) * plan.Lc; j1 = j0 + plan.Lc; }
    if (j1 > mpad) j1 = mpad;

    for (int g0 = j0; g0 < j1; g0 += kSymGroup) {
        __syncthreads();
        {   // stage 32 pair records (contiguous in the packed array) with the padded stride
            const float4* src = reinterpret_cast<const float4*>(colpack + (size_t)(g0 >> 1) * REC);
            constexpr int F4 = REC / 4;
            for (int t = tid; t < 32 * F4; t += kSymThreads) {
                const int rec = t / F4, f = t - rec * F4;
                reinterpret_cast<float4*>(tile + rec * STRIDE)[f] = src[t];
            }
        }
        __syncthreads();
        if (chunk == 0) {
            // diagonal block: every ORDERED pair, row side only (plain Op::pair, broadcast reads); columns >= M skipped
            const int nvalid = (M - g0 < kSymGroup) ? (M - g0 > 0 ? M - g0 : 0) : kSymGroup;
            const int npair = nvalid >> 1;
            for (int pr = 0; pr < npair; ++pr) {
                const float4* rec = reinterpret_cast<const float4*>(tile + pr * STRIDE);
                F2 c[NF];
#pragma unroll
                for (int k = 0; k < NF / 2; ++k) {
                    const float4 v = rec[k];
                    c[2 * k] = f2(v.x, v.y);
                    c[2 * k + 1] = f2(v.z, v.w);
                }
#pragma unroll
                for (int r = 0; r < kSymR; ++r) Op::template pair<F2>(prm, row[r], c, acc[r]);
            }
            if (nvalid & 1) {
                const float* rec = tile + npair * STRIDE;
                float c[NF], tmp[NACC];
#pragma unroll
                for (int k = 0; k < NF; ++k) c[k] = rec[2 * k];
#pragma unroll
                for (int r = 0; r < kSymR; ++r) {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) tmp[k] = 0.f;
                    Op::template pair<float>(prm, row[r], c, tmp);
#pragma unroll
                    for (int k = 0; k < NACC; ++k) acc[r][k] = vadd(acc[r][k], f2(tmp[k], 0.f));
                }
            }
            continue;
        }
        F2 cacc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) cacc[k] = f2(0.f, 0.f);
        if (rows_ragged || g0 + kSymGroup > M) sym_ring_round<Op, true, kSymR>(prm, row, tile, STRIDE, lane, rmask, g0, M, acc, cacc);
        else sym_ring_round<Op, false, kSymR>(prm, row, tile, STRIDE, lane, rmask, g0, M, acc, cacc);
        // lane l holds the sums of column pair l over this warp's 64 rows: add the four warps in warp order
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            float a, b;
            f2_unpack(cacc[k], a, b);
            xch[((warp * NACC + k) * 32 + lane) * 2] = a;
            xch[((warp * NACC + k) * 32 + lane) * 2 + 1] = b;
        }
        __syncthreads();
        for (int t = tid; t < NACC * kSymGroup; t += kSymThreads) {
            const int k = t / kSymGroup, cc = t - k * kSymGroup;            // column cc of the group = pair cc/2, half cc&1
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kSymThreads / 32; ++w) v += xch[((w * NACC + k) * 32 + (cc >> 1)) * 2 + (cc & 1)];
            colpart[((size_t)I * NACC + k) * mpad + g0 + cc] = v;
        }
    }
    // row partials of this item
    float* rp = rowpart + (size_t)item * NACC * kSymRows;
#pragma unroll
    for (int r = 0; r < kSymR; ++r)
#pragma unroll
        for (int k = 0; k < NACC; ++k) rp[k * kSymRows + warp * (32 * kSymR) + r * 32 + lane] = f2_sum(acc[r][k]);
}

// Finish: 32 rows per CTA, 4 thread groups per row.  Group g adds the row's item partials it, it+4, ... and the column
// partials of the row blocks J = g, g+4, ... (4 independent loads in flight per step), the groups are added in group order
// through shared memory, group 0 runs Op::finish.  Fixed order => deterministic.
template <class Op>
__global__ void __launch_bounds__(128) sym_finish_kernel(typename Op::Params prm, const float* __restrict__ rowpart,
                                                         const float* __restrict__ colpart, float* __restrict__ blockscal,
                                                         SymPlan plan) {
    constexpr int NACC = Op::NACC, NSCAL = Op::NSCAL, FR = 32, G = 4;
    __shared__ float red[32];
    __shared__ float xch[G * NACC * FR];
    const int r = threadIdx.x & (FR - 1), g = threadIdx.x / FR;
    const int i = blockIdx.x * FR + r;
    float rs[NSCAL > 0 ? NSCAL : 1];
#pragma unroll
    for (int k = 0; k < NSCAL; ++k) rs[k] = 0.f;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const bool valid = i < plan.M;
    if (valid) {
        const int I = i / kSymRows, rl = i - I * kSymRows;
        const int mpad = plan.ngroups_total * kSymGroup;
        const int it1 = plan.prefix[I + 1];
        for (int it = plan.prefix[I] + g; it < it1; it += 4 * G) {
            float v[4][NACC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = it + u * G;
                const float* rp = rowpart + (size_t)(t < it1 ? t : it) * NACC * kSymRows + rl;
#pragma unroll
                for (int k = 0; k < NACC; ++k) v[u][k] = rp[k * kSymRows];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (it + u * G < it1) {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) acc[k] += v[u][k];
                }
        }
        for (int J = g; J < I; J += 4 * G) {
            float v[4][NACC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = J + u * G;
                const float* cp = colpart + (size_t)(t < I ? t : J) * NACC * mpad + i;
#pragma unroll
                for (int k = 0; k < NACC; ++k) v[u][k] = cp[(size_t)k * mpad];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (J + u * G < I) {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) acc[k] += v[u][k];
                }
        }
    }
    if (g > 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) xch[(g * NACC + k) * FR + r] = acc[k];
    }
    __syncthreads();
    if (valid && g == 0) {
        for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[k] += xch[(g2 * NACC + k) * FR + r];
        }
        typename Op::Row row;
        Op::load_row(prm, i, row);
        Op::finish(prm, i, row, acc, rs);
    }
#pragma unroll
    for (int k = 0; k < NSCAL; ++k) {           // row scalars: block partials in a fixed tree, summed by scalar_reduce_kernel
        const float v = block_sum(rs[k], red);
        if (threadIdx.x == 0) blockscal[(size_t)blockIdx.x * NSCAL + k] = v;
    }
}

// packed columns (same layout as the general engine) + symmetric kernel + finish
template <class Op>
inline int run_pair_sym(const typename Op::Params& prm, int M, float* scal_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    static_assert(Op::PACKED, "symmetric engine: packed Ops");
    const SymPlan plan = sym_make_plan(M, device_info().sms);
    const int mpad = plan.ngroups_total * kSymGroup;
    const int npadcol = (mpad + 127) / 128 * 128;
    const size_t col_bytes = align_up((size_t)npadcol * Op::NF * 4, 256);
    const size_t row_bytes = align_up(sym_rowpart_floats(plan, Op::NACC) * 4, 256);
    const size_t cpart_bytes = align_up(sym_colpart_floats(plan, Op::NACC) * 4, 256);
    const int nfin = (M + 31) / 32;
    const size_t scal_bytes = align_up((size_t)nfin * (Op::NSCAL > 0 ? Op::NSCAL : 1) * 4, 256);
    if (ws == nullptr || col_bytes + row_bytes + cpart_bytes + scal_bytes > ws_bytes) return DICP_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 127) != 0) return DICP_EBADARG;
    float* colpack = (float*)ws;
    float* rowpart = (float*)((char*)ws + col_bytes);
    float* colpart = (float*)((char*)ws + col_bytes + row_bytes);
    float* blockscal = (float*)((char*)ws + col_bytes + row_bytes + cpart_bytes);
    pack_kernel_p<Op><<<(npadcol + 255) / 256, 256, 0, st>>>(prm, colpack, M, npadcol);
    sym_pair_kernel<Op><<<plan.items, kSymThreads, 0, st>>>(prm, colpack, rowpart, colpart, plan);
    sym_finish_kernel<Op><<<nfin, 128, 0, st>>>(prm, rowpart, colpart, blockscal, plan);
    launch_counter() += 3;
    if (Op::NSCAL > 0 && scal_out != nullptr) {
        scalar_reduce_kernel<<<1, 256, 0, st>>>(blockscal, nfin, Op::NSCAL, scal_out, 0);
        launch_counter() += 1;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DICP_OK : (int)e;
}

template <class Op, int R>
__global__ void __launch_bounds__(kSymThreads) rect_pair_kernel(typename Op::Params prm, const float* __restrict__ colpack,
                                                                float* __restrict__ rowpart, float* __restrict__ colpart,
                                                                RectPlan plan) {
    constexpr int NF = Op::NF, NACC = Op::NACC, NACC_COL = Op::NACC_COL, REC = 2 * NF, STRIDE = SymStride<REC>::value;
    __shared__ __align__(16) float tile[32 * STRIDE];
    __shared__ float xch[(kSymThreads / 32) * 32 * 2 * NACC_COL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int item = blockIdx.x, I = item / plan.nchunks, chunk = item - I * plan.nchunks;
    const int Mrows = plan.Mrows, Ncols = plan.Ncols, mpad = plan.ngroups_total * kSymGroup;
    typename Op::Row row[R];
    float rmask[R];
    const int RPC = plan.rpc;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int ri = I * RPC + warp * (32 * R) + r * 32 + lane;
        rmask[r] = ri < Mrows ? 1.f : 0.f;
        Op::load_row(prm, ri < Mrows ? ri : Mrows - 1, row[r]);
    }
    F2 acc[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[r][k] = f2(0.f, 0.f);
    const bool rows_ragged = (I + 1) * RPC > Mrows;
    const int j0 = chunk * plan.Lc;
    int j1 = j0 + plan.Lc;
    if (j1 > mpad) j1 = mpad;
    for (int g0 = j0; g0 < j1; g0 += kSymGroup) {
        __syncthreads();
        {
            const float4* src = reinterpret_cast<const float4*>(colpack + (size_t)(g0 >> 1) * REC);
            constexpr int F4 = REC / 4;
            for (int t = tid; t < 32 * F4; t += kSymThreads) {
                const int rec = t / F4, f = t - rec * F4;
                reinterpret_cast<float4*>(tile + rec * STRIDE)[f] = src[t];
            }
        }
        __syncthreads();
        F2 cacc[NACC_COL];
#pragma unroll
        for (int k = 0; k < NACC_COL; ++k) cacc[k] = f2(0.f, 0.f);
        if (rows_ragged || g0 + kSymGroup > Ncols) sym_ring_round<Op, true, R>(prm, row, tile, STRIDE, lane, rmask, g0, Ncols, acc, cacc);
        else sym_ring_round<Op, false, R>(prm, row, tile, STRIDE, lane, rmask, g0, Ncols, acc, cacc);
#pragma unroll
        for (int k = 0; k < NACC_COL; ++k) {
            float a, b;
            f2_unpack(cacc[k], a, b);
            xch[((warp * NACC_COL + k) * 32 + lane) * 2] = a;
            xch[((warp * NACC_COL + k) * 32 + lane) * 2 + 1] = b;
        }
        __syncthreads();
        for (int t = tid; t < NACC_COL * kSymGroup; t += kSymThreads) {
            const int k = t / kSymGroup, cc = t - k * kSymGroup;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < kSymThreads / 32; ++w) v += xch[((w * NACC_COL + k) * 32 + (cc >> 1)) * 2 + (cc & 1)];
            colpart[((size_t)I * NACC_COL + k) * mpad + g0 + cc] = v;
        }
    }
    float* rp = rowpart + (size_t)item * NACC * RPC;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < NACC; ++k) rp[k * RPC + warp * (32 * R) + r * 32 + lane] = f2_sum(acc[r][k]);
}

// rows: add the chunks' partials in chunk order, Op::finish
template <class Op>
__global__ void __launch_bounds__(128) rect_finish_rows_kernel(typename Op::Params prm, const float* __restrict__ rowpart,
                                                               RectPlan plan) {
    constexpr int NACC = Op::NACC;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= plan.Mrows) return;
    const int I = i / plan.rpc, rl = i - I * plan.rpc;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    for (int c = 0; c < plan.nchunks; ++c) {
        const float* rp = rowpart + (size_t)(I * plan.nchunks + c) * NACC * plan.rpc + rl;
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] += rp[k * plan.rpc];
    }
    typename Op::Row row;
    Op::load_row(prm, i, row);
    Op::finish(prm, i, row, acc, nullptr);
}

// columns: 32 columns x 4 thread groups per CTA; group g adds row blocks g, g+4, ... (4 loads in flight), groups in order
template <class Op>
__global__ void __launch_bounds__(128) rect_finish_cols_kernel(typename Op::Params prm, const float* __restrict__ colpart,
                                                               RectPlan plan) {
    constexpr int NC = Op::NACC_COL, FR = 32, G = 4;
    __shared__ float xch[G * NC * FR];
    const int r = threadIdx.x & (FR - 1), g = threadIdx.x / FR;
    const int j = blockIdx.x * FR + r;
    const int mpad = plan.ngroups_total * kSymGroup;
    const bool valid = j < plan.Ncols;
    float acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = 0.f;
    if (valid) {
        for (int J = g; J < plan.nrb; J += 4 * G) {
            float v[4][NC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = J + u * G;
                const float* cp = colpart + (size_t)(t < plan.nrb ? t : J) * NC * mpad + j;
#pragma unroll
                for (int k = 0; k < NC; ++k) v[u][k] = cp[(size_t)k * mpad];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (J + u * G < plan.nrb) {
#pragma unroll
                    for (int k = 0; k < NC; ++k) acc[k] += v[u][k];
                }
        }
    }
    if (g > 0) {
#pragma unroll
        for (int k = 0; k < NC; ++k) xch[(g * NC + k) * FR + r] = acc[k];
    }
    __syncthreads();
    if (valid && g == 0) {
        for (int g2 = 1; g2 < G; ++g2) {
#pragma unroll
            for (int k = 0; k < NC; ++k) acc[k] += xch[(g2 * NC + k) * FR + r];
        }
        Op::finish_col(prm, j, acc);
    }
}

// packed columns + rectangular ring kernel + the two finish kernels.  DICP_EWORKSPACE if `ws` is too small (the caller
// then falls back to the two-pass path).
template <class Op>
inline int run_pair_rect(const typename Op::Params& prm, int Mrows, int Ncols, void* ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int R = Op::RECT_R;
    const RectPlan plan = rect_make_plan(Mrows, Ncols, R, device_info().sms);
    if (ws == nullptr || rect_workspace_bytes(plan, Op::NF, Op::NACC, Op::NACC_COL) > ws_bytes) return DICP_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 127) != 0) return DICP_EBADARG;
    const int mpad = plan.ngroups_total * kSymGroup;
    const int npadcol = (mpad + 127) / 128 * 128;
    const size_t col_bytes = align_up((size_t)npadcol * Op::NF * 4, 256);
    const size_t row_bytes = align_up((size_t)plan.items * Op::NACC * plan.rpc * 4, 256);
    float* colpack = (float*)ws;
    float* rowpart = (float*)((char*)ws + col_bytes);
    float* colpart = (float*)((char*)ws + col_bytes + row_bytes);
    pack_kernel_p<Op><<<(npadcol + 255) / 256, 256, 0, st>>>(prm, colpack, Ncols, npadcol);
    rect_pair_kernel<Op, R><<<plan.items, kSymThreads, 0, st>>>(prm, colpack, rowpart, colpart, plan);
    rect_finish_rows_kernel<Op><<<(Mrows + 127) / 128, 128, 0, st>>>(prm, rowpart, plan);
    rect_finish_cols_kernel<Op><<<(Ncols + 31) / 32, 128, 0, st>>>(prm, colpart, plan);
    launch_counter() += 4;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DICP_OK : (int)e;
}

// Point sets beyond kSymMaxBlocks * 256 points: super-blocks of kSymSuper points.  The pairs inside a super-block go through
// the symmetric kernel (rows = columns = the super-block), the pairs between super-blocks A < B through the rectangular ring
// (rows from A, columns from B, Op::pair_sym: both sides), every launch ADDING to the outputs (zeroed first) in a fixed order.
static constexpr int kSymSuper = 32768;

// workspace of one super-block tile (symmetric or rectangular), for any adjoint Op (NF <= 12, NACC <= 6)
inline size_t sym_blocked_workspace_bytes(int sms) {
    const SymPlan sp = sym_make_plan(kSymSuper, sms);
    const size_t mpad = (size_t)sp.ngroups_total * kSymGroup, npadcol = (mpad + 127) / 128 * 128;
    const size_t symb = align_up(npadcol * 12 * 4, 256) + align_up(sym_rowpart_floats(sp, 6) * 4, 256) +
                        align_up(sym_colpart_floats(sp, 6) * 4, 256) + align_up((size_t)((kSymSuper + 31) / 32) * 4, 256);
    const RectPlan rp = rect_make_plan(kSymSuper, kSymSuper, 2, sms);
    const size_t rectb = rect_workspace_bytes(rp, 12, 6, 6);
    return (symb > rectb ? symb : rectb) + 4096;
}

template <class Op>
inline int run_pair_sym_blocked(const typename Op::Params& prm0, int M, void* ws, size_t ws_bytes, cudaStream_t st) {
    static_assert(Op::NSCAL == 0, "blocked symmetric evaluation: adjoint Ops (no row scalars)");
    constexpr int D = Op::NF / 4;                                   // (q', p, a, u) records
    typename Op::Params prm = prm0;
    if (!prm.accumulate) {
        cudaMemsetAsync(prm.gq, 0, (size_t)M * D * sizeof(float), st);
        cudaMemsetAsync(prm.gp, 0, (size_t)M * D * sizeof(float), st);
    }
    prm.accumulate = 1;
    const int nsb = (M + kSymSuper - 1) / kSymSuper;
    for (int A = 0; A < nsb; ++A) {
        const int a0 = A * kSymSuper, na = (M - a0 < kSymSuper) ? M - a0 : kSymSuper;
        typename Op::Params pa = prm;                               // row (and, for the diagonal tile, column) view of block A
        pa.q += (size_t)a0 * D; pa.p += (size_t)a0 * D; pa.a += (size_t)a0 * D; pa.u += (size_t)a0 * D;
        pa.gq += (size_t)a0 * D; pa.gp += (size_t)a0 * D;
        pa.col0 = 0;
        int rc = sym_applicable(na) ? run_pair_sym<Op>(pa, na, nullptr, ws, ws_bytes, st)
                                    : run_pair<Op>(pa, na, na, nullptr, 0, ws, ws_bytes, st);      // short last block
        if (rc != DICP_OK) return rc;
        for (int B = A + 1; B < nsb; ++B) {
            const int b0 = B * kSymSuper, nb = (M - b0 < kSymSuper) ? M - b0 : kSymSuper;
            pa.col0 = b0 - a0;
            rc = run_pair_rect<Op>(pa, na, nb, ws, ws_bytes, st);
            if (rc != DICP_OK) return rc;
        }
    }
    return DICP_OK;
}

#endif  // __CUDACC__

}  // namespace dicp
