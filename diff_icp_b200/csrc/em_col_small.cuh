// Column statistics of the EM step for FEW components (C <= 64): one launch, all lanes busy.
//
// The M step (/root/reference/diffICP/core/GMM.py:432-458; torch twin :286-297) needs, per component c, the log-domain
// sums over ALL points n of gamma_nc, gamma_nc (x_n - mu_c), gamma_nc |x_n - mu_c|^2 (EmCol in ops_em.cuh).  Through the
// general pair engine this is "C rows x N columns": with C = 50 (atlas) or 8 only 50 / 8 of the 128 threads of a CTA hold a
// row, and the result needs three launches (pack, pair, split merge).  Here a CTA takes a contiguous range of points, its
// 128 threads form G = 128 / C groups of C lanes; the points are staged chunk by chunk in shared memory and every group
// sweeps its own share of the chunk for all C components (same EmCol::pair arithmetic, same online-max rescaling); the
// groups are merged in group order (EmCol::combine, the log-sum-exp merge), the CTA's partial goes to the workspace, and
// the CTAs' partials are merged in two levels by "last CTA to finish" tickets (groups of 16 CTAs, then the groups) and the
// statistics written.  Fixed orders everywhere: deterministic, no floating-point atomics.
#pragma once
#include "ops_em.cuh"
#include "small_step.cuh"     // last_cta

namespace dicp {

static constexpr int kEmColMaxC = 64;
static constexpr int kEmColChunk = 512;      // points staged per step

static constexpr int kEmColGroup = 16;     // CTAs per level-1 merge group
#ifndef DICP_EM_CTAS_PER_SM
#define DICP_EM_CTAS_PER_SM 8
#endif
static constexpr int kEmColCtasPerSm = DICP_EM_CTAS_PER_SM;   // CTAs per SM of the column-statistics kernels (latency-bound: more resident warps)
static constexpr int kEmRowR = 4;          // points per thread in the row passes
static constexpr int kEmRowRows = 128 * kEmRowR;
#ifndef DICP_EM_ROW_RFULL
#define DICP_EM_ROW_RFULL 2                // points per thread swept together in the FULL row pass (9-11 packed accumulators per point)
#endif
#ifndef DICP_EM_ROW_MINB
#define DICP_EM_ROW_MINB 5                 // resident CTAs asked of the compiler for the row-pass kernel (register cap)
// measured on B200 (full pass, 640k x 50 2-D / 1.07M x 20 3-D / 4M x 8 3-D): 4 points, no cap (148 registers) 48.9 / 50.9 / 77.6 us;
// 2 points 44.5 / 51.2 / 89.7; 4 points, 4 CTAs 47.0 / 46.4 / 76.8; 2 points, 5 CTAs (94 registers) 43.2 / 46.7 / 75.3; 1 point,
// 6 CTAs 46.5 / 46.8 / 83.8 -- the pass is bound by its instruction count (~50 per point and component pair), not by occupancy
#endif

// acc <- merge (in a fixed order) of the partials s in [s0, s1) of component c: thread group g takes s0+g, s0+g+G, ...
// (loads of a batch of 8 issued before they are combined), then the groups are merged in group order through `xch`.
// Valid result in the threads with g == 0.  All threads of the CTA must call it.
template <int D>
DICP_D void em_col_merge(const float* part, int s0, int s1, int g, int G, int c, int C, bool work, float* xch, float* acc) {
    using Op = EmCol<D>;
    constexpr int NACC = Op::NACC;
    Op::init(acc);
    if (work) {
        for (int s = s0 + g; s < s1; s += 8 * G) {
            float b[8][NACC];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int ss = s + u * G;
                if (ss < s1) {
#pragma unroll
                    for (int k = 0; k < NACC; ++k) b[u][k] = __ldcg(&part[((size_t)ss * NACC + k) * C + c]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (s + u * G < s1) Op::combine(acc, b[u]);
        }
    }
    __syncthreads();
    if (work && g > 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) xch[(g * NACC + k) * C + c] = acc[k];
    }
    __syncthreads();
    if (work && g == 0) {
        for (int g2 = 1; g2 < G; ++g2) {
            float b[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) b[k] = xch[(g2 * NACC + k) * C + c];
            Op::combine(acc, b);
        }
    }
}

// accp += EmCol pairs of component `row` with the staged point pairs [a, b): four pairs per block (EmCol::pair_cols: one branch
// per block, four independent chains), then the rest one by one.  Same values as one call of pair() per staged pair.
template <int D>
DICP_D void em_col_sweep(const EmParams& P, const typename EmCol<D>::Row& row, const float4* sp, int a, int b, F2* accp) {
    using Op = EmCol<D>;
    constexpr int NF = Op::NF, U = 4;
    int t = a;
    for (; t + U <= b; t += U) {
        F2 cc[U][NF];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int k = 0; k < NF / 2; ++k) {
                const float4 v = sp[(t + u) * (NF / 2) + k];
                cc[u][2 * k] = f2(v.x, v.y);
                cc[u][2 * k + 1] = f2(v.z, v.w);
            }
        }
        Op::template pair_cols<F2, U>(P, row, cc, accp);
    }
    for (; t < b; ++t) {
        F2 cc[NF];
#pragma unroll
        for (int k = 0; k < NF / 2; ++k) {
            const float4 v = sp[t * (NF / 2) + k];
            cc[2 * k] = f2(v.x, v.y);
            cc[2 * k + 1] = f2(v.z, v.w);
        }
        Op::template pair<F2>(P, row, cc, accp);
    }
}

// Tail shared by the column-statistics kernels: two-level merge of the CTAs' partials (a single serial merge of hundreds of
// partials by one CTA is a chain of dependent L2 round trips): the last CTA of every group of kEmColGroup consecutive CTAs
// merges that group's partials into a level-2 partial, and the last of those mergers merges the level-2 partials and writes
// the statistics.  `acc` holds the CTA's own partial in the threads with g == 0 on entry.
template <int D>
DICP_D void em_col_publish_and_merge(const EmParams& P, int C, int g, int G, int c, bool work, float* xch, float* acc,
                                     const typename EmCol<D>::Row& row, float* __restrict__ part,
                                     unsigned* __restrict__ counter) {
    using Op = EmCol<D>;
    constexpr int NACC = Op::NACC;
    const int tid = threadIdx.x, nsplit = gridDim.x;
    if (work && g == 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) part[((size_t)blockIdx.x * NACC + k) * C + c] = acc[k];
    }
    const int ngroups = (nsplit + kEmColGroup - 1) / kEmColGroup;
    const int grp = blockIdx.x / kEmColGroup;
    const int s0 = grp * kEmColGroup, s1 = (s0 + kEmColGroup < nsplit) ? s0 + kEmColGroup : nsplit;
    if (!last_cta(&counter[1 + grp], (unsigned)(s1 - s0))) return;
    float* part2 = part + (size_t)nsplit * NACC * C;
    em_col_merge<D>(part, s0, s1, g, G, c, C, work, xch, acc);
    if (work && g == 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) part2[((size_t)grp * NACC + k) * C + c] = acc[k];
    }
    if (tid == 0) counter[1 + grp] = 0u;
    if (!last_cta(&counter[0], (unsigned)ngroups)) return;
    em_col_merge<D>(part2, 0, ngroups, g, G, c, C, work, xch, acc);
    if (work && g == 0) Op::finish(P, c, row, acc, nullptr);
    if (tid == 0) counter[0] = 0u;
}

template <int D>
__global__ void __launch_bounds__(128) em_col_small_kernel(EmParams P, int N, int C, float* __restrict__ part,
                                                           unsigned* __restrict__ counter) {
    using Op = EmCol<D>;
    constexpr int NACC = Op::NACC, NF = Op::NF, REC = 2 * NF;       // staged pair records: (x' (D), T2 [, pad]) of two points
    __shared__ __align__(16) float pts[(kEmColChunk / 2) * REC];
    __shared__ float xch[128 * NACC];
    const int tid = threadIdx.x;
    const int G = 128 / C;
    const int g = tid / C, c = tid - g * C;
    const bool work = g < G;
    const int nsplit = gridDim.x;
    const int per = (N + nsplit - 1) / nsplit;
    const int n0 = blockIdx.x * per, n1 = (n0 + per < N) ? n0 + per : N;

    typename Op::Row row;
    if (work) Op::load_row(P, c, row);
    F2 accp[NACC];
    Op::init_packed(accp);
    for (int j0 = n0; j0 < n1; j0 += kEmColChunk) {
        const int n = (n1 - j0 < kEmColChunk) ? n1 - j0 : kEmColChunk;
        const int npad = (n + 1) & ~1;                              // an odd count gets a null partner (contributes exactly 0)
        __syncthreads();
        for (int t = tid; t < npad; t += 128) {
            float rec[Op::COLF4 * 4];
            Op::pack_col(P, t < n ? j0 + t : N, N, rec);            // index N => the Op's null record
            float* dst = pts + (t >> 1) * REC + (t & 1);
#pragma unroll
            for (int k = 0; k < NF; ++k) dst[2 * k] = rec[k];
        }
        __syncthreads();
        if (work) {
            // packed fp32: two points per call (EmCol::pair<F2>), shares of whole pairs per group
            const int npair = npad >> 1;
            const int a = (int)(((long long)npair * g) / G), b = (int)(((long long)npair * (g + 1)) / G);
            const float4* sp = reinterpret_cast<const float4*>(pts);
            em_col_sweep<D>(P, row, sp, a, b, accp);
        }
    }
    float acc[NACC];
    Op::unpack_acc(accp, acc);
    // groups -> one partial per component, in group order
    if (work && g > 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) xch[(g * NACC + k) * C + c] = acc[k];
    }
    __syncthreads();
    if (work && g == 0) {
        for (int g2 = 1; g2 < G; ++g2) {
            float b[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) b[k] = xch[(g2 * NACC + k) * C + c];
            Op::combine(acc, b);
        }
    }
    em_col_publish_and_merge<D>(P, C, g, G, c, work, xch, acc, row, part, counter);
}

// ---- first sweep of the EM step in ONE pass over the points (C <= 64) ----------------------------------------------------------
// Row log-sum-exp AND the column statistics from a single read of X: a CTA takes groups of 512 points; per group
//   phase A  every thread owns 4 points and sweeps the components (resident in shared memory, EmRow<D,true>::pair<F2>):
//            T2_n = log2 sum_c 2^t2_nc stays in registers;
//   staging  the thread writes its points' packed records (x'_n, T2_n) to shared memory -- what em_col_small_kernel packs
//            from global memory, here without the round trip of T2 through HBM;
//   phase B  the threads regroup as G = 128 / C groups of C lanes, lane c of group g sweeps the group's share of the staged
//            points for component c (EmCol::pair<F2>, online-max rescaling), accumulators live in registers across groups.
// Then the CTA partial and the two-level ticket merge exactly as in em_col_small_kernel.  Algorithmic HBM traffic of this
// sweep: 4 D bytes per point (X once); 8 + 14 FP32 operations and 2 MUFU.EX2 per (point, component).
template <int D>
__global__ void __launch_bounds__(128) em_lse_col_small_kernel(EmParams P, int N, int C, int passes,
                                                               float* __restrict__ part, unsigned* __restrict__ counter) {
    using OpR = EmRow<D, true>;
    using OpC = EmCol<D>;
    constexpr int NFR = OpR::NF, RECR = 2 * NFR, PF4R = NFR / 2;
    constexpr int NFC = OpC::NF, RECC = 2 * NFC, NACC = OpC::NACC;
    constexpr int R = kEmRowR, ROWS = kEmRowRows;
    static_assert(ROWS == kEmColChunk, "one staged chunk per row group");
    __shared__ __align__(16) float cols[(kEmColMaxC / 2) * RECR];
    __shared__ __align__(16) float pts[(ROWS / 2) * RECC];
    __shared__ float xch[128 * NACC];
    DICP_EM_STATE_PROLOGUE(P)
    const int tid = threadIdx.x;
    const int Cpad = (C + 1) & ~1;
    for (int j = tid; j < Cpad; j += 128) {
        float rec[OpR::COLF4 * 4];
        OpR::pack_col(P, j, C, rec);
        float* dst = cols + (j >> 1) * RECR + (j & 1);
#pragma unroll
        for (int k = 0; k < NFR; ++k) dst[2 * k] = rec[k];
    }
    const int G = 128 / C;
    const int g = tid / C, c = tid - g * C;
    const bool work = g < G;
    typename OpC::Row rowc;
    if (work) OpC::load_row(P, c, rowc);
    F2 accp[NACC];
    OpC::init_packed(accp);
    const float4* spc = reinterpret_cast<const float4*>(cols);
    const float4* spp = reinterpret_cast<const float4*>(pts);
    const int ncp = Cpad >> 1;
    __syncthreads();
    for (int ps = 0; ps < passes; ++ps) {
        const long long base = ((long long)blockIdx.x * passes + ps) * ROWS;
        if (base >= N) break;
        const int n = (N - base < ROWS) ? (int)(N - base) : ROWS;
        const int npad = (n + 1) & ~1;
        // phase A: row log-sum-exp of this thread's 4 points
        typename OpR::Row row[R];
        F2 acc[R][OpR::NACC];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int t = tid + r * 128;
            OpR::load_row(P, (int)(base + (t < n ? t : n - 1)), row[r]);
            OpR::init_packed(acc[r]);
        }
        for (int Pp = 0; Pp < ncp; ++Pp) {
            F2 cc[NFR];
#pragma unroll
            for (int k = 0; k < PF4R; ++k) {
                const float4 v = spc[Pp * PF4R + k];
                cc[2 * k] = f2(v.x, v.y);
                cc[2 * k + 1] = f2(v.z, v.w);
            }
            OpR::template pair_rows<F2, R>(P, row, cc, acc);
        }
        __syncthreads();                       // the previous group's phase B is done with `pts`
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int t = tid + r * 128;
            if (t < npad) {
                float a[OpR::NACC];
                OpR::unpack_acc(acc[r], a);
                float* dst = pts + (t >> 1) * RECC + (t & 1);
                if (t < n) {
#pragma unroll
                    for (int k = 0; k < D; ++k) dst[2 * k] = row[r].x[k];
                    dst[2 * D] = a[OpR::A_M] + lg2f_fast(a[OpR::A_S]);
                } else {                       // null partner of an odd last point (EmCol::pack_col's padding record)
#pragma unroll
                    for (int k = 0; k < D; ++k) dst[2 * k] = DICP_FAR;
                    dst[2 * D] = 1.0e30f;
                }
#pragma unroll
                for (int k = D + 1; k < NFC; ++k) dst[2 * k] = 0.f;
            }
        }
        __syncthreads();
        // phase B: column statistics of the staged points
        if (work) {
            const int npair = npad >> 1;
            const int a = (int)(((long long)npair * g) / G), b = (int)(((long long)npair * (g + 1)) / G);
            em_col_sweep<D>(P, rowc, spp, a, b, accp);
        }
    }
    float acc[NACC];
    OpC::unpack_acc(accp, acc);
    if (work && g > 0) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) xch[(g * NACC + k) * C + c] = acc[k];
    }
    __syncthreads();
    if (work && g == 0) {
        for (int g2 = 1; g2 < G; ++g2) {
            float b[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) b[k] = xch[(g2 * NACC + k) * C + c];
            OpC::combine(acc, b);
        }
    }
    em_col_publish_and_merge<D>(P, C, g, G, c, work, xch, acc, rowc, part, counter);
}
// CTAs of the fused sweep: up to kEmColCtasPerSm per SM, whole 512-point groups each
inline void em_lse_col_small_grid(long long N, int sms, int* blocks, int* passes) {
    const long long groups = (N + kEmRowRows - 1) / kEmRowRows;
    const long long cap = (long long)sms * kEmColCtasPerSm;
    long long ps = (groups + cap - 1) / cap;
    if (ps < 1) ps = 1;
    *passes = (int)ps;
    *blocks = (int)((groups + ps - 1) / ps);
}

// ---- row passes of the EM step for FEW components (C <= 64): one launch ------------------------------------------------------
// Through the general engine a row pass is pack + pair kernel (TMA pipeline, one 128-column tile of which <= 64 columns are
// real) + scalar reduction.  With the components resident in shared memory a thread simply sweeps them for its 4 rows (packed
// fp32, EmRow::pair<F2>): no pipeline, coalesced row loads / stores, and the four free-energy sums of the full pass are block
// partials added in block order by scalar_reduce_kernel.  HBM-bound for C <~ 9 (12D + 8 bytes per point and EM step).

template <int D, bool LITE>
__global__ void __launch_bounds__(128, DICP_EM_ROW_MINB) em_row_small_kernel(EmParams P, int N, int C, int passes,
                                                           float* __restrict__ blockscal) {
    using Op = EmRow<D, LITE>;
    constexpr int NF = Op::NF, NACC = Op::NACC, NSCAL = Op::NSCAL, PF4 = NF / 2, REC = 2 * NF;
    static_assert(NF % 2 == 0, "packed records");
    __shared__ __align__(16) float cols[(kEmColMaxC / 2) * REC];
    __shared__ float red[32];
    DICP_EM_STATE_PROLOGUE(P)
    const int tid = threadIdx.x;
    const int Cpad = (C + 1) & ~1;                                   // an odd count gets a null partner record
    for (int j = tid; j < Cpad; j += 128) {
        float c[Op::COLF4 * 4];
        Op::pack_col(P, j, C, c);
        float* dst = cols + (j >> 1) * REC + (j & 1);
#pragma unroll
        for (int k = 0; k < NF; ++k) dst[2 * k] = c[k];
    }
    __syncthreads();
    const float4* sp = reinterpret_cast<const float4*>(cols);
    const int npair = Cpad >> 1;
    float scal[NSCAL > 0 ? NSCAL : 1];
#pragma unroll
    for (int k = 0; k < NSCAL; ++k) scal[k] = 0.f;
    // `passes` groups of kEmRowRows rows per CTA: the staging of the components is paid once per CTA; a group is swept RS rows of
    // a thread at a time
    constexpr int RS = LITE ? kEmRowR : DICP_EM_ROW_RFULL;
    static_assert(kEmRowR % RS == 0, "row sweep");
    for (int ps = 0; ps < passes * (kEmRowR / RS); ++ps) {
        typename Op::Row row[RS];
        F2 acc[RS][NACC];
        const int base = blockIdx.x * passes * kEmRowRows + ps * (128 * RS) + tid;
        if (base - tid >= N) break;
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = base + r * 128;
            Op::load_row(P, i < N ? i : N - 1, row[r]);
            Op::init_packed(acc[r]);
        }
        for (int Pp = 0; Pp < npair; ++Pp) {
            F2 c[NF];
#pragma unroll
            for (int k = 0; k < PF4; ++k) {
                const float4 v = sp[Pp * PF4 + k];
                c[2 * k] = f2(v.x, v.y);
                c[2 * k + 1] = f2(v.z, v.w);
            }
            Op::template pair_rows<F2, RS>(P, row, c, acc);
        }
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int i = base + r * 128;
            if (i < N) {
                float a[NACC], rs[NSCAL > 0 ? NSCAL : 1];
                Op::unpack_acc(acc[r], a);
                Op::finish(P, i, row[r], a, rs);
#pragma unroll
                for (int k = 0; k < NSCAL; ++k) scal[k] += rs[k];
            }
        }
    }
    if constexpr (NSCAL > 0) {
        // block partials; summed in block order by scalar_reduce_kernel (a single ticket counter would serialise the atomics
        // of thousands of CTAs: measured ~10 ns each)
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) {
            const float v = block_sum(scal[k], red);
            if (tid == 0) blockscal[(size_t)blockIdx.x * NSCAL + k] = v;
        }
    }
}
inline size_t em_row_small_workspace(long long N) { return 256 + (size_t)((N + kEmRowRows - 1) / kEmRowRows) * 4 * 4; }

// M step on the column statistics (/root/reference/diffICP/core/GMM.py:286-297 torch twin, :442-456 KeOps formulation):
//   w'_c = (m_c + log2 S0_c) ln 2,  mu'_c = mu_c + B_c / S0_c,  log pi'_c = w'_c - LSE(w'),
//   nds2 = N D sigma'^2 = sum_c 2^m_c (A_c - |B_c|^2 / S0_c)   [sig_mode 1: distances to the NEW centroids]
//                       = sum_c 2^m_c A_c                        [sig_mode 2: distances to the OLD centroids],  0: not wanted.
// One CTA; reductions over c with the fixed block tree (deterministic).  out_scal = { nds2, LSE(w') }.
template <int D>
__global__ void __launch_bounds__(256) em_mstep_kernel(const float* __restrict__ stats, const float* __restrict__ mu_old,
                                                       const float* __restrict__ w_old, int C, int do_mu, int do_w,
                                                       int sig_mode, float* __restrict__ mu_new, float* __restrict__ w_new,
                                                       float* __restrict__ lpi_new, float* __restrict__ out_scal) {
    __shared__ float red[32];
    __shared__ float bc;
    const int tid = threadIdx.x;
    float wmax = -INFINITY, nd = 0.f;
    for (int c = tid; c < C; c += 256) {
        const float* st = stats + (size_t)c * (D + 3);
        const float m = st[0], S0 = st[1], A = st[2 + D];
        float b2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float B = st[2 + k];
            b2 = fmaf(B, B, b2);
            mu_new[(size_t)c * D + k] = do_mu ? mu_old[(size_t)c * D + k] + B / S0 : mu_old[(size_t)c * D + k];
        }
        const float w = do_w ? (m + log2f(S0)) * kLn2 : w_old[c];
        w_new[c] = w;
        wmax = fmaxf(wmax, w);
        if (sig_mode == 1) nd += exp2f(m) * (A - b2 / S0);
        else if (sig_mode == 2) nd += exp2f(m) * A;
    }
    // max over c
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((tid & 31) == 0) red[tid >> 5] = wmax;
    __syncthreads();
    if (tid == 0) {
        float v = red[0];
        for (int k = 1; k < 8; ++k) v = fmaxf(v, red[k]);
        bc = v;
    }
    __syncthreads();
    wmax = bc;
    __syncthreads();
    float se = 0.f;
    for (int c = tid; c < C; c += 256) se += expf(w_new[c] - wmax);
    se = block_sum(se, red);
    if (tid == 0) bc = wmax + logf(se);
    __syncthreads();
    const float lse = bc;
    for (int c = tid; c < C; c += 256) lpi_new[c] = w_new[c] - lse;
    __syncthreads();
    nd = block_sum(nd, red);
    if (tid == 0) { out_scal[0] = nd; out_scal[1] = lse; }
}

// ---- EM loop with its state on the device (one CUDA graph replay per GMM_opt, no host read between steps) --------------------
// Host loop being replaced: GaussianMixtureUnif.EM_optimization (/root/reference/diffICP/core/GMM.py:330-357) around EM_step
// (:402-496): per step the host reads sigma' and the free-energy sums, updates the model and tests |FE - FE_prev| < tol |FE_prev|.
// Here sigma, kappa, the normalisation constant, FE_prev and the stop flag live in `state` (doubles, EmState); a step is
//   em_lse_col_small_kernel -> em_mstep_state_kernel -> em_row_small_kernel<D,false> -> scalar_reduce_kernel -> em_state_finalize_kernel
// and every kernel returns at once when the stop flag is set, so max_iterations steps can be enqueued back to back: the model
// freezes at the step that met the criterion, exactly where the host loop returns.  As the body of a WHILE conditional node of
// a CUDA graph (dicp_em_loop_create) the last kernel of a step decides on the device whether another step runs, so exactly the
// steps the host loop would execute are executed, from ONE graph launch.  Scalars are computed in double like the
// host code (Python floats), the stop test in fp32 like the reference's 0-d fp32 tensors.
template <int D>
__global__ void __launch_bounds__(256) em_mstep_state_kernel(const float* __restrict__ stats, const float* __restrict__ mu_old,
                                                             const float* __restrict__ w_old, int C, int do_mu, int do_w,
                                                             int sig_mode, float* __restrict__ mu_new, float* __restrict__ w_new,
                                                             float* __restrict__ lpi_new, double* __restrict__ state) {
    __shared__ float red[32];
    __shared__ float bc;
    if (state[ES_DONE] != 0.0) return;
    const int tid = threadIdx.x;
    float wmax = -INFINITY, nd = 0.f;
    for (int c = tid; c < C; c += 256) {
        const float* st = stats + (size_t)c * (D + 3);
        const float m = st[0], S0 = st[1], A = st[2 + D];
        float b2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float B = st[2 + k];
            b2 = fmaf(B, B, b2);
            mu_new[(size_t)c * D + k] = do_mu ? mu_old[(size_t)c * D + k] + B / S0 : mu_old[(size_t)c * D + k];
        }
        const float w = do_w ? (m + log2f(S0)) * kLn2 : w_old[c];
        w_new[c] = w;
        wmax = fmaxf(wmax, w);
        if (sig_mode == 1) nd += exp2f(m) * (A - b2 / S0);
        else if (sig_mode == 2) nd += exp2f(m) * A;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((tid & 31) == 0) red[tid >> 5] = wmax;
    __syncthreads();
    if (tid == 0) {
        float v = red[0];
        for (int k = 1; k < 8; ++k) v = fmaxf(v, red[k]);
        bc = v;
    }
    __syncthreads();
    wmax = bc;
    __syncthreads();
    float se = 0.f;
    for (int c = tid; c < C; c += 256) se += expf(w_new[c] - wmax);
    se = block_sum(se, red);
    if (tid == 0) bc = wmax + logf(se);
    __syncthreads();
    const float lse = bc;
    for (int c = tid; c < C; c += 256) lpi_new[c] = w_new[c] - lse;
    __syncthreads();
    nd = block_sum(nd, red);
    if (tid == 0) {
        // sigma' = sqrt(max(N D sigma'^2, 0) / (D N))  (core/GMM.py:296 / :455), or the old sigma when it is not optimised
        double sg = state[ES_SIGMA];
        if (sig_mode != 0) {
            const double ndd = (double)nd;
            sg = sqrt((ndd > 0.0 ? ndd : 0.0) / ((double)D * state[ES_N]));
        }
        state[ES_SIGMA_NEW] = sg;
    }
}

template <int D>
__global__ void __launch_bounds__(256) em_state_finalize_kernel(const float* __restrict__ scal4, int C, int keops_sem,
                                                                const float* __restrict__ mu_new, const float* __restrict__ w_new,
                                                                const float* __restrict__ lpi_new, float* __restrict__ mu,
                                                                float* __restrict__ w, float* __restrict__ lpi,
                                                                float* __restrict__ wl2, double* __restrict__ state,
                                                                int use_cond, cudaGraphConditionalHandle cond) {
    if (state[ES_DONE] != 0.0) return;
    const int tid = threadIdx.x;
    const double sg = state[ES_SIGMA_NEW];
    const double lgn_new = (double)D * (log(sg) + 0.5 * log(2.0 * 3.141592653589793));
    const float lgnf = (float)lgn_new;
    for (int i = tid; i < C * D; i += 256) mu[i] = mu_new[i];
    for (int c = tid; c < C; c += 256) {
        w[c] = w_new[c];
        const float l = lpi_new[c];
        lpi[c] = l;
        wl2[c] = (l - lgnf) * 1.4426950408889634f;
    }
    __syncthreads();                               // every thread has read the old state
    if (tid == 0) {
        const double P = scal4[0], Q = scal4[1], SQ = scal4[2], N = state[ES_N];
        const double lgn = keops_sem ? lgn_new : state[ES_LGN];             // core/GMM.py:485-488 vs :312-317
        const double inv2s2 = 1.0 / (2.0 * sg * sg);
        const double Cfe = P * inv2s2 + Q + N * lgn;
        const double FE = Cfe + SQ * inv2s2;
        const float FEf = (float)FE, lastf = (float)state[ES_LAST_FE], tolf = (float)state[ES_TOL];
        const bool stop = state[ES_HAVE_LAST] != 0.0 && tolf >= 0.f && fabsf(FEf - lastf) < tolf * fabsf(lastf);
        state[ES_SIGMA] = sg;
        state[ES_KAPPA] = (double)(float)(sqrt(0.5 * 1.4426950408889634) / (double)(float)sg);   // gauss_const((float)sigma).kappa
        state[ES_LGN] = lgn_new;
        state[ES_CFE] = Cfe;
        state[ES_FE] = FE;
        state[ES_LAST_FE] = (double)FEf;
        state[ES_HAVE_LAST] = 1.0;
        const double steps = state[ES_STEPS] + 1.0;
        state[ES_STEPS] = steps;
        if (stop) state[ES_DONE] = 1.0;
        // body of a WHILE node of a CUDA graph (dicp_em_loop_*): run another step?
        if (use_cond) cudaGraphSetConditional(cond, (!stop && steps < state[ES_MAXIT]) ? 1u : 0u);
    }
}

// ---- multi-GPU EM step: the buffer of the ONE all-reduce, and the M step on the reduced buffer ---------------------------------
// (core/GMM.py's _EM_optimization_pipelined; the frames of an atlas are sharded over ranks)
// pack:  buf = [ S0, B, A of every component rescaled from the local exponent m_c to the rank-agreed one m_ref_c (C x (D+2)) |
//                flag = 1 if some m_c - m_ref_c > 100 (the rescaled sums could overflow) | extra (n_extra plain partial sums) ]
// One launch instead of ~10 element-wise ones; every rank's buf is then summed by NCCL.
template <int D>
__global__ void __launch_bounds__(256) em_reduce_pack_kernel(const float* __restrict__ stats, const float* __restrict__ m_ref,
                                                             int C, const float* __restrict__ extra, int n_extra,
                                                             float* __restrict__ buf) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    int over = 0;
    for (int c = threadIdx.x; c < C; c += 256) {
        const float* st = stats + (size_t)c * (D + 3);
        const float d = st[0] - m_ref[c];
        if (d > 100.f) over = 1;
        const float sc = exp2f(fminf(d, 120.f));
#pragma unroll
        for (int k = 0; k < D + 2; ++k) buf[(size_t)c * (D + 2) + k] = st[1 + k] * sc;
    }
    if (over) any = 1;                       // benign race: every writer stores 1
    __syncthreads();
    float* tail = buf + (size_t)C * (D + 2);
    if (threadIdx.x == 0) tail[0] = any ? 1.f : 0.f;
    for (int k = threadIdx.x; k < n_extra; k += 256) tail[1 + k] = extra[k];
}

// M step on the all-reduced buffer (same formulas as em_mstep_kernel with m = m_ref), plus what the host loop needs next:
//   m_next_c = round(m_ref_c + log2 max(S0_c, 1e-30))        the exponent every rank agrees on for the next step
//   host     = [ N D sigma'^2 | the n_extra reduced sums | flag sum | 1 if some merged S0_c < 1e-30 ]   (ONE small D2H read)
template <int D>
__global__ void __launch_bounds__(256) em_mstep_merged_kernel(const float* __restrict__ buf, const float* __restrict__ m_ref,
                                                              const float* __restrict__ mu_old, const float* __restrict__ w_old,
                                                              int C, int do_mu, int do_w, int sig_mode, int n_extra,
                                                              float* __restrict__ mu_new, float* __restrict__ w_new,
                                                              float* __restrict__ lpi_new, float* __restrict__ m_next,
                                                              float* __restrict__ host) {
    __shared__ float red[32];
    __shared__ float bc;
    __shared__ int vanished;
    const int tid = threadIdx.x;
    if (tid == 0) vanished = 0;
    __syncthreads();
    float wmax = -INFINITY, nd = 0.f;
    int van = 0;
    for (int c = tid; c < C; c += 256) {
        const float* st = buf + (size_t)c * (D + 2);
        const float m = m_ref[c], S0 = st[0], A = st[1 + D];
        if (S0 < 1e-30f) van = 1;
        float b2 = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float B = st[1 + k];
            b2 = fmaf(B, B, b2);
            mu_new[(size_t)c * D + k] = do_mu ? mu_old[(size_t)c * D + k] + B / S0 : mu_old[(size_t)c * D + k];
        }
        const float w = do_w ? (m + log2f(S0)) * kLn2 : w_old[c];
        w_new[c] = w;
        wmax = fmaxf(wmax, w);
        if (sig_mode == 1) nd += exp2f(m) * (A - b2 / S0);
        else if (sig_mode == 2) nd += exp2f(m) * A;
        m_next[c] = rintf(m + log2f(fmaxf(S0, 1e-30f)));
    }
    if (van) vanished = 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if ((tid & 31) == 0) red[tid >> 5] = wmax;
    __syncthreads();
    if (tid == 0) {
        float v = red[0];
        for (int k = 1; k < 8; ++k) v = fmaxf(v, red[k]);
        bc = v;
    }
    __syncthreads();
    wmax = bc;
    __syncthreads();
    float se = 0.f;
    for (int c = tid; c < C; c += 256) se += expf(w_new[c] - wmax);
    se = block_sum(se, red);
    if (tid == 0) bc = wmax + logf(se);
    __syncthreads();
    const float lse = bc;
    for (int c = tid; c < C; c += 256) lpi_new[c] = w_new[c] - lse;
    __syncthreads();
    nd = block_sum(nd, red);
    const float* tail = buf + (size_t)C * (D + 2);
    if (tid == 0) {
        host[0] = nd;
        host[1 + n_extra] = tail[0];
        host[2 + n_extra] = vanished ? 1.f : 0.f;
    }
    for (int k = tid; k < n_extra; k += 256) host[1 + k] = tail[1 + k];
}

// number of point ranges (CTAs): up to kEmColCtasPerSm per SM, at least one chunk of points each
inline int em_col_small_splits(long long N, int sms) {
    long long s = (N + kEmColChunk - 1) / kEmColChunk;
    const long long cap = (long long)sms * kEmColCtasPerSm;
    if (s > cap) s = cap;
    return s < 1 ? 1 : (int)s;
}
// workspace: counters (1 + groups words, zeroed before every launch) | level-1 partials | level-2 partials
inline size_t em_col_small_counter_bytes(long long N, int sms) {
    const size_t words = 1 + (size_t)(em_col_small_splits(N, sms) + kEmColGroup - 1) / kEmColGroup;
    return (words * 4 + 255) / 256 * 256;
}
inline size_t em_col_small_workspace(long long N, int C, int sms) {
    const size_t ns = (size_t)em_col_small_splits(N, sms);
    return em_col_small_counter_bytes(N, sms) + (ns + (ns + kEmColGroup - 1) / kEmColGroup) * 8 * (size_t)C * 4;
}

}  // namespace dicp
