// Tiled N x M pair-reduction engine (sm_100a).
//
// One template serves every "for each row i: reduce over all columns j of f(row_i, col_j)" kernel on
// the diffICP hot path: the ten Gaussian kernel sums and their VJPs, the fused Hamiltonian right-hand
// side and its adjoint, and the GMM row pass.  An `Op` supplies
//
//   struct Params;                     raw device pointers + scalars (passed by value to the kernel)
//   THREADS, R, TILE, COLF4            CTA size, rows per thread, columns per tile, float4 per packed column
//   NACC, NSCAL                        accumulators per row, row-scalars that are summed over all rows
//   pack_col(prm, j, N, float* c)      build the packed record of column j (records j >= N are padding, never visited)
//   load_row(prm, i, Row&)             read row i from the raw arrays (pre-scaling coordinates)
//   init(acc), pair(prm, row, c, acc)  the per-pair arithmetic  (c = COLF4*4 floats of the packed column)
//   combine(a, b)                      how two partial accumulators merge across column splits (default: +)
//   finish(prm, i, row, acc, scal)     write the outputs of row i, fill its NSCAL row-scalars
//
// Data movement: columns are packed ONCE per launch into a 16-byte aligned, pre-scaled record array
// (pack_kernel), then streamed tile by tile into shared memory by the TMA engine with 1-D bulk copies
// (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP), NSTAGE-deep.  Every thread keeps R rows and their
// accumulators in registers and reads the staged column records with broadcast LDS.128.  The kernels are
// FFMA / MUFU.EX2 bound (D = 2,3 is not a dense contraction): no tensor cores on purpose.
//
// Parallelism: grid = (row blocks, column splits).  When there are too few rows to fill 148 SMs the
// column range is split; partial accumulators go to a workspace and `finish_kernel` merges them in a fixed
// order (deterministic, no float atomics).  Row-scalars (dcost, Hamiltonian pieces, ...) are block-reduced
// with a fixed tree and summed over blocks by `scalar_reduce_kernel`, also deterministic.
#pragma once
#include <atomic>
#include <type_traits>
#include <vector>
#include "common.cuh"

namespace dicp {

static constexpr int kStages = 3;

#ifndef DICP_WAVES
#define DICP_WAVES 4          // column splits are chosen so that the grid is about this many waves of resident CTAs
#endif
#ifndef DICP_COL_UNROLL
#define DICP_COL_UNROLL 2
#endif
#define DICP_STR2(x) #x
#define DICP_STR(x) DICP_STR2(x)

struct PairPlan {
    int M, N;
    int ntiles;      // column tiles of Op::TILE
    int nrb;         // row blocks
    int nsplit;      // column splits (grid.y)
    size_t col_bytes, part_bytes, scal_bytes;
    size_t total_bytes() const { return col_bytes + part_bytes + scal_bytes; }
};

#if defined(__CUDACC__)

template <class Op>
__global__ void pack_kernel(typename Op::Params prm, float4* __restrict__ colpack, int N, int Npad) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    float c[Op::COLF4 * 4];
    Op::pack_col(prm, j, N, c);
#pragma unroll
    for (int k = 0; k < Op::COLF4; ++k)
        colpack[(size_t)j * Op::COLF4 + k] = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
}

template <class Op>
__global__ void __launch_bounds__(Op::THREADS, Op::MINB)
pair_kernel(typename Op::Params prm, const float4* __restrict__ colpack, float* __restrict__ part,
            float* __restrict__ blockscal, int M, int N, int ntiles) {
    constexpr int R = Op::R, TILE = Op::TILE, CF4 = Op::COLF4, NACC = Op::NACC, NSCAL = Op::NSCAL;
    constexpr uint32_t STAGE_BYTES = TILE * CF4 * 16;
    __shared__ __align__(128) float4 stage[kStages][TILE * CF4];
    __shared__ __align__(8) uint64_t full[kStages];
    __shared__ float red[32];

    const int tid = threadIdx.x;
    const int nsplit = gridDim.y;
    const int t0 = (int)(((long long)blockIdx.y * ntiles) / nsplit);
    const int t1 = (int)(((long long)(blockIdx.y + 1) * ntiles) / nsplit);
    const int nt = t1 - t0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            if (s < nt) {
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_g2s(stage[s], colpack + (size_t)(t0 + s) * TILE * CF4, STAGE_BYTES, &full[s]);
            }
    }

    typename Op::Row row[R];
    float acc[R][NACC];
    const int rbase = blockIdx.x * (Op::THREADS * R) + tid;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = rbase + r * Op::THREADS;
        Op::load_row(prm, i < M ? i : M - 1, row[r]);
        Op::init(acc[r]);
    }

    for (int t = 0; t < nt; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const float4* sp = stage[s];
        const int left = N - (t0 + t) * TILE;            // the last tile may be partial: padded records are never visited
        const int ncol = left < TILE ? left : TILE;
_Pragma(DICP_STR(unroll DICP_COL_UNROLL))
        for (int j = 0; j < ncol; ++j) {
            float c[CF4 * 4];
#pragma unroll
            for (int k = 0; k < CF4; ++k) {
                float4 v = sp[j * CF4 + k];
                c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) Op::pair(prm, row[r], c, acc[r]);
        }
        __syncthreads();   // every thread is done with stage s -> it may be refilled
        if (tid == 0 && t + kStages < nt) {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(stage[s], colpack + (size_t)(t0 + t + kStages) * TILE * CF4, STAGE_BYTES, &full[s]);
        }
    }

    if (nsplit == 1) {
        float scal[NSCAL > 0 ? NSCAL : 1];
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) scal[k] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = rbase + r * Op::THREADS;
            if (i < M) {
                float rs[NSCAL > 0 ? NSCAL : 1];
                Op::finish(prm, i, row[r], acc[r], rs);
#pragma unroll
                for (int k = 0; k < NSCAL; ++k) scal[k] += rs[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) {
            float v = block_sum(scal[k], red);
            if (tid == 0) blockscal[(size_t)blockIdx.x * NSCAL + k] = v;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = rbase + r * Op::THREADS;
            if (i < M) {
                float* dst = part + (size_t)blockIdx.y * NACC * M + i;       // layout [split][k][row]: coalesced
#pragma unroll
                for (int k = 0; k < NACC; ++k) dst[(size_t)k * M] = acc[r][k];
            }
        }
    }
}

// ---- packed variant: two columns per lane-pair (FFMA2 / FADD2 / FMUL2) -----------------------------------------
// Column records are stored per PAIR of columns, interleaved: pair P holds NF float2 = (c_{2P}[k], c_{2P+1}[k]), so one
// 64-bit register carries the same field of two columns and a float4 LDS delivers two such fields.
// Ops that offer pair_rows<V, R>() (all R rows of a thread against one packed column record in one call)
template <class Op, class = void>
struct has_fused_rows : std::false_type {};
template <class Op>
struct has_fused_rows<Op, std::enable_if_t<Op::FUSED_ROWS>> : std::true_type {};

template <class Op>
__global__ void pack_kernel_p(typename Op::Params prm, float* __restrict__ colpack, int N, int Npad) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Npad) return;
    float c[Op::COLF4 * 4];
    Op::pack_col(prm, j, N, c);
    float* dst = colpack + (size_t)(j >> 1) * (2 * Op::NF) + (j & 1);
#pragma unroll
    for (int k = 0; k < Op::NF; ++k) dst[2 * k] = c[k];
}

template <class Op>
__global__ void __launch_bounds__(Op::THREADS, Op::MINB)
pair_kernel_p(typename Op::Params prm, const float4* __restrict__ colpack, float* __restrict__ part,
              float* __restrict__ blockscal, int M, int N, int ntiles) {
    constexpr int R = Op::R, TILE = Op::TILE, NF = Op::NF, NACC = Op::NACC, NSCAL = Op::NSCAL;
    static_assert(NF % 2 == 0 && TILE % 2 == 0, "packed records");
    constexpr int TP = TILE / 2;              // column pairs per tile
    constexpr int PF4 = NF / 2;               // float4 per column pair
    constexpr uint32_t STAGE_BYTES = TP * PF4 * 16;
    __shared__ __align__(128) float4 stage[kStages][TP * PF4];
    __shared__ __align__(8) uint64_t full[kStages];
    __shared__ float red[32];

    const int tid = threadIdx.x;
    const int nsplit = gridDim.y;
    const int t0 = (int)(((long long)blockIdx.y * ntiles) / nsplit);
    const int t1 = (int)(((long long)(blockIdx.y + 1) * ntiles) / nsplit);
    const int nt = t1 - t0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s)
            if (s < nt) {
                mbar_expect_tx(&full[s], STAGE_BYTES);
                bulk_g2s(stage[s], colpack + (size_t)(t0 + s) * TP * PF4, STAGE_BYTES, &full[s]);
            }
    }

    typename Op::Row row[R];
    F2 acc[R][NACC];
    const int rbase = blockIdx.x * (Op::THREADS * R) + tid;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = rbase + r * Op::THREADS;
        Op::load_row(prm, i < M ? i : M - 1, row[r]);
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[r][k] = f2(0.f, 0.f);
        Op::init_packed(acc[r]);
    }

    for (int t = 0; t < nt; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const float4* sp = stage[s];
        const int left = N - (t0 + t) * TILE;
        const int ncol = left < TILE ? left : TILE;
        // PAD_NULL Ops: the padded partner of an odd last column is a record that contributes exactly nothing
        const int npair = Op::PAD_NULL ? (ncol + 1) >> 1 : ncol >> 1;
_Pragma(DICP_STR(unroll DICP_COL_UNROLL))
        for (int P = 0; P < npair; ++P) {
            F2 c[NF];
#pragma unroll
            for (int k = 0; k < PF4; ++k) {
                float4 v = sp[P * PF4 + k];
                c[2 * k] = f2(v.x, v.y);
                c[2 * k + 1] = f2(v.z, v.w);
            }
            if constexpr (has_fused_rows<Op>::value) {
                Op::template pair_rows<F2, R>(prm, row, c, acc);     // R rows, one rarely taken branch (EM row pass)
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) Op::template pair<F2>(prm, row[r], c, acc[r]);
            }
        }
        if (!Op::PAD_NULL && (ncol & 1)) {    // odd trailing column of the whole problem: one-column form
            float c[NF];
#pragma unroll
            for (int k = 0; k < PF4; ++k) {
                float4 v = sp[npair * PF4 + k];
                c[2 * k] = v.x;
                c[2 * k + 1] = v.z;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float tmp[NACC];
#pragma unroll
                for (int k = 0; k < NACC; ++k) tmp[k] = 0.f;
                Op::template pair<float>(prm, row[r], c, tmp);
#pragma unroll
                for (int k = 0; k < NACC; ++k) acc[r][k] = vadd(acc[r][k], f2(tmp[k], 0.f));
            }
        }
        __syncthreads();
        if (tid == 0 && t + kStages < nt) {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            bulk_g2s(stage[s], colpack + (size_t)(t0 + t + kStages) * TP * PF4, STAGE_BYTES, &full[s]);
        }
    }

    float accf[R][NACC];
#pragma unroll
    for (int r = 0; r < R; ++r) Op::unpack_acc(acc[r], accf[r]);

    if (nsplit == 1) {
        float scal[NSCAL > 0 ? NSCAL : 1];
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) scal[k] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = rbase + r * Op::THREADS;
            if (i < M) {
                float rs[NSCAL > 0 ? NSCAL : 1];
                Op::finish(prm, i, row[r], accf[r], rs);
#pragma unroll
                for (int k = 0; k < NSCAL; ++k) scal[k] += rs[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NSCAL; ++k) {
            float v = block_sum(scal[k], red);
            if (tid == 0) blockscal[(size_t)blockIdx.x * NSCAL + k] = v;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = rbase + r * Op::THREADS;
            if (i < M) {
                float* dst = part + (size_t)blockIdx.y * NACC * M + i;       // layout [split][k][row]: coalesced
#pragma unroll
                for (int k = 0; k < NACC; ++k) dst[(size_t)k * M] = accf[r][k];
            }
        }
    }
}

// Merge the column-split partials of each row, then run the Op's finish.  A CTA of 128 threads handles 128 / G rows with
// G thread groups per row: group g merges splits g, g+G, g+2G, ... in that order, the G group results are merged in group
// order through shared memory (fixed order => deterministic).  G = 1 is the plain serial merge in split order; G = 16 is
// used when a few rows face many splits (column statistics of the EM step: 50 rows x thousands of splits).
template <class Op>
__global__ void finish_kernel(typename Op::Params prm, const float* __restrict__ part,
                              float* __restrict__ blockscal, int M, int nsplit, int G) {
    constexpr int NACC = Op::NACC, NSCAL = Op::NSCAL;
    __shared__ float red[32];
    __shared__ float xch[128 * NACC];
    const int FR = 128 / G;                            // rows per CTA
    const int r = threadIdx.x % FR, g = threadIdx.x / FR;
    const int i = blockIdx.x * FR + r;
    float rs[NSCAL > 0 ? NSCAL : 1];
#pragma unroll
    for (int k = 0; k < NSCAL; ++k) rs[k] = 0.f;
    float acc[NACC];
    const bool have = i < M && g < nsplit;
    if (have) {
        const float* src = part + (size_t)g * NACC * M + i;
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = src[(size_t)k * M];
        for (int s = g + G; s < nsplit; s += 4 * G) {          // 4 splits' loads in flight, combined in split order
            float b[4][NACC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ss = s + u * G;
                src = part + (size_t)(ss < nsplit ? ss : s) * NACC * M + i;
#pragma unroll
                for (int k = 0; k < NACC; ++k) b[u][k] = src[(size_t)k * M];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (s + u * G < nsplit) Op::combine(acc, b[u]);
        }
    }
    if (G > 1) {
        if (have && g > 0) {
#pragma unroll
            for (int k = 0; k < NACC; ++k) xch[(g * NACC + k) * FR + r] = acc[k];
        }
        __syncthreads();
        if (have && g == 0) {
            const int ng = nsplit < G ? nsplit : G;
            for (int g2 = 1; g2 < ng; ++g2) {
                float b[NACC];
#pragma unroll
                for (int k = 0; k < NACC; ++k) b[k] = xch[(g2 * NACC + k) * FR + r];
                Op::combine(acc, b);
            }
        }
    }
    if (have && g == 0) {
        typename Op::Row row;
        Op::load_row(prm, i, row);
        Op::finish(prm, i, row, acc, rs);
    }
#pragma unroll
    for (int k = 0; k < NSCAL; ++k) {
        float v = block_sum(rs[k], red);
        if (threadIdx.x == 0) blockscal[(size_t)blockIdx.x * NSCAL + k] = v;
    }
}

// thread groups per row of finish_kernel
DICP_HD int finish_groups(int M, int nsplit) { return (nsplit >= 8 && M <= 4096) ? 16 : 1; }

// out[k] (+)= scale * sum_b blockscal[b][k]   -- single CTA, fixed order.
__global__ void scalar_reduce_kernel(const float* __restrict__ blockscal, int nblocks, int nscal,
                                     float* __restrict__ out, int accumulate) {
    __shared__ float red[32];
    for (int k = 0; k < nscal; ++k) {
        float v = 0.f;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += blockscal[(size_t)b * nscal + k];
        v = block_sum(v, red);
        if (threadIdx.x == 0) out[k] = accumulate ? out[k] + v : v;
        __syncthreads();
    }
}

// ---- host side --------------------------------------------------------------------------------
// number of kernel launches issued by this library (a claim the bench reports as gpu_launches)
inline std::atomic<unsigned long long>& launch_counter() {          // atomic: frames may be registered from several host threads
    static std::atomic<unsigned long long> n{0};
    return n;
}
struct DeviceInfo {
    int sms = 0;
};
inline const DeviceInfo& device_info() {
    static DeviceInfo info = [] {
        DeviceInfo d;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        if (d.sms <= 0) d.sms = 148;
        return d;
    }();
    return info;
}

template <class Op>
inline int op_occupancy() {
    static int occ = [] {
        int o = 0;
        if constexpr (Op::PACKED) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pair_kernel_p<Op>, Op::THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, pair_kernel<Op>, Op::THREADS, 0);
        return o > 0 ? o : 1;
    }();
    return occ;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Upper bound of the workspace any Op needs for an (M rows, N columns) launch on this device.
// The Ops are held to these limits by static_asserts in make_plan.
static constexpr int kMaxSplit = 1024;
static constexpr int kMaxColF4 = 4, kMaxAcc = 16, kMaxScal = 8, kMaxRowsPerCta = 512, kMaxOcc = 16;
inline size_t pair_workspace_bound(long long M, long long N) {
    const long long sms = device_info().sms;
    if (M < 1) M = 1;
    if (N < 1) N = 1;
    size_t col = align_up((size_t)(N + 256) * kMaxColF4 * 16, 256);
    long long rows_a = (long long)DICP_WAVES * sms * kMaxOcc * kMaxRowsPerCta + M;      // nsplit <= waves*slots/nrb
    long long rows_b = M * ((N + 127) / 128);                        // nsplit <= number of column tiles
    long long rows = rows_a < rows_b ? rows_a : rows_b;
    size_t part = align_up((size_t)rows * kMaxAcc * 4, 256);
    size_t scal = align_up((size_t)((M + 127) / 128 + 1 + 4096 / 8) * kMaxScal * 4, 256);    // finish CTAs: <= M/8 when M <= 4096
    return col + part + scal + 1024;
}

template <class Op>
inline PairPlan make_plan(int M, int N) {
    static_assert(Op::COLF4 <= kMaxColF4 && Op::NACC <= kMaxAcc && Op::NSCAL <= kMaxScal, "workspace bound");
    static_assert(Op::THREADS * Op::R <= kMaxRowsPerCta && Op::TILE >= 128, "workspace bound");
    PairPlan p{};
    p.M = M; p.N = N;
    p.ntiles = (N + Op::TILE - 1) / Op::TILE;
    if (p.ntiles < 1) p.ntiles = 1;
    const int rows_per_cta = Op::THREADS * Op::R;
    p.nrb = (M + rows_per_cta - 1) / rows_per_cta;
    int occ = op_occupancy<Op>();
    if (occ > kMaxOcc) occ = kMaxOcc;
    const long long slots = (long long)device_info().sms * occ;
    long long s = (DICP_WAVES * slots) / p.nrb;
    if (s < 1) s = 1;
    if (s > p.ntiles) s = p.ntiles;
    if (s > kMaxSplit) s = kMaxSplit;          // few rows x very many columns (EM column statistics): the serial part of the
                                               // split merge grows with s while the grid is already > 1 wave
    if ((long long)p.nrb >= slots) s = 1;
    p.nsplit = (int)s;
    p.col_bytes = align_up((size_t)p.ntiles * Op::TILE * Op::COLF4 * 16, 256);
    p.part_bytes = p.nsplit > 1 ? align_up((size_t)p.nsplit * M * Op::NACC * 4, 256) : 0;
    const int fr = 128 / finish_groups(M, p.nsplit);
    int nfin = p.nsplit > 1 ? (M + fr - 1) / fr : p.nrb;
    p.scal_bytes = align_up((size_t)(nfin > p.nrb ? nfin : p.nrb) * (Op::NSCAL > 0 ? Op::NSCAL : 1) * 4, 256);
    return p;
}

// Pack columns, run the pair kernel (+ finish / scalar reduction when needed).
// scal_out: device pointer receiving the NSCAL summed row-scalars (may be null if NSCAL == 0).
template <class Op>
inline int run_pair(const typename Op::Params& prm, int M, int N, float* scal_out, int scal_accumulate,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    if (M <= 0) return DICP_OK;
    PairPlan p = make_plan<Op>(M, N);
    if (ws == nullptr || p.total_bytes() > ws_bytes) return DICP_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 127) != 0) return DICP_EBADARG;
    char* base = (char*)ws;
    float4* colpack = (float4*)base;
    float* part = (float*)(base + p.col_bytes);
    float* blockscal = (float*)(base + p.col_bytes + p.part_bytes);
    const int Npad = p.ntiles * Op::TILE;
    dim3 grid(p.nrb, p.nsplit);
    if constexpr (Op::PACKED) {
        pack_kernel_p<Op><<<(Npad + 255) / 256, 256, 0, st>>>(prm, (float*)colpack, N, Npad);
        pair_kernel_p<Op><<<grid, Op::THREADS, 0, st>>>(prm, colpack, part, blockscal, M, N, p.ntiles);
    } else {
        pack_kernel<Op><<<(Npad + 255) / 256, 256, 0, st>>>(prm, colpack, N, Npad);
        pair_kernel<Op><<<grid, Op::THREADS, 0, st>>>(prm, colpack, part, blockscal, M, N, p.ntiles);
    }
    launch_counter() += 2;
    int nblk = p.nrb;
    if (p.nsplit > 1) {
        const int G = finish_groups(M, p.nsplit);
        nblk = (M + 128 / G - 1) / (128 / G);
        finish_kernel<Op><<<nblk, 128, 0, st>>>(prm, part, blockscal, M, p.nsplit, G);
        launch_counter() += 1;
    }
    if (Op::NSCAL > 0 && scal_out != nullptr) {
        scalar_reduce_kernel<<<1, 256, 0, st>>>(blockscal, nblk, Op::NSCAL, scal_out, scal_accumulate);
        launch_counter() += 1;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? DICP_OK : (int)e;
}

#endif  // __CUDACC__

// ---- host emulation (tests only): the same Op arithmetic, executed row by row on the CPU ----------
template <class Op>
inline void run_pair_host(const typename Op::Params& prm, int M, int N, float* scal_out) {
    double scal[Op::NSCAL > 0 ? Op::NSCAL : 1] = {0};
    for (int i = 0; i < M; ++i) {
        typename Op::Row row;
        Op::load_row(prm, i, row);
        float acc[Op::NACC];
        Op::init(acc);
        for (int j = 0; j < N; ++j) {
            float c[Op::COLF4 * 4];
            Op::pack_col(prm, j, N, c);
            if constexpr (Op::PACKED) Op::template pair<float>(prm, row, c, acc);
            else Op::pair(prm, row, c, acc);
        }
        float rs[Op::NSCAL > 0 ? Op::NSCAL : 1] = {0};
        Op::finish(prm, i, row, acc, rs);
        for (int k = 0; k < Op::NSCAL; ++k) scal[k] += rs[k];
    }
    if (scal_out)
        for (int k = 0; k < Op::NSCAL; ++k) scal_out[k] = (float)scal[k];
}

// Rectangular both-sides evaluation (tests only): every (row, column) pair once with Op::pair_sym; rows finished with
// Op::finish, columns with Op::finish_col -- what the device's rectangular ring engine computes, in a simple order.
template <class Op>
inline void run_rect_host(const typename Op::Params& prm, int Mrows, int Ncols) {
    std::vector<float> cacc((size_t)Ncols * Op::NACC_COL, 0.f);
    for (int i = 0; i < Mrows; ++i) {
        typename Op::Row row;
        Op::load_row(prm, i, row);
        float acc[Op::NACC] = {0};
        for (int j = 0; j < Ncols; ++j) {
            float c[Op::COLF4 * 4];
            Op::pack_col(prm, j, Ncols, c);
            Op::template pair_sym<float>(prm, row, c, acc, &cacc[(size_t)j * Op::NACC_COL]);
        }
        Op::finish(prm, i, row, acc, nullptr);
    }
    for (int j = 0; j < Ncols; ++j) Op::finish_col(prm, j, &cacc[(size_t)j * Op::NACC_COL]);
}

// Symmetric evaluation (tests only): every unordered pair {i, j}, i < j, visited ONCE with Op::pair_sym, the diagonal
// pairs with Op::pair -- what the device's symmetric engine (sym_engine.cuh) computes, in a simple order.
template <class Op>
inline void run_pair_host_sym(const typename Op::Params& prm, int M, float* scal_out = nullptr) {
    std::vector<float> acc((size_t)M * Op::NACC, 0.f);
    for (int i = 0; i < M; ++i) {
        typename Op::Row row;
        Op::load_row(prm, i, row);
        float c[Op::COLF4 * 4];
        Op::pack_col(prm, i, M, c);
        Op::template pair<float>(prm, row, c, &acc[(size_t)i * Op::NACC]);
        for (int j = i + 1; j < M; ++j) {
            Op::pack_col(prm, j, M, c);
            Op::template pair_sym<float>(prm, row, c, &acc[(size_t)i * Op::NACC], &acc[(size_t)j * Op::NACC]);
        }
    }
    double scal[Op::NSCAL > 0 ? Op::NSCAL : 1] = {0};
    for (int i = 0; i < M; ++i) {
        typename Op::Row row;
        Op::load_row(prm, i, row);
        float rs[Op::NSCAL > 0 ? Op::NSCAL : 1] = {0};
        Op::finish(prm, i, row, &acc[(size_t)i * Op::NACC], rs);
        for (int k = 0; k < Op::NSCAL; ++k) scal[k] += rs[k];
    }
    if (scal_out)
        for (int k = 0; k < Op::NSCAL; ++k) scal_out[k] = (float)scal[k];
}

}  // namespace dicp
