// EXPERIMENT (not part of the product ABI): classic RhsQQ, D = 3, with packed f32x2 math over PAIRS OF COLUMNS.
// Built only with -DDICP_EXPERIMENT_F2; used to decide whether FFMA2/FADD2/FMUL2 relieve the dispatch stalls that cap the
// scalar kernels at ~75 % of the FFMA pipe.
#pragma once
#include "pair_engine.cuh"

namespace dicp {

struct F2 { unsigned long long v; };
DICP_D F2 f2(float a, float b) { F2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
DICP_D void f2_unpack(F2 x, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
DICP_D F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
DICP_D F2 mul2(F2 a, F2 b) { F2 r; asm("mul.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DICP_D F2 sub2(F2 a, F2 b) { F2 r; asm("sub.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
DICP_D F2 add2(F2 a, F2 b) { F2 r; asm("add.f32x2 %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// pack: column pair P -> 3 float4: (qx0,qx1,qy0,qy1) (qz0,qz1,px0,px1) (py0,py1,pz0,pz1)
__global__ void exp_pack_kernel(const float* q, const float* p, float kappa, float4* out, int N) {
    int P = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * P >= N) return;
    float v[2][6];
    for (int h = 0; h < 2; ++h) {
        int j = 2 * P + h;
        for (int k = 0; k < 3; ++k) {
            v[h][k] = j < N ? (q[(size_t)j * 3 + k] - q[k]) * kappa : DICP_FAR;
            v[h][3 + k] = j < N ? p[(size_t)j * 3 + k] : 0.f;
        }
    }
    out[(size_t)P * 3 + 0] = make_float4(v[0][0], v[1][0], v[0][1], v[1][1]);
    out[(size_t)P * 3 + 1] = make_float4(v[0][2], v[1][2], v[0][3], v[1][3]);
    out[(size_t)P * 3 + 2] = make_float4(v[0][4], v[1][4], v[0][5], v[1][5]);
}

template <int R>
__global__ void __launch_bounds__(128) exp_f2_kernel(const float* __restrict__ q, const float* __restrict__ p, float kappa,
                                                     float alpha, const float4* __restrict__ colpack, float* __restrict__ part,
                                                     int M, int npairs_total) {
    constexpr int TP = 64;                      // column pairs per tile (= 128 columns)
    constexpr uint32_t STAGE_BYTES = TP * 3 * 16;
    __shared__ __align__(128) float4 stage[kStages][TP * 3];
    __shared__ __align__(8) uint64_t full[kStages];
    const int tid = threadIdx.x, nsplit = gridDim.y;
    const int ntiles = npairs_total / TP;
    const int t0 = (int)(((long long)blockIdx.y * ntiles) / nsplit), t1 = (int)(((long long)(blockIdx.y + 1) * ntiles) / nsplit);
    const int nt = t1 - t0;
    if (tid == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); fence_barrier_init(); }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < kStages; ++s)
            if (s < nt) { mbar_expect_tx(&full[s], STAGE_BYTES); bulk_g2s(stage[s], colpack + (size_t)(t0 + s) * TP * 3, STAGE_BYTES, &full[s]); }
    F2 rq[R][3], rp[R][3], vq[R][3], T[R][3];
    const int rbase = blockIdx.x * (128 * R) + tid;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = rbase + r * 128; if (i >= M) i = M - 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float a = (q[(size_t)i * 3 + k] - q[k]) * kappa, b = p[(size_t)i * 3 + k];
            rq[r][k] = f2(a, a); rp[r][k] = f2(b, b); vq[r][k] = f2(0.f, 0.f); T[r][k] = f2(0.f, 0.f);
        }
    }
    for (int t = 0; t < nt; ++t) {
        const int s = t % kStages;
        mbar_wait(&full[s], (uint32_t)((t / kStages) & 1));
        const float4* sp = stage[s];
#pragma unroll 2
        for (int P = 0; P < TP; ++P) {
            float4 a = sp[P * 3], b = sp[P * 3 + 1], c = sp[P * 3 + 2];
            F2 Q[3] = {f2(a.x, a.y), f2(a.z, a.w), f2(b.x, b.y)};
            F2 Pm[3] = {f2(b.z, b.w), f2(c.x, c.y), f2(c.z, c.w)};
#pragma unroll
            for (int r = 0; r < R; ++r) {
                F2 z0 = sub2(rq[r][0], Q[0]), z1 = sub2(rq[r][1], Q[1]), z2 = sub2(rq[r][2], Q[2]);
                F2 r2 = fma2(z2, z2, fma2(z1, z1, mul2(z0, z0)));
                F2 w = fma2(rp[r][2], Pm[2], fma2(rp[r][1], Pm[1], mul2(rp[r][0], Pm[0])));
                float ra, rb; f2_unpack(r2, ra, rb);
                F2 K = f2(ex2_neg(ra), ex2_neg(rb));
                F2 Kw = mul2(K, w);
                vq[r][0] = fma2(K, Pm[0], vq[r][0]); vq[r][1] = fma2(K, Pm[1], vq[r][1]); vq[r][2] = fma2(K, Pm[2], vq[r][2]);
                T[r][0] = fma2(Kw, z0, T[r][0]); T[r][1] = fma2(Kw, z1, T[r][1]); T[r][2] = fma2(Kw, z2, T[r][2]);
            }
        }
        __syncthreads();
        if (tid == 0 && t + kStages < nt) { mbar_expect_tx(&full[s], STAGE_BYTES); bulk_g2s(stage[s], colpack + (size_t)(t0 + t + kStages) * TP * 3, STAGE_BYTES, &full[s]); }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = rbase + r * 128;
        if (i < M) {
            float* dst = part + ((size_t)blockIdx.y * M + i) * 6;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float a, b; f2_unpack(vq[r][k], a, b); dst[k] = a + b;
                f2_unpack(T[r][k], a, b); dst[3 + k] = alpha * (a + b);
            }
        }
    }
}

__global__ void exp_finish_kernel(const float* part, float* vq, float* dp, int M, int nsplit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float a[6] = {0, 0, 0, 0, 0, 0};
    for (int s = 0; s < nsplit; ++s) for (int k = 0; k < 6; ++k) a[k] += part[((size_t)s * M + i) * 6 + k];
    for (int k = 0; k < 3; ++k) { vq[(size_t)i * 3 + k] = a[k]; dp[(size_t)i * 3 + k] = a[3 + k]; }
}

}  // namespace dicp
