// Micro-benchmark: how close can the instruction MIX of the fused pair kernels get to the FP32 roof of a B200 SM?
// Each variant runs a long unrolled loop of packed-fp32 instructions (fma/add/mul.rn.f32x2 -> FFMA2/FADD2/FMUL2) with
// optional MUFU.EX2, broadcast LDS.128 and 32-bit operand forms mixed in at the ratios of pair_kernel_p<RhsQQ<3,1,1,2>>
// (per 16 pairs: 152 FFMA2, 56 FADD2, 32 FMUL2, 16 MUFU, 12 LDS.128, 24 MOV), at several occupancies.
// Prints FP32 lane-operations per SM per clock (roof: 128).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/experiments/pipe_mix scripts/experiments/pipe_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0,%1;" : "=f"(r) : "f"(x)); return r; }

// VARIANT bits: 1 = use add/mul mix, 2 = MUFU, 4 = LDS, 8 = broadcast (scalar) operands, 16 = scalar FFMA instead of packed,
// 32 = the adds / multiplies of the mix written as FFMA2 with a constant operand (a + b = fma(a, 1, b), a * b = fma(a, b, -0)):
// bit-identical results, different pipe cost
template <int V, int NCH>
__global__ void __launch_bounds__(64) mix_kernel(int iters, float* out, float seed, float onev, float nzerov) {
    __shared__ float4 tile[64];
    if (threadIdx.x < 64) tile[threadIdx.x] = make_float4(seed, seed * 0.5f, 1.f - seed, 0.25f);
    __syncthreads();
    u64 acc[NCH];
    float facc[2 * NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) { acc[k] = pk(seed + k, seed - k); facc[2 * k] = seed + k; facc[2 * k + 1] = seed - k; }
    float rowa = seed * 1.0001f, rowb = seed * 0.9999f;
    u64 colv = pk(seed * 0.5f, seed * 0.25f);
    float e0 = seed, e1 = seed * 0.5f;
    for (int it = 0; it < iters; ++it) {
        if (V & 4) {
            float4 t = tile[(it + 0) & 63];             // broadcast LDS.128 (same address for all lanes)
            colv = pk(t.x, t.y);
            rowb = t.z;
        }
        if (V & 16) {
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) facc[k] = fmaf(facc[k], rowa, rowb);
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) facc[k] = fmaf(facc[k], rowb, rowa);
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) facc[k] = fmaf(facc[k], rowa, e0);
        } else {
            const u64 ra = (V & 8) ? pk(rowa, rowa) : pk(rowa, rowb);
            const u64 rb = (V & 8) ? pk(rowb, rowb) : pk(rowb, rowa);
            const u64 one = pk(onev, onev), nzero = pk(nzerov, nzerov);      // run-time values: ptxas cannot fold them back
            // 3 rounds x NCH instructions = 30 packed instructions for NCH = 10: per round FFMA2 : FADD2 : FMUL2 ~ 19 : 7 : 4
#pragma unroll
            for (int k = 0; k < NCH; ++k) acc[k] = fma2(acc[k], ra, colv);
            if (V & 2) { e0 = ex2(-e0 * 0.5f - 1.f); }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if ((V & 1) && k < 7) acc[k] = (V & 32) ? fma2(acc[k], one, rb) : add2(acc[k], rb);
                else acc[k] = fma2(acc[k], rb, ra);
            }
            if (V & 2) { e1 = ex2(-e1 * 0.5f - 1.f); }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if ((V & 1) && k < 4) acc[k] = (V & 32) ? fma2(acc[k], ra, nzero) : mul2(acc[k], ra);
                else acc[k] = fma2(acc[k], ra, (V & 2) ? pk(e0, e1) : colv);
            }
        }
    }
    float s = e0 + e1;
#pragma unroll
    for (int k = 0; k < NCH; ++k) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(acc[k])); s += a + b + facc[2 * k] + facc[2 * k + 1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V, int NCH>
double run(int ctas_per_sm, int sms, float clock_ghz, float* out) {
    const int iters = 20000;
    dim3 grid(sms * ctas_per_sm), block(64);
    mix_kernel<V, NCH><<<grid, block>>>(100, out, 0.37f, 1.f, -0.f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<V, NCH><<<grid, block>>>(iters, out, 0.37f, 1.f, -0.f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double laneops = (double)grid.x * 64 * iters * 3.0 * NCH * 2.0;      // FP32 lane-operations (packed counts 2)
    return laneops / (best * 1e-3) / (sms * clock_ghz * 1e9);                  // per SM per clock
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float ghz = khz * 1e-6f;
    float* out; cudaMalloc(&out, (size_t)sms * 32 * 64 * sizeof(float));
    printf("device %s, %d SMs, %.3f GHz (attribute; lanes/SM/clk below assume this clock)\n", prop.name, sms, ghz);
    printf("%-44s %8s %8s %8s %8s\n", "variant (lane-ops/SM/clk, roof 128)", "4 CTA", "8 CTA", "12 CTA", "16 CTA");
#define ROW(name, V, NCH) { printf("%-44s", name); for (int c : {4, 8, 12, 16}) printf(" %8.1f", run<V, NCH>(c, sms, ghz, out)); printf("\n"); }
    ROW("scalar FFMA x30, 20 chains", 16, 10)
    ROW("FFMA2 x30, 10 chains", 0, 10)
    ROW("FFMA2 x30, 10 chains, broadcast operands", 8, 10)
    ROW("FFMA2/FADD2/FMUL2 19:7:4", 1, 10)
    ROW("  + broadcast operands", 9, 10)
    ROW("  + 2 MUFU.EX2 per 30", 3, 10)
    ROW("  + 2 MUFU + 1 LDS.128 per 30", 7, 10)
    ROW("  + 2 MUFU + 1 LDS.128 + broadcast", 15, 10)
    ROW("mix as all-FFMA2 (add=fma(a,1,b), mul=fma(a,b,-0))", 33, 10)
    ROW("  all-FFMA2 + broadcast", 41, 10)
    ROW("  all-FFMA2 + 2 MUFU", 35, 10)
    ROW("  all-FFMA2 + 2 MUFU + LDS + broadcast", 47, 10)
    ROW("  all-FFMA2 + 2 MUFU + LDS + broadcast, 16 ch", 47, 16)
    ROW("FFMA2 x18, 6 chains", 0, 6)
    ROW("FFMA2 mix+MUFU+LDS, 6 chains", 15, 6)
    ROW("FFMA2 x48, 16 chains", 0, 16)
    ROW("FFMA2 mix+MUFU+LDS+bc, 16 chains", 15, 16)
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
