// Common device helpers for the diffICP B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DICP_OK 0
#define DICP_EBADARG (-1)
#define DICP_EUNSUPPORTED (-2)
#define DICP_EWORKSPACE (-3)

// Padding coordinate for packed column tiles: far enough that K underflows to exactly 0 and
// every polynomial-in-z term times K is 0 (3 * (3e18)^2 = 2.7e37 < FLT_MAX).
#define DICP_FAR 3.0e18f

#if defined(__CUDACC__)
#define DICP_HD __host__ __device__ __forceinline__
#define DICP_D __device__ __forceinline__
#else
#define DICP_HD inline
#define DICP_D inline
#endif

namespace dicp {

// ---- math -----------------------------------------------------------------------------------
// exp2 of a non-positive argument given as its NEGATION: returns 2^(-t).  On the device this is a
// single MUFU.EX2 with the negate modifier folded in (checked in SASS: "MUFU.EX2 R, -R").
DICP_HD float ex2_neg(float t) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-t));
    return r;
#else
    return exp2f(-t);
#endif
}
DICP_HD float ex2f_fast(float t) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
#else
    return exp2f(t);
#endif
}
DICP_HD float lg2f_fast(float t) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
#else
    return log2f(t);
#endif
}

// ---- lane-generic arithmetic: V = float (one column) or F2 (two columns packed in one 64-bit register, evaluated with
// the Blackwell packed-fp32 instructions FFMA2 / FADD2 / FMUL2; a row-side scalar broadcasts through the .F32 operand
// form, which ptxas selects for f2(a,a)).  The Ops' per-pair formulas are written once against this interface.
DICP_HD float vfma(float a, float b, float c) { return fmaf(a, b, c); }
DICP_HD float vmul(float a, float b) { return a * b; }
DICP_HD float vadd(float a, float b) { return a + b; }
DICP_HD float vsub(float a, float b) { return a - b; }
DICP_HD float vex2n(float a) { return ex2_neg(a); }
template <class V> DICP_HD V vbc(float a);
template <> DICP_HD float vbc<float>(float a) { return a; }

// F2: two fp32 lanes in one 64-bit register.  Device: packed PTX instructions; host (tests): plain per-lane arithmetic.
struct F2 { unsigned long long v; };
DICP_HD F2 f2(float a, float b) {
    F2 r;
#if defined(__CUDA_ARCH__)
    asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(a), "f"(b));
#else
    float t[2] = {a, b};
    __builtin_memcpy(&r.v, t, 8);
#endif
    return r;
}
DICP_HD void f2_unpack(F2 x, float& a, float& b) {
#if defined(__CUDA_ARCH__)
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v));
#else
    float t[2];
    __builtin_memcpy(t, &x.v, 8);
    a = t[0]; b = t[1];
#endif
}
#if defined(__CUDA_ARCH__)
#define DICP_F2_OP3(name, ptx)                                                                                      \
    DICP_HD F2 name(F2 a, F2 b, F2 c) { F2 r; asm(ptx " %0,%1,%2,%3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#define DICP_F2_OP2(name, ptx, hostop)                                                                              \
    DICP_HD F2 name(F2 a, F2 b) { F2 r; asm(ptx " %0,%1,%2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
#else
#define DICP_F2_OP3(name, ptx)                                                                                      \
    DICP_HD F2 name(F2 a, F2 b, F2 c) { float a0, a1, b0, b1, c0, c1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1);    \
        f2_unpack(c, c0, c1); return f2(fmaf(a0, b0, c0), fmaf(a1, b1, c1)); }
#define DICP_F2_OP2(name, ptx, hostop)                                                                              \
    DICP_HD F2 name(F2 a, F2 b) { float a0, a1, b0, b1; f2_unpack(a, a0, a1); f2_unpack(b, b0, b1);                  \
        return f2(a0 hostop b0, a1 hostop b1); }
#endif
DICP_F2_OP3(vfma, "fma.rn.f32x2")
DICP_F2_OP2(vmul, "mul.rn.f32x2", *)
DICP_F2_OP2(vadd, "add.rn.f32x2", +)
DICP_F2_OP2(vsub, "sub.rn.f32x2", -)
DICP_HD F2 vex2n(F2 a) { float x, y; f2_unpack(a, x, y); return f2(ex2_neg(x), ex2_neg(y)); }
template <> DICP_HD F2 vbc<F2>(float a) { return f2(a, a); }
DICP_HD float f2_sum(F2 a) { float x, y; f2_unpack(a, x, y); return x + y; }
// lane helpers for Ops that keep a per-row scalar state (running maximum) next to packed sums
DICP_HD float vlane0(float a) { return a; }
DICP_HD float vlane0(F2 a) { float x, y; f2_unpack(a, x, y); return x; }
DICP_HD float vhmax(float a) { return a; }
DICP_HD float vhmax(F2 a) { float x, y; f2_unpack(a, x, y); return fmaxf(x, y); }

#if defined(__CUDACC__)
// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) ------------------------------
DICP_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DICP_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DICP_D void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
DICP_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DICP_D uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
DICP_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes multiple of 16, 16B aligned)
DICP_D void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- reductions -------------------------------------------------------------------------------
DICP_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DICP_D float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum, result valid in thread 0.  `red` is shared scratch of >= 32 floats.
// Deterministic: fixed shuffle tree, fixed warp order.
DICP_D float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    if (wid == 0) {
        t = lane < nw ? red[lane] : 0.f;
        t = warp_sum(t);
    }
    return t;
}
#endif  // __CUDACC__

}  // namespace dicp
