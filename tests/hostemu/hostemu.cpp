// CPU emulation of the pair-engine Ops: TEST INFRASTRUCTURE ONLY (built by tests/hostemu/build.sh with g++).
// Runs exactly the Op arithmetic and the dispatch logic of the device library (csrc/dispatch.cuh) row by row
// on the CPU, so that formulas, packing offsets and constants can be checked against the oracle in the
// build container, which has no GPU.  It is not linked into libdicp_b200.so and never used by the product.
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../diff_icp_b200/csrc/dispatch.cuh"

using namespace dicp;

extern "C" {

int emu_ksum(int D, unsigned mask, float sigma, const float* x, int64_t M, const float* y, int64_t N,
             const float* b, const float* c, const float* d,
             float* o_base, float* o_redscal, float* o_red, float* o_grad, float* o_dd, float* o_gend,
             float* o_hess, float* o_lap, float* o_gradlap, float* o_minsq, float* o_dot) {
    float* outs[11] = {o_base, o_redscal, o_red, o_grad, o_dd, o_gend, o_hess, o_lap, o_gradlap, o_minsq, o_dot};
    HostExec ex;
    return ksum_entry(ex, D, mask, sigma, x, M, y, N, b, c, d, outs);
}

int emu_rhs_forward(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                    const float* x, int64_t Nx, float* vq, float* dp, float* vx, float* scal) {
    HostExec ex;
    return rhs_forward_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, vq, dp, vx, scal);
}

int emu_rhs_forward_sym(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                        const float* x, int64_t Nx, float* vq, float* dp, float* vx, float* scal) {
    HostExec ex;
    ex.sym = true;
    return rhs_forward_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, vq, dp, vx, scal);
}

int emu_rhs_adjoint_sym(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                        const float* x, int64_t Nx, const float* a, const float* u, const float* wx, const float* gc,
                        float* gq, float* gp, float* gx) {
    HostExec ex;
    ex.sym = true;
    return rhs_adjoint_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, a, u, wx, gc, gq, gp, gx);
}

int emu_rhs_adjoint(int D, int withlogdet, float sigma, float eta, const float* q, const float* p, int64_t M,
                    const float* x, int64_t Nx, const float* a, const float* u, const float* wx, const float* gc,
                    float* gq, float* gp, float* gx) {
    HostExec ex;
    return rhs_adjoint_entry(ex, D, withlogdet, sigma, eta, q, p, M, x, Nx, a, u, wx, gc, gq, gp, gx);
}

}

extern "C" {
int emu_em_rowpass(int D, int lite, float sigma_old, const float* X, int64_t N, const float* mu_old, const float* wl2,
                   int64_t C, const float* mu_new, const float* lpi_new, float* T2, float* Y, float* rowP, float* rowQ,
                   float* sq, float* scal4) {
    HostExec ex;
    return em_rowpass_entry(ex, D, lite, sigma_old, X, N, mu_old, wl2, C, mu_new, lpi_new, T2, Y, rowP, rowQ, sq, scal4);
}
int emu_em_colstats(int D, float sigma_old, const float* X, int64_t N, const float* T2, const float* mu_old,
                    const float* wl2, int64_t C, float* stats) {
    HostExec ex;
    return em_colstats_entry(ex, D, sigma_old, X, N, T2, mu_old, wl2, C, stats);
}
}

