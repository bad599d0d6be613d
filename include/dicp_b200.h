/* dicp_b200.h -- C ABI of the B200-native diffICP hot path (libdicp_b200.so).
 *
 * Drop-in boundary: these entry points are what the reference's two dispatch seams would bind
 *   - GenKernel.set_computversion   /root/reference/diffICP/tools/kernel.py:91-110   (ten kernel reductions)
 *   - GaussianMixtureUnif.set_computversion  /root/reference/diffICP/core/GMM.py:126-144  (EM_step)
 * plus the fused forms of their callers LDDMMModel.ODE / Shoot (core/LDDMM.py:176-227, 286-299) and the
 * integrator updates (tools/integrators.py:20-51).  See INTEGRATION.md for the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (row-major (n,D), D in {2,3}) unless stated;
 *   - the caller allocates every output and the workspace; the library never allocates, frees or keeps pointers;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation, no host
 *     reads of device results: every call is CUDA-graph capturable;
 *   - return value: 0 = ok, <0 = DICP_E* argument error, >0 = cudaError_t of the launch;
 *   - workspace: at least dicp_pair_workspace_bytes(rows, cols) bytes, 128-byte aligned, private to the
 *     call sequence on that stream;
 *   - results are deterministic (no floating-point atomics).
 */
#ifndef DICP_B200_H
#define DICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DICP_OK 0
#define DICP_EBADARG (-1)
#define DICP_EUNSUPPORTED (-2)
#define DICP_EWORKSPACE (-3)

/* output selectors of dicp_ksum (bit mask; any supported subset is computed in ONE sweep) */
#define DICP_K_BASE 1u      /* KBase      kernel.py:131/178   sum_j K                    -> (M)   */
#define DICP_K_REDSCAL 2u   /* KRedScal   kernel.py:135/182   sum_j K d_j                -> (M)   */
#define DICP_K_RED 4u       /* KRed       kernel.py:138/186   sum_j K b_j                -> (M,D) */
#define DICP_K_GRAD 8u      /* GradKRed   kernel.py:142/190   sum_j gradK                -> (M,D) */
#define DICP_K_DD 16u       /* DDKRed     kernel.py:151/198   sum_j d_dK b_j^d           -> (M,D) */
#define DICP_K_GEND 32u     /* GenDKRed   kernel.py:155/202   sum_j gradK (c_i.b_j)      -> (M,D) */
#define DICP_K_HESS 64u     /* HessKRed   kernel.py:160/284   sum_j HessK (c_i-b_j)      -> (M,D) */
#define DICP_K_LAP 128u     /* LapKRed    kernel.py:164/206   sum_j LapK                 -> (M)   */
#define DICP_K_GRADLAP 256u /* GradLapKRed kernel.py:168/289  sum_j grad LapK            -> (M,D) */
#define DICP_K_MINSQ 512u   /* check_coverage kernel.py:324   min_j |x_i-y_j|^2          -> (M)   */
#define DICP_K_DOT 1024u    /* GradKRed_rev kernel.py:147/194 (rows=y, cols=x, b=d)      -> (M)   */

int dicp_version(void);

/* number of SMs of the current device (148 on B200) */
int dicp_sm_count(void);

/* Engine selection.  The (q,q) passes of dicp_rhs_adjoint run on the symmetric engine (every unordered pair evaluated
 * once) for M >= 4096 (beyond 65536 points in super-blocks of 32768), the (x,q) adjoint on its rectangular form.  The
 * engine is a PER-CALL argument of dicp_rhs_forward / dicp_rhs_adjoint (re-entrant: nothing process-global is consulted
 * when it is >= 0):
 *   DICP_ENGINE_DEFAULT        the process-wide default below
 *   DICP_ENGINE_GENERAL        every ordered pair through the general tiled engine
 *   DICP_ENGINE_SYMMETRIC      symmetric / rectangular ring engines for the adjoint passes (the default's default)
 *   DICP_ENGINE_SYMMETRIC_ALL  also the (q,q) pass of dicp_rhs_forward (measured slower on B200, kept for experiments)
 * Results of the engines agree to fp32 rounding.
 * dicp_sym_mode(mode) sets the process-wide DEFAULT (atomic; initial value 1 or the DICP_SYM environment variable) used by
 * calls that pass DICP_ENGINE_DEFAULT and by the small-support stage kernels; mode < 0 only queries.  Returns the previous
 * default. */
#define DICP_ENGINE_DEFAULT (-1)
#define DICP_ENGINE_GENERAL 0
#define DICP_ENGINE_SYMMETRIC 1
#define DICP_ENGINE_SYMMETRIC_ALL 2
int dicp_sym_mode(int mode);

/* number of kernel launches this library has issued so far in this process (launches recorded into a CUDA graph
 * are counted once, at capture) */
unsigned long long dicp_launch_count(void);          /* atomic counter: safe with several host threads */

/* Upper bound of the workspace needed by any pair kernel with `rows` rows and `cols` columns. */
size_t dicp_pair_workspace_bytes(int64_t rows, int64_t cols);

/* Gaussian kernel reductions, K(z) = exp(-|z|^2 / (2 sigma^2)), z = x_i - y_j.
 * b: (N,D) column vectors, c: (M,D) row vectors, d: (N) column scalars (null when unused).
 * o_*: outputs, null unless selected in `mask`. */
int dicp_ksum(int D, unsigned mask, float sigma,
              const float* x, int64_t M, const float* y, int64_t N,
              const float* b, const float* c, const float* d,
              float* o_base, float* o_redscal, float* o_red, float* o_grad, float* o_dd, float* o_gend,
              float* o_hess, float* o_lap, float* o_gradlap, float* o_minsq, float* o_dot,
              void* workspace, size_t workspace_bytes, void* stream);

/* Fused right-hand side of the Hamiltonian ODE (LDDMMModel.ODE, core/LDDMM.py:176-227).
 *   withlogdet: 0 = no divergence cost, 1 = dcost = -sum div v at the data points (x if given, else q)
 *   eta: 0 (classic / hybrid) or 1/lambda (logdet model)
 *   x / Nx: optional external data points (null / 0 => data points are the support points)
 * outputs: vq (M,D) = dq/dt, dp (M,D) = dp/dt, vx (Nx,D) = dx/dt,
 *          scal[4] = { dcost, A = sum p.KRed(q,q,p), B = sum p.GradKRed(q,q), C = sum LapKRed(q,q) }
 *          (B is only produced when withlogdet && !x or eta != 0; C only when eta != 0; else 0)
 *          H(q,p) = A/2 - eta*B - eta^2*C/2  (core/LDDMM.py:142-159). */
int dicp_rhs_forward(int D, int withlogdet, float sigma, float eta,
                     const float* q, const float* p, int64_t M, const float* x, int64_t Nx,
                     float* vq, float* dp, float* vx, float* scal,
                     void* workspace, size_t workspace_bytes, void* stream, int engine);

/* Adjoint (vector-Jacobian product) of dicp_rhs_forward: given cotangents a (of vq), u (of dp), wx (of vx)
 * and the device scalar gc (of dcost; null => 0), writes gq, gp (M,D) and gx (Nx,D). */
int dicp_rhs_adjoint(int D, int withlogdet, float sigma, float eta,
                     const float* q, const float* p, int64_t M, const float* x, int64_t Nx,
                     const float* a, const float* u, const float* wx, const float* gc,
                     float* gq, float* gp, float* gx,
                     void* workspace, size_t workspace_bytes, void* stream, int engine);

/* ---- GMM EM step (GaussianMixtureUnif.EM_step, core/GMM.py:236-325 torch twin, :402-529 KeOps formulation) ----
 * Row pass over points x components.  Responsibilities come from the OLD parameters
 *      t_nc = w_c - LSE(w) - |x_n - mu_old_c|^2 / (2 sigma_old^2) - D (ln sigma_old + ln(2 pi)/2)
 * given as wl2[c] = (w_c - LSE(w) - D(ln sigma_old + ln(2 pi)/2)) * log2(e); targets and free-energy sums use the NEW
 * ones (mu_new, lpi_new = natural-log mixture weights).
 *   lite != 0 : only T2[n] = log2 sum_c exp(t_nc) is produced (needed before the column pass)
 *   lite == 0 : also Y (N,D) = sum_c gamma_nc mu_new_c, optional per-point rowP, rowQ, sq (null = skip), and
 *               scal4 = { P = sum_n (sum_c gamma |mu_new_c|^2 - |Y_n|^2),  Q = sum_nc gamma (ln gamma - lpi_new_c),
 *                         SQ = sum_n |x_n - Y_n|^2,  DS = sum_nc gamma |x_n - mu_old_c|^2 }
 *               so that  Cfe = P/(2 sigma'^2) + Q + N D (ln sigma'' + ln(2 pi)/2),  FE = Cfe + SQ/(2 sigma'^2)
 *               (core/GMM.py:312-317 / :485-488, :527) for whatever sigma', sigma'' the caller's variant prescribes. */
int dicp_em_rowpass(int D, int lite, float sigma_old, const float* X, int64_t N, const float* mu_old,
                    const float* wl2, int64_t C, const float* mu_new, const float* lpi_new,
                    float* T2, float* Y, float* rowP, float* rowQ, float* sq, float* scal4,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Column pass: log-domain sufficient statistics of every component (M step, core/GMM.py:286-297 / :442-456):
 * stats[c] = { m_c, S0_c, B_c (D), A_c } with, for l_nc = log2 gamma_nc = wl2_c - kappa^2|x_n-mu_c|^2 - T2_n,
 *   m_c ~ max_n l_nc (a reference exponent),  S0_c = sum_n 2^(l_nc-m_c),  B_c = sum_n 2^(l_nc-m_c) (x_n - mu_old_c),
 *   A_c = sum_n 2^(l_nc-m_c) |x_n - mu_old_c|^2.
 * Then  w'_c = (m_c + log2 S0_c) ln 2,  mu'_c = mu_old_c + B_c/S0_c,
 *       N D sigma'^2 = sum_c 2^m_c (A_c - |B_c|^2/S0_c)  [distances to the NEW mu, KeOps formulation :453-455]
 *                    = sum_c 2^m_c A_c                    [distances to the OLD mu, torch twin :263,296].
 * Statistics of disjoint point subsets (frames on different GPUs) merge by max on m_c and rescaled sums. */
int dicp_em_colstats(int D, float sigma_old, const float* X, int64_t N, const float* T2, const float* mu_old,
                     const float* wl2, int64_t C, float* stats, void* workspace, size_t workspace_bytes, void* stream);

/* One EM step with the loop state on the device, for CUDA-graph replay of a whole EM_optimization (core/GMM.py:330-357 around
 * EM_step :402-496 / :236-325) without a host read between steps.  C <= 64 components, no outlier term.
 *   state: 16 doubles -- [0] sigma, [1] kappa = sqrt(log2(e)/2)/sigma as rounded to fp32, [2] D(ln sigma + ln(2 pi)/2),
 *          [3] stop flag, [4] 1 if [5] holds a previous free energy, [5] previous FE (as rounded to fp32), [6] steps executed,
 *          [7] Cfe, [8] FE of the last executed step, [9] number of points N, [10] tol (< 0: never stop), [11] scratch,
 *          [12] step limit (dicp_em_loop_* only).  The caller initialises [0,1,2,9,10,12] and zeroes the rest.
 *   mu, w, lpi (= w - LSE(w)), wl2 (= (lpi - state[2]) log2(e)): the CURRENT parameters, updated in place by the step;
 *   mu_new, w_new, lpi_new, stats (C, D+3), T2 (N), scal4: scratch;  Y (N,D): targets of the last executed step.
 * The step does nothing once state[3] != 0 (set when |FE - FE_prev| < tol |FE_prev| in fp32, the reference's test), so
 * max_iterations steps may be enqueued back to back; sig_mode / keops_sem as in dicp_em_mstep / the two EM orderings.
 * dicp_em_loop_create builds a CUDA graph whose only node is a WHILE conditional node with one such step as its body: the last
 * kernel of the step sets the loop condition on the device (another step iff not stopped and state[6] < state[12]), so
 * dicp_em_loop_launch executes exactly the steps the reference's host loop would, from one graph launch; state[12] = the step
 * limit (>= 1) is set by the caller before every launch.  All pointers are baked into the graph.  Returns null on failure. */
size_t dicp_em_state_workspace_bytes(int64_t N, int64_t C);
void* dicp_em_loop_create(int D, const float* X, int64_t N, int64_t C, float* mu, float* w, float* lpi, float* wl2, float* mu_new,
                          float* w_new, float* lpi_new, float* stats, float* Y, float* T2, float* scal4, double* state, int do_mu,
                          int do_w, int sig_mode, int keops_sem, void* workspace, size_t workspace_bytes, void* stream);
int dicp_em_loop_launch(void* loop, void* stream);
void dicp_em_loop_destroy(void* loop);
int dicp_em_state_step(int D, const float* X, int64_t N, int64_t C, float* mu, float* w, float* lpi, float* wl2, float* mu_new,
                       float* w_new, float* lpi_new, float* stats, float* Y, float* T2, float* scal4, double* state, int do_mu,
                       int do_w, int sig_mode, int keops_sem, void* workspace, size_t workspace_bytes, void* stream);

/* Multi-GPU EM step (frames sharded over ranks): the buffer of the step's ONE all-reduce and the M step on it.
 * dicp_em_reduce_pack: buf = [ S0, B (D), A of every component, rescaled from the local exponent stats[c][0] to the exponent
 *   m_ref[c] every rank agrees on (C x (D+2)) | flag (1 if some local exponent exceeds m_ref by more than 100) | extra[0..n_extra) ]
 *   -- (C (D+2) + 1 + n_extra) floats, summed over ranks by the caller (NCCL).
 * dicp_em_mstep_merged: dicp_em_mstep on the reduced buf with m = m_ref, plus m_next[c] = round(m_ref[c] + log2 max(S0_c, 1e-30))
 *   and host = [ N D sigma'^2 | the n_extra reduced sums | flag sum | 1 if some merged S0_c < 1e-30 ] (2 + 1 + n_extra floats,
 *   the one device-to-host read of the step).  Replaces the partial sums of core/GMM.py:286-297 / :442-456 across frames. */
int dicp_em_reduce_pack(int D, const float* stats, const float* m_ref, int64_t C, const float* extra, int n_extra, float* buf,
                        void* stream);
int dicp_em_mstep_merged(int D, const float* buf, const float* m_ref, const float* mu_old, const float* w_old, int64_t C,
                         int do_mu, int do_w, int sig_mode, int n_extra, float* mu_new, float* w_new, float* lpi_new,
                         float* m_next, float* host, void* stream);

/* First sweep of the EM step -- row log-sum-exp AND column statistics -- as one call: same statistics as dicp_em_rowpass(lite)
 * followed by dicp_em_colstats (the six KeOps reductions of core/GMM.py:410-415, :443-455 collapse into this sweep and the
 * full row pass).  With C <= 64 components it is ONE launch and ONE read of X (the row LSE never leaves the chip; T2_scratch
 * may be null); with more components it runs the two sweeps of the general engine and T2_scratch (N floats) is required. */
int dicp_em_lse_colstats(int D, float sigma_old, const float* X, int64_t N, const float* mu_old, const float* wl2, int64_t C,
                         float* T2_scratch, float* stats, void* workspace, size_t workspace_bytes, void* stream);

/* M step on the column statistics, one launch (core/GMM.py:286-297 / :442-456):
 *   mu_new = do_mu ? mu_old + B/S0 : mu_old;   w_new = do_w ? (m + log2 S0) ln 2 : w_old;   lpi_new = w_new - LSE(w_new);
 *   out_scal = { N D sigma'^2 (sig_mode 1: sum_c 2^m_c (A_c - |B_c|^2/S0_c), 2: sum_c 2^m_c A_c, 0: 0), LSE(w_new) }.
 * With do_mu = do_w = sig_mode = 0 (frozen mixture) `stats` may be null. */
int dicp_em_mstep(int D, const float* stats, const float* mu_old, const float* w_old, int64_t C, int do_mu, int do_w,
                  int sig_mode, float* mu_new, float* w_new, float* lpi_new, float* out_scal, void* stream);

/* lgam (N,C) = log_softmax_c(w_c - |x_n-mu_c|^2/(2 sigma^2)) (core/GMM.py:221-232) and/or argmax (N) int64
 * (first index wins ties, core/GMM.py:677-680); either output may be null. */
int dicp_log_resp(int D, float sigma, const float* X, int64_t N, const float* mu, const float* w, int64_t C,
                  float* lgam, long long* argmax, void* stream);

/* ---- Fused integrator stages for small supports (M <= dicp_small_max_support(); any number of data points) ----
 * The grid / decimated support schemes of DiffPSR (core/PSR.py:430-493) give tens to hundreds of support points: one
 * right-hand-side evaluation is then microseconds of arithmetic and launch latency dominates.  These entry points do a
 * whole Euler / Ralston STAGE in one launch.  Flat vectors: state / cotangent [q (M,D) | p (M,D) | x (Nx,D) | cost],
 * F = [vq | dp | vx | dcost, A, B, C] (S+3 floats), G = [gq | gp | gx | 0] (S floats).
 * workspace: dicp_small_workspace_bytes(M, Nx) bytes, ZERO-initialised before its first use (it holds ticket counters
 * that the kernels reset themselves), private to one launch sequence.
 *
 * Kernel forms behind the two entry points (chosen by the library from M, Nx and the frame count; same arguments, results
 * equal to fp32 rounding): up to 64 support points one launch per stage (adjoint: ring form, every (x,q) pair once); above, a
 * forward stage compiled for 128 registers and an adjoint stage of two launches (ring rounds over 64-column groups packed from
 * the state vector + a finish launch that merges the column sums); without data points or with too few CTAs to fill the SMs,
 * the one-launch form that splits the data columns over CTAs.
 *
 * forward stage:  F = rhs(s_eval);  out = base + c_this*F + c_other*other   (other, out nullable)
 *   Euler step            s_eval = base = s_t, c_this = h
 *   Ralston stage 1 / 2   (tools/integrators.py:42-48)  c_this = 2h/3  /  base = s_t, other = F1, c_this = 3h/4, c_other = h/4 */
int dicp_small_max_support(void);
size_t dicp_small_workspace_bytes(int64_t M, int64_t Nx);
int dicp_small_rhs_step(int D, int withlogdet, float sigma, float eta, int64_t M, int64_t Nx, const float* s_eval,
                        const float* base, const float* other, float c_this, float c_other, float* out, float* F,
                        void* workspace, size_t workspace_bytes, void* stream);
/* adjoint stage:  G = J_rhs(s_eval)^T lam;  out = base + c_this*G + c_other*other + add   (other, add, out nullable;
 * out must not alias lam). */
int dicp_small_adj_step(int D, int withlogdet, float sigma, float eta, int64_t M, int64_t Nx, const float* s_eval,
                        const float* lam, const float* base, const float* other, const float* add, float c_this,
                        float c_other, float* out, float* G, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Batched (multi-frame) registration closure for small supports -------------------------------------------------
 * DiffPSR.Reg_opt (core/PSR.py:521-569) optimises K independent frames, each through LDDMMModel.Optimize
 * (core/LDDMM.py:338-398).  These entry points evaluate the closure of ALL frames with the same launches
 * (blockIdx.y = frame).  Frame k owns floats [k*fstride, (k+1)*fstride) of every state-like buffer (state, cotangent, F,
 * G: same flat layouts as above, with the frame's own sizes); dims = (K,2) int32 device array {M_k, Nx_k};
 * active = (K) int32 device array or null: frames with active[k] == 0 are skipped.  maxM / maxNx bound the frames' sizes
 * (maxM <= dicp_small_max_support()); fstride >= 2*maxM*D + maxNx*D + 4.
 * workspace of the two stage calls: K * ws_frame_bytes bytes, ws_frame_bytes >= dicp_small_workspace_bytes(maxM, maxNx)
 * and a multiple of 16, ZERO-initialised before its first use. */
int dicp_batch_rhs_step(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                        int64_t maxM, int64_t maxNx, int64_t fstride, const float* s_eval, const float* base,
                        const float* other, float c_this, float c_other, float* out, float* F, void* workspace,
                        size_t ws_frame_bytes, void* stream);
int dicp_batch_adj_step(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                        int64_t maxM, int64_t maxNx, int64_t fstride, const float* s_eval, const float* lam,
                        const float* base, const float* other, const float* add, float c_this, float c_other, float* out,
                        float* G, void* workspace, size_t ws_frame_bytes, void* stream);
/* p part of state0[k] <- X[k*xstride : k*xstride + M_k*D], cost entry <- 0 (the trial momenta of one L-BFGS round). */
int dicp_batch_set_p(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t fstride, const float* X,
                     int64_t xstride, float* state0, void* stream);
/* Quadratic data loss of every frame on its arrival state (x part if Nx_k > 0, else q part):
 * loss[k*lstride] = sum_n inv[k*ystride+n] |z_n - y[(k*ystride+n)*D ..]|^2, g_end data part = 2 inv (z - y).
 * workspace: dicp_batch_quad_workspace_bytes(K) bytes, zero before the first use. */
size_t dicp_batch_quad_workspace_bytes(int K);
int dicp_batch_quad_loss(int D, int K, const int* dims, const int* active, int64_t max_points, int64_t fstride,
                         const float* state_end, const float* y, const float* inv, int64_t ystride, float* g_end,
                         float* loss, int64_t lstride, void* workspace, size_t workspace_bytes, void* stream);
/* out[k*ostride ..] = { dcost(0), A, B, C, cost(1), (untouched: data loss), -, - | lam_p + lam_reg * vq(0) } with nscal
 * leading scalars (>= 6); lam == null: scalars only (forward-only evaluation). */
int dicp_batch_closure_out(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t fstride,
                           float lam_reg, const float* lam, const float* F0, const float* state_end, float* out,
                           int64_t ostride, int nscal, void* stream);
/* The WHOLE closure of every active frame -- what one closure() call of tools/optim.py:32-50 computes for one frame of
 * DiffPSR.Reg_opt (core/PSR.py:521-569 -> core/LDDMM.py:338-398: Shoot :286-299, trajloss :318-334, QuadLossFunctor
 * core/PSR.py:498-516, backward()): Euler shoot from (q0, p0 = X[k]), lambda*H(q0,p0) + cost(1), quadratic data loss
 * against (y, inv), adjoint sweep, d loss / d p0 -- in ONE launch: one thread-block cluster per frame, stages separated by
 * cluster barriers instead of kernel launches (csrc/cluster_closure.cuh; replaces 1 + nt + 1 + nt + 1 launches of the
 * dicp_batch_* stage kernels above; same values up to fp32 summation order).  eta = 0 models with data points, Euler.
 * traj: (nt+1, K, fstride) states, time-major with stride tstride; traj[0] must hold q0 and x0 of every frame, the data points
 * x(t) are written to traj[t] (support points and momenta of t > 0 are NOT written).  out as in dicp_batch_closure_out
 * (scalars 1 = A, 4 = cost(1), 5 = data loss).  The launch shape (64 or 128 threads per CTA, 8 or 16 CTAs per frame) is chosen
 * from the sizes and the frame count K.  dicp_batch_closure_cluster_rows returns the rows handled per CTA (> 0) when this form
 * applies to the given sizes (<= 64 support points, <= 16 x 2048 data points per frame), else 0; the launcher returns
 * DICP_EUNSUPPORTED in that case. */
int dicp_batch_closure_cluster_rows(int D, float eta, int scheme_euler, int64_t maxM, int64_t maxNx, int nt, int K);
int dicp_batch_closure_cluster(int D, int withlogdet, float sigma, float eta, int K, const int* dims, const int* active,
                               int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                               const float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride,
                               float lam_reg, float* out, int64_t ostride, int nscal, void* stream);
/* ---- Lock-step L-BFGS with the optimiser state on the device (csrc/lbfgs_device.cuh) -------------------------------------------
 * The per-frame state machines of dicp_lbfgs_* (torch.optim.LBFGS with strong-Wolfe line search as tools/optim.py:26,56 uses
 * it), one warp per frame, running in a kernel right after dicp_batch_closure_cluster: a lock-step round needs no host round
 * trip.  The caller owns all memory (device pointers below) and initialises it: ints zero except [0] = n_k, [1] = line search
 * on/off; dbl zero except [1] = 1 (H_diag), [18] = NaN (last closure value), [19] = +inf (best closure value); vec slot 0 = the
 * parameters (fp64).  Host-readable slots: ints [6] closure evaluations, [7] L-BFGS iterations, [16] has-best flag;
 * dbl [18], [19]; vec slot 0 (current parameters); best_x (fp32 parameters of the best closure value so far). */
#define DICP_LBFGS_DEV_NI 24
#define DICP_LBFGS_DEV_ND 24
#define DICP_LBFGS_DEV_NV 12
typedef struct dicp_lbfgs_dev {
    int K;                       /* frames */
    long long stride;            /* row stride of the per-frame vectors (>= max n_k) */
    int history, max_iter, max_eval, max_ls;
    double tol_grad, tol_change, lr, c1, c2;
    int* ints;                   /* (K, DICP_LBFGS_DEV_NI) */
    double* dbl;                 /* (K, DICP_LBFGS_DEV_ND) */
    double* vec;                 /* (K, DICP_LBFGS_DEV_NV, stride) */
    float* best_x;               /* (K, stride) */
    double *dirs, *stps;         /* (K, history, stride) curvature pairs (ring buffers) */
    double *ro, *al;             /* (K, history) */
    unsigned* counters;          /* 4 words: ticket, waiting frames of the round, rounds since begin, waiting frames at the end */
} dicp_lbfgs_dev;
/* optimizer.step() begins for the frames with mask[k] != 0 (null = all): writes their trial points into X (K, xstride) and the
 * flags active (K), the input buffers of dicp_batch_closure_cluster. */
int dicp_lbfgs_dev_begin(const dicp_lbfgs_dev* L, const unsigned char* mask, float* X, int64_t xstride, int* active, void* stream);
/* One lock-step round: dicp_batch_closure_cluster (same arguments) + the optimiser kernel, which consumes (loss, gradient) of
 * every waiting frame from `out` and writes the next trial points / active flags.  counters[3] = frames still waiting. */
int dicp_lbfgs_dev_round(const dicp_lbfgs_dev* L, int D, int withlogdet, float sigma, float eta, int K, const int* dims,
                         int* active, int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                         float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride, float lam_reg, float* out,
                         int64_t ostride, int nscal, void* stream);
/* The rounds of one optimizer.step() of ALL frames as ONE CUDA graph launch: a WHILE conditional node whose body is such a round
 * and whose condition (some frame still waits and fewer than max_rounds rounds ran) the optimiser kernel sets on the device.
 * All pointers are baked into the graph.  Returns null on failure. */
void* dicp_lbfgs_dev_loop_create(const dicp_lbfgs_dev* L, int D, int withlogdet, float sigma, float eta, int K, const int* dims,
                                 int* active, int64_t maxM, int64_t maxNx, int64_t fstride, int nt, float* traj, int64_t tstride,
                                 float* X, int64_t xstride, const float* y, const float* inv, int64_t ystride, float lam_reg,
                                 float* out, int64_t ostride, int nscal, int max_rounds);
int dicp_lbfgs_dev_loop_launch(void* loop, void* stream);
void dicp_lbfgs_dev_loop_destroy(void* loop);

/* counts[k*ntimes + t] += #{data points of frame k farther than `radius` from every support point at stored time t}
 * (GaussKernel.check_coverage over a whole trajectory, tools/kernel.py:324-329 as used in core/PSR.py:559-566);
 * traj: time-major, time point t of frame k at traj + t*tstride + k*fstride.  counts must be zeroed by the caller. */
int dicp_batch_coverage(int D, int K, const int* dims, const int* active, int64_t maxM, int64_t maxNx, int64_t fstride,
                        const float* traj, int64_t tstride, int ntimes, float radius, int* counts, void* stream);

/* ---- Lock-step L-BFGS over K independent problems (HOST arithmetic; pointers below are HOST pointers) ---------------
 * One optimiser per frame with the settings and the algorithm of torch.optim.LBFGS(line_search_fn="strong_wolfe") as
 * used by LBFGS_optimization (tools/optim.py:26): state persists over steps, strong-Wolfe line search with cubic
 * interpolation (c1 = 1e-4, c2 = 0.9, at most 25 trials), or a fixed unit step when line_search == 0 (:77).
 * Protocol:  begin_step(mask) ; loop { n = pending(X, active) ; if n == 0 break ; <evaluate the active frames at the
 * rows of X> ; feed(losses, grads) }.  X and grads are (K, stride) fp32 row-major, losses (K); only rows of frames that
 * were pending are read.  The best closure value seen by a frame and its parameters are tracked (tools/optim.py:42-44). */
void* dicp_lbfgs_create(int K, const int64_t* n, int64_t stride, int max_iter, int max_eval, int history,
                        double tolerance_grad, double tolerance_change);
void dicp_lbfgs_destroy(void* h);
int dicp_lbfgs_set_x(void* h, int k, const float* x);
int dicp_lbfgs_get_x(void* h, int k, float* x, int best);
int dicp_lbfgs_reset(void* h, int k, int line_search);
int dicp_lbfgs_begin_step(void* h, const uint8_t* mask);
int dicp_lbfgs_pending(void* h, float* X, uint8_t* active);
int dicp_lbfgs_feed(void* h, const float* losses, const float* grads);
/* out4 = { last closure value, best closure value, closure evaluations, L-BFGS iterations } of frame k */
int dicp_lbfgs_stats(void* h, int k, double* out4);
/* all frames at once: X (K, stride) <- current (best == 0) or best-so-far parameters; out (K, 4) <- the statistics above */
int dicp_lbfgs_get_all(void* h, float* X, int best);
int dicp_lbfgs_stats_all(void* h, double* out);

/* ---- Set-up helpers on point sets (not on the per-iteration path) ---------------------------
 * out[i] = second smallest squared distance from x_i to the points of x (the smallest is x_i itself): the Kmin(2)
 * reduction of intrinsic_scale (tools/point_sets.py:13-26); intrinsic scale = sqrt(mean(out)). */
int dicp_min2_sqdist(int D, const float* x, int64_t N, float* out, void* stream);
/* Greedy decimation with radius R (tools/point_sets.py:102-133): repeatedly keep the uncovered point with most uncovered
 * neighbours within R (smallest index on ties) and cover its neighbours.  One pick = one launch; nothing N x N is stored.
 * dicp_decimate_steps enqueues `nsteps` picks (restart != 0 first resets the state: all points uncovered); picks after
 * completion are no-ops.  kept (N int32, device) receives the kept indices in pick order.
 * dicp_decimate_status copies { number kept so far, done flag } to the HOST array nkept_done and synchronises the
 * stream (the only synchronising entry point of this library; the caller loops steps / status until done).
 * workspace: dicp_decimate_workspace_bytes(N) bytes, private to one decimation. */
size_t dicp_decimate_workspace_bytes(int64_t N);
int dicp_decimate_steps(int D, const float* x, int64_t N, float radius, int restart, int nsteps, int* kept, void* workspace,
                        size_t workspace_bytes, void* stream);
int dicp_decimate_status(const void* workspace, int64_t N, int* nkept_done, void* stream);

/* Quadratic data loss of the registration step (DiffPSR.QuadLossFunctor, core/PSR.py:498-516):
 *   loss[0] = sum_n inv[n] |x_n - y_n|^2,   g[n,:] = 2 inv[n] (x_n - y_n)      (inv[n] = 1 / (2 sigma_s(n)^2)). */
int dicp_quad_loss(int D, const float* x, const float* y, const float* inv, int64_t n, float* g, float* loss,
                   void* workspace, size_t workspace_bytes, void* stream);

/* out = a + alpha*f1 + beta*f2  (f2 may be null) over n floats: the Euler / Ralston state updates
 * (tools/integrators.py:27-29, 42-48) and the adjoint accumulations. */
int dicp_axpy(int64_t n, float* out, const float* a, float alpha, const float* f1, float beta, const float* f2,
              void* stream);

/* Pipe-throughput probes used by bench.py to measure the FP32 / SFU roofline denominators live:
 * which = 0: FFMA, 1: FFMA2 (packed f32x2), 2: MUFU.EX2.  Runs `iters` dependent-chain steps x 8 chains per
 * thread on `blocks` x 256 threads; out[blocks*256] receives a checksum.  ops per thread = iters * 8
 * (FFMA2 counts 2 FMAs per instruction). */
int dicp_pipe_probe(int which, int blocks, int iters, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DICP_B200_H */
