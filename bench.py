#!/usr/bin/env python
"""bench.py -- headline benchmark of the diffICP hot path on B200.

Workload (BASELINE.json configs[1], "api two-point-set diffeomorphic ICP matching, 3D synthetic clouds 20k x 20k"):
one STEP = one slice of an ICP_two_set iteration on a 20 000-point 3-D cloud registered to a 20 000-centroid GMM with
dense support (support points = data points, the only 3-D-capable scheme of the reference):
  (1) one EM step of the GMM (E step 20k x 20k, mu/w frozen, sigma re-estimated; api/ICP_two_set.py:179-187) giving the
      quadratic targets; with N > 1 GPUs every rank holds its own frame and the sigma statistics are all-reduced (NCCL) --
      the only collective of the algorithm;
  (2) one L-BFGS closure evaluation of the registration: geodesic shoot (Euler, nt = 10) + trajectory loss + quadratic
      data loss, then the adjoint sweep (= `L.backward()` of the reference, tools/optim.py:34-47).  The reference spends
      > 99 % of an ICP iteration in these evaluations (~23 per frame per outer iteration).
The LDDMM model is the API default of the two-set path: the full logdet model; --variant selects
hybrid / classic.

Metric: Gaussian kernel pairs / second, where one "pair" is one (i,j) visit in the E step, in one right-hand-side
evaluation or in one adjoint evaluation: pairs/step = M*C + 2 * nt * M^2 (implementation independent: the reference
visits each such pair 2-7 times per evaluation with separate reductions, this build once).

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference ...                      CPU arm: the oracle port of the reference's algorithm

Prints ONE JSON line (rank 0).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_POINTS = 20000
DIM = 3
NT = 10
SIGMA_LDDMM = 0.2
LAMBDA_LDDMM = 500.0
SIGMA_GMM = 0.1
METRIC = "gaussian_kernel_pairs_per_s"
UNIT = "pairs/s"


def make_workload(seed, M=M_POINTS, D=DIM):
    """xA ~ U[0,1]^3; targets y = xA warped by a smooth random field + N(0, 0.01^2)."""
    g = torch.Generator().manual_seed(seed)
    xA = torch.rand(M, D, generator=g)
    cen = torch.rand(8, D, generator=g)
    amp = 0.05 * torch.randn(8, D, generator=g)
    w = torch.exp(-((xA[:, None, :] - cen[None]) ** 2).sum(-1) / (2 * 0.3 ** 2))
    y = xA + w @ amp + 0.01 * torch.randn(M, D, generator=g)
    p0 = 1e-3 * torch.randn(M, D, generator=g)
    return xA.contiguous(), y.contiguous(), p0.contiguous()


def pairs_per_step(M, nt=NT):
    return float(M) * M + 2.0 * nt * M * M


# ------------------------------------------------------------------------------------------------------------------
# clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port of its torch twin) on a bounded row sample of the same workload
def cpu_sample_seconds(variant, m_rows, xA, p0, threads):
    """One right-hand-side evaluation + its reverse-mode gradient for `m_rows` rows against all M columns, done the
    way the reference does it: separate dense reductions (tools/kernel.py:186-203) + torch autograd."""
    from oracle.kernels import GaussOracle
    torch.set_num_threads(threads)
    K = GaussOracle(SIGMA_LDDMM, DIM, chunk=256)
    idx = torch.arange(0, xA.shape[0], max(1, xA.shape[0] // m_rows))[:m_rows]
    q = xA.clone().requires_grad_(True)
    p = p0.clone().requires_grad_(True)
    t0 = time.perf_counter()
    qs, ps = q[idx], p[idx]
    eta = 1.0 / LAMBDA_LDDMM if variant == "logdet" else 0.0
    vq = K.KRed(qs, q, p)
    Gq = K.GenDKRed(qs, q, p, ps)
    L = (vq ** 2).sum() + (Gq ** 2).sum()
    if variant in ("hybrid", "logdet"):
        L = L + (ps * K.GradKRed(qs, q)).sum()
    if variant == "logdet":
        vq2 = K.GradKRed(qs, q)
        Gq2 = K.HessKRed(qs, q, p, ps)
        Gq3 = K.GradLapKRed(qs, q)
        L = L + eta * (vq2 ** 2).sum() + eta * (Gq2 ** 2).sum() + eta ** 2 * (Gq3 ** 2).sum() + eta * K.LapKRed(qs, q).sum()
    L.backward()
    dt_rhs = time.perf_counter() - t0
    # E step of the same rows against all M centroids (dense (m,M) block, core/GMM.py:263-317)
    t0 = time.perf_counter()
    with torch.no_grad():
        d2 = ((xA[idx][:, None, :] - xA[None, :, :]) ** 2).sum(-1)
        t = -d2 / (2 * SIGMA_GMM ** 2)
        T = t.logsumexp(1)
        gam = (t - T[:, None]).exp()
        Y = gam @ xA
        _ = (gam * d2).sum() + (gam * (t - T[:, None])).sum() + (Y ** 2).sum()
    dt_em = time.perf_counter() - t0
    # same workload mix as the GPU arm: a step has ONE E step per nt (RHS + adjoint) evaluations, so the sampled E step
    # enters with weight 1/nt in both the time and the pair count
    return dt_rhs + dt_em / NT, len(idx) * xA.shape[0] * (2.0 + 1.0 / NT)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    xA, y, p0 = make_workload(1234)
    m_rows = 256
    cpu_sample_seconds(args.variant, 64, xA, p0, threads)            # warm-up of the thread pool
    for _ in range(max(0, args.warmup - 1)):
        cpu_sample_seconds(args.variant, m_rows, xA, p0, threads)
    tot_t, tot_pairs = 0.0, 0.0
    for _ in range(args.steps):
        dt, npairs = cpu_sample_seconds(args.variant, m_rows, xA, p0, threads)
        tot_t += dt
        tot_pairs += npairs
    value = tot_pairs / tot_t
    sample = (f"{m_rows} of {M_POINTS} rows x {M_POINTS} columns per step: one RHS evaluation + autograd backward, plus "
              f"1/{NT} of one dense E step (the mix of a full step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.variant),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(variant):
    return {"workload": "two_set_3d_20k_dense: LBFGS closure evaluation (shoot + loss + adjoint), "
                        f"M=N={M_POINTS}, D={DIM}, Euler nt={NT}, sigma_LDDMM={SIGMA_LDDMM}, lambda={LAMBDA_LDDMM}",
            "model_variant": variant, "pairs_per_step": pairs_per_step(M_POINTS), "gmm": f"C={M_POINTS} centroids frozen, sigma optimised",
            "l2": "flushed between timed steps (256 MiB memset, outside the event pairs)"}


# ------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    from diff_icp_b200 import ops
    from diff_icp_b200.core.LDDMM import LDDMMModel

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    spec = {"device": dev, "dtype": torch.float32}
    xA, y, p0 = make_workload(1234 + rank)
    M = xA.shape[0]
    LM = LDDMMModel(sigma=SIGMA_LDDMM, D=DIM, lambd=LAMBDA_LDDMM, spec=spec, version=args.variant, scheme="Euler", nt=NT)
    LM.use_cuda_graph = not args.no_graph
    inv2s2 = 1.0 / (2 * SIGMA_GMM ** 2)
    inv_w = torch.full((M,), inv2s2, device=dev)
    q_d, y_d, p_d = xA.to(dev), y.to(dev), p0.to(dev)

    from diff_icp_b200.core.GMM import GaussianMixtureUnif
    GMM = GaussianMixtureUnif(y.to(dev), sigma=SIGMA_GMM, spec=spec)
    GMM.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
    if world > 1:
        from diff_icp_b200.dist import StatsComm
        GMM.comm = StatsComm()

    def closure(q, p_init, yy):
        """(1) EM step on the current points -> quadratic targets; (2) the L-BFGS closure exactly as DiffPSR.Reg_opt /
        LDDMMModel.Optimize evaluate it: shoot + lambda*H + cost + quadratic data loss + adjoint, one captured launch
        sequence, loss and gradient returned to the host."""
        GMM.sigma = SIGMA_GMM
        Y, Cfe, FE = GMM.EM_step(q)
        dataloss = lambda x: ((x - Y) ** 2).sum() * inv2s2
        dataloss.targets, dataloss.inv2sig2 = Y, inv_w
        evaluate = LM.closure_evaluator(dataloss, q)
        return evaluate(p_init)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # launches per step (counted on an eager step: graph replays do not pass through the launch counter)
    lib = ops.load()
    LM.use_cuda_graph = False
    closure(q_d, p_d, y_d)
    torch.cuda.synchronize()
    c0 = lib.dicp_launch_count()
    closure(q_d, p_d, y_d)
    torch.cuda.synchronize()
    launches_per_step = int(lib.dicp_launch_count() - c0)
    LM.use_cuda_graph = not args.no_graph

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)       # nvidia-smi needs ~0.2 s to deliver its first sample: start before warm-up
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        closure(q_d, p_d, y_d)
    barrier()

    # ---- device-resident timing ------------------------------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for e0, e1 in ev:
        flush.zero_()
        e0.record()
        closure(q_d, p_d, y_d)
        e1.record()
    barrier()
    ms_total = sum(e0.elapsed_time(e1) for e0, e1 in ev)

    # ---- end to end through the public API with host buffers -------------------------------------------------
    q_h, y_h, p_h = xA.pin_memory(), y.pin_memory(), p0.pin_memory()
    g_h = torch.empty_like(p0).pin_memory()
    h2d = (q_h.numel() + y_h.numel() + p_h.numel()) * 4
    d2h = g_h.numel() * 4 + 4

    def e2e_step():
        q = q_h.to(dev, non_blocking=True)
        yy = y_h.to(dev, non_blocking=True)
        pp = p_h.to(dev, non_blocking=True)
        GMM.mu = yy                          # the template (GMM centroids) also arrives from the host
        L, g = closure(q, pp, yy)            # loss (float) and gradient (pinned host buffer): the D2H read of the step
        g_h.copy_(g)
        return float(L)

    e2e_step()
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    for e0, e1 in ev2:
        e0.record()
        e2e_step()
        e1.record()
    barrier()
    wall_e2e = time.perf_counter() - t0
    ms_e2e_total = max(sum(e0.elapsed_time(e1) for e0, e1 in ev2), 0.0)
    if rank == 0 and world == 1:             # short runs: keep the GPU under the same load until a few samples exist
        t_wait = time.perf_counter()
        while len(sampler.rows) < 5 and time.perf_counter() - t_wait < 3.0:
            closure(q_d, p_d, y_d)
    clocks = sampler.stop() if rank == 0 else None

    # ---- second half of BASELINE.json's metric: groupwise PSR iteration time (frames sharded over the ranks) ----------
    groupwise = None
    if not args.no_groupwise:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from groupwise_iteration import run_groupwise
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):          # the algorithm's own progress / warning prints
            groupwise = run_groupwise(rank, world, dev, GMM.comm, iters=3)
            if world > 1:       # the 64-frame atlas is latency-bound once sharded: also report it at 64 frames PER rank
                weak = run_groupwise(rank, world, dev, GMM.comm, iters=3, weak=True)
                if groupwise is not None and weak is not None:
                    groupwise["weak_scaling_64_frames_per_gpu"] = {k: weak[k] for k in
                                                                   ("frames", "scaling", "gmm_opt_ms", "reg_opt_ms",
                                                                    "iteration_ms_steady", "FE", "sigma")}

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e_total, wall_e2e * 1e3], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_e2e_total, wall_ms = (float(v) for v in t)
        wall_e2e = wall_ms / 1e3
    if rank != 0:
        return

    pps = pairs_per_step(M)
    value = world * pps * args.steps / (ms_total * 1e-3)
    e2e_time = max(ms_e2e_total * 1e-3, wall_e2e)            # host-visible time: includes the D2H sync of every step
    e2e_value = world * pps * args.steps / e2e_time

    # ---- roofline of the dominant kernel (adjoint pair kernel), timed alone with CUDA events ------------------
    roof = kernel_roofline(args, LM, q_d, p_d, dev, ops)

    # ---- CPU baseline beside it (rank 0, bounded sample) -------------------------------------------------------
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline:
        cpu_sample_seconds(args.variant, 64, xA, p0, threads)
        tt, pp_ = 0.0, 0.0
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < 10.0:
            dt, npairs = cpu_sample_seconds(args.variant, 256, xA, p0, threads)
            tt += dt
            pp_ += npairs
        cpu = {"value": pp_ / tt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"256 of {M} rows x {M} columns per evaluation (RHS + autograd backward + 1/{NT} E step), repeated for 10 s"}

    if groupwise is not None and cpu is not None:
        # CPU reference for the groupwise iteration: ONE closure evaluation of one frame (the reference's algorithm: separate
        # reductions + autograd through the Euler loop), scaled by the measured ~23.5 closures per frame
        from oracle.lddmm import LDDMMOracle
        from groupwise_iteration import spiral_frames
        torch.set_num_threads(threads)
        fr = spiral_frames(1, groupwise["points_per_frame"])[0]
        nq = groupwise["support_points"]
        side = int(round(nq ** 0.5))
        gx = torch.linspace(float(fr[:, 0].min()), float(fr[:, 0].max()), side)
        gy = torch.linspace(float(fr[:, 1].min()), float(fr[:, 1].max()), max(1, nq // side))
        qg = torch.stack(torch.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2)
        OR = LDDMMOracle(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10)
        pc = (1e-3 * torch.randn(qg.shape)).requires_grad_(True)
        t0 = time.perf_counter()
        Lc, _ = OR.loss(qg, pc, fr, fr + 0.01, 50.0)
        Lc.backward()
        tc = time.perf_counter() - t0
        groupwise["cpu_reference_estimate_ms"] = 1e3 * tc * 23.5 * groupwise["frames"]
        groupwise["cpu_reference_note"] = (f"oracle port, {threads} threads: one closure evaluation of one frame took "
                                           f"{1e3 * tc:.0f} ms; x 23.5 closures/frame x {groupwise['frames']} frames (EM excluded)")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.variant),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_time / args.steps},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "cuda_graph": LM.use_cuda_graph, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "groupwise_psr_iteration": groupwise,
    }
    print(json.dumps(line), flush=True)


def kernel_roofline(args, LM, q, p, dev, ops):
    """Time dicp_rhs_adjoint and dicp_rhs_forward alone (CUDA events, median of 20), measure the FP32 / SFU pipe peaks
    live with the probe kernels, and report the binding-pipe fraction."""
    from diff_icp_b200 import shooting
    M, D = q.shape
    sp = LM._spec_for(M, 0, dev)
    ws = torch.empty(int(ops.load().dicp_pair_workspace_bytes(M, M)), dtype=torch.uint8, device=dev)
    state = torch.cat([q.reshape(-1), p.reshape(-1), torch.zeros(1, device=dev)])
    lam = torch.randn(sp.S, device=dev)
    F = torch.zeros(sp.S + 3, device=dev)
    G = torch.zeros(sp.S, device=dev)

    def timeit(fn, n=20):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    t_adj = timeit(lambda: shooting._vjp(sp, state, lam, G, ws))
    t_fwd = timeit(lambda: shooting._rhs(sp, state, F, ws))

    sms = ops.load().dicp_sm_count()
    peaks = {}
    for name, which, per in (("ffma", 0, 1), ("mufu_ex2", 2, 1)):
        blocks, iters = sms * 8, 8000
        out = torch.empty(blocks * 256, device=dev)
        ops.pipe_probe(which, blocks, 100, out)
        t = min(timeit(lambda: ops.pipe_probe(which, blocks, iters, out), n=5) for _ in range(2))
        peaks[name] = blocks * 256 * iters * 8 * per / t

    # algorithmic FP32 instructions / MUFU per pair of the adjoint and forward kernels (DESIGN.md §6, D = 3)
    alg = dict(ALG_WORK[args.variant])
    symmetric = bool(ops.load().dicp_sym_mode(-1)) and M >= 4096
    if not symmetric:
        alg["adj_fp32"] = alg["adj_fp32_ordered"]
    pairs = float(M) * M
    fp_rate_adj = alg["adj_fp32"] * pairs / t_adj
    sfu_rate_adj = (0.5 if symmetric else 1.0) * pairs / t_adj          # one MUFU.EX2 per evaluated pair
    frac_adj = max(fp_rate_adj / peaks["ffma"], sfu_rate_adj / peaks["mufu_ex2"])
    fp_rate_fwd = alg["fwd_fp32"] * pairs / t_fwd
    frac_fwd = max(fp_rate_fwd / peaks["ffma"], pairs / t_fwd / peaks["mufu_ex2"])
    em = em_roofline(dev, timeit, peaks)
    return {
        "bound": "fp32_pipe",
        "kernel": ("sym_pair_kernel<AdjQQ*> (dicp_rhs_adjoint, symmetric engine)" if symmetric
                   else "pair_kernel_p<AdjQQ*> (dicp_rhs_adjoint)"),
        "achieved": 2 * fp_rate_adj / 1e12, "peak": 2 * peaks["ffma"] / 1e12, "unit": "TFLOP/s", "frac": frac_adj,
        "traffic": NCU_DRAM_BYTES.get(args.variant),
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the adjoint kernel, ncu --set full "
                        "(profiles/r01c_ncu_full_bench_kernels.csv); algorithmic bytes = 2 x 20000 x 48 B = 1.92 MB",
        "note": "achieved = algorithmic FP32 instructions per ordered pair x M^2 pairs / event time, x2 flop; the adjoint "
                "evaluates every UNORDERED pair once (symmetric engine), its per-ordered-pair count is half the unordered "
                "one; peak = FFMA issue rate measured live by dicp_pipe_probe (x2 flop), of measured; frac = binding-pipe "
                "utilisation",
        "adjoint": {"s_per_launch": t_adj, "pairs_per_s": pairs / t_adj, "fp32_per_pair": alg["adj_fp32"], "frac": frac_adj},
        "forward": {"s_per_launch": t_fwd, "pairs_per_s": pairs / t_fwd, "fp32_per_pair": alg["fwd_fp32"], "frac": frac_fwd},
        "measured_peaks": {"ffma_per_s": peaks["ffma"], "mufu_ex2_per_s": peaks["mufu_ex2"], "sms": sms},
        "em_step": em,
    }


def em_roofline(dev, timeit, peaks):
    """E step: the three fused passes timed alone as device time (each pass captured 10x in a CUDA graph, so host launch
    latency does not pollute these ~10-100 us calls), at three shapes:
      atlas   640k points x 50 components, 2-D  (configs[2]): pipe-bound, small => partly latency-bound
      two_set 20k points x 20k components, 3-D  (configs[1], the E step inside this bench's step): pipe-bound, large
      few_components  4M points x 8 components, 3-D: the regime (C <~ 9) where HBM binds
    north_star asks for the achieved HBM GB/s of the E step next to its FP32 / SFU fraction: both are reported, against the
    measured HBM peak of MEASURED_PEAKS.json (fallback: the same pool's 6549 GB/s) and the live FFMA / MUFU probes."""
    import math
    from diff_icp_b200 import em_ops
    hbm = 6549.4
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "MEASURED_PEAKS.json"
    except Exception:
        src = "fallback (MEASURED_PEAKS.json absent)"

    def graph_time(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        return timeit(g.replay, n=10) / reps

    out = {"hbm_peak_gbs": hbm, "hbm_peak_source": src,
           "note": "device time per call = pack + pair kernel (+ split merge / scalar reduction), 10 calls per CUDA-graph "
                   "replay; algorithmic bytes per point: 12D+8 per EM step (SURVEY.md 8d): pass 1 reads 4D and writes 4, "
                   "the column pass reads 4D+4, pass 2 reads 4D+4 and writes 4D"}
    for tag, (N, C, D, sig) in {"atlas": (640000, 50, 2, 0.05), "two_set": (20000, 20000, 3, 0.1),
                                "few_components": (4000000, 8, 3, 0.2)}.items():
        g = torch.Generator().manual_seed(7)
        X = torch.rand(N, D, generator=g).to(dev)
        mu = torch.rand(C, D, generator=g).to(dev)
        w = torch.zeros(C, device=dev)
        lgn = D * (math.log(sig) + 0.5 * math.log(2 * math.pi))
        wl2 = ((w - torch.logsumexp(w, 0) - lgn) * 1.4426950408889634).contiguous()
        lpi = (w - torch.logsumexp(w, 0)).contiguous()
        T2 = em_ops.rowpass(sig, X, mu, wl2)
        res = {"workload": f"{N} points x {C} components, D={D}"}
        # name: (callable, algorithmic FP32 instr / pair, algorithmic HBM bytes / point)
        passes = {
            "row_lite": (lambda: em_ops.rowpass(sig, X, mu, wl2), 8, 4 * D + 4),
            "col_stats": (lambda: em_ops.colstats(sig, X, T2, mu, wl2), 14, 4 * D + 4),
            "row_full": (lambda: em_ops.rowpass(sig, X, mu, wl2, mu, lpi), 17, 4 * D + 4 + 4 * D),
        }
        total = 0.0
        for name, (fn, fp, nbytes) in passes.items():
            t = graph_time(fn)
            total += t
            pairs = float(N) * C
            res[name] = {"s_per_call": t, "pairs_per_s": pairs / t, "fp32_per_pair": fp,
                         "frac_fp32": fp * pairs / t / peaks["ffma"], "frac_sfu": pairs / t / peaks["mufu_ex2"],
                         "hbm_gbs": nbytes * N / t / 1e9, "frac_hbm": nbytes * N / t / 1e9 / hbm}
        res["em_step_s"] = total
        res["binding"] = max(("fp32", max(res[k]["frac_fp32"] for k in passes)), ("sfu", max(res[k]["frac_sfu"] for k in passes)),
                             ("hbm", max(res[k]["frac_hbm"] for k in passes)), key=lambda kv: kv[1])[0]
        out[tag] = res
        del X, T2
    return out


# DRAM bytes per launch of the adjoint kernel at 20k x 20k (one ncu --set full capture per variant, profiles/):
# dram__bytes_read.sum + dram__bytes_write.sum; the row / column partials of the symmetric engine (~40 MB) stay in L2
NCU_DRAM_BYTES = {"classic": None, "hybrid": None, "logdet": 1990144}

# algorithmic FP32 instruction counts per ORDERED pair, D = 3 (hand count of the formulas in csrc/ops_rhs.cuh).
# The adjoint (q,q) pass runs on the symmetric engine: an unordered pair costs the shared part once plus both sides'
# accumulations -- 53 / 74 / 99 operations (classic / hybrid / logdet) instead of 2 x 41 / 56 / 83 -- i.e. 26.5 / 37 / 49.5
# per ordered pair.  With DICP_SYM=0 (general engine) the counts are the ordered ones.
ALG_WORK = {
    "classic": {"fwd_fp32": 16, "adj_fp32": 26.5, "adj_fp32_ordered": 41},
    "hybrid": {"fwd_fp32": 19, "adj_fp32": 37.0, "adj_fp32_ordered": 56},
    "logdet": {"fwd_fp32": 30, "adj_fp32": 49.5, "adj_fp32_ordered": 83},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="logdet", choices=["classic", "hybrid", "logdet"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-groupwise", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
