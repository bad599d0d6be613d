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
    python bench.py --impl reference ...                      CPU arm: the unmodified reference package (baseline/_ref)

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
# CPU arm: the REFERENCE ITSELF (unmodified package from baseline/_ref or /root/reference through oracle/ref_loader.py, torch
# twin because pykeops is absent) on a bounded sample of the same workload; the oracle port only if the package is missing
CPU_SAMPLE_M = 1536          # the step at M = N = 1536 instead of 20 000: ~1 s of CPU work per step on the box's 16 cores


def load_reference_or_none():
    try:
        from oracle import ref_loader
        if ref_loader.find_root() is None:
            return None
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):          # the package prints on import
            return ref_loader.load_reference()
    except Exception as e:          # noqa: BLE001 -- reported in the JSON line
        print(f"[bench] reference package could not be loaded: {e!r}", file=sys.stderr)
        return None


def reference_closure(ref, version, scheme, nt, sigma, lam, q, p0, x, y, inv2s2):
    """One L-BFGS closure evaluation exactly as the reference runs it (tools/optim.py:34-47 calling the `lossfunc` of
    core/LDDMM.py:363-371): Shoot + trajloss + quadratic data loss (core/PSR.py:498-516), then L.backward() through its
    own integrator loop and torch-twin reductions."""
    LM = ref.LDDMM.LDDMMModel(sigma=sigma, D=q.shape[1], lambd=lam, spec={"device": "cpu", "dtype": torch.float32},
                              version=version, computversion="torch", scheme=scheme, nt=nt)
    p = p0.clone().requires_grad_(True)
    shoot = LM.Shoot(q, p, x)
    moved = shoot[-1][0] if x is None else shoot[-1][3]
    L = LM.trajloss(shoot) + ((moved - y) ** 2).sum() * inv2s2
    L.backward()
    return float(L), p.grad


def cpu_step_reference(ref, variant, xA, y, p0, m, threads):
    """The bench step (one EM step + one closure evaluation) of the reference at M = N = m: every m-th point of the same
    clouds.  Returns (seconds, pairs)."""
    torch.set_num_threads(threads)
    idx = torch.arange(0, xA.shape[0], max(1, xA.shape[0] // m))[:m]
    q, ys, ps = xA[idx].contiguous(), y[idx].contiguous(), p0[idx].contiguous()
    t0 = time.perf_counter()
    G = ref.GMM.GaussianMixtureUnif(ys, sigma=SIGMA_GMM, spec={"device": "cpu", "dtype": torch.float32}, computversion="torch")
    G.to_optimize = {"mu": False, "sigma": True, "w": False, "eta0": False}
    Y, Cfe, FE = G.EM_step(q)                                    # core/GMM.py:236-325 (dense (m,m) block)
    reference_closure(ref, variant, "Euler", NT, SIGMA_LDDMM, LAMBDA_LDDMM, q, ps, None, Y, 1.0 / (2 * SIGMA_GMM ** 2))
    return time.perf_counter() - t0, pairs_per_step(len(idx))


def cpu_step_port(variant, xA, y, p0, m, threads):
    """Same step through the oracle port (only when the reference package is not available on this machine)."""
    from oracle.gmm import GMMOracle
    from oracle.lddmm import LDDMMOracle
    torch.set_num_threads(threads)
    idx = torch.arange(0, xA.shape[0], max(1, xA.shape[0] // m))[:m]
    q, ys, ps = xA[idx].contiguous(), y[idx].contiguous(), p0[idx].contiguous()
    t0 = time.perf_counter()
    O = GMMOracle(ys, SIGMA_GMM, to_optimize={"mu": False, "sigma": True, "w": False, "eta0": False})
    Y, _, _ = O.em_step(q, variant="torch")
    OR = LDDMMOracle(sigma=SIGMA_LDDMM, D=DIM, lambd=LAMBDA_LDDMM, version=variant, scheme="Euler", nt=NT)
    p = ps.clone().requires_grad_(True)
    L, _ = OR.loss(q, p, None, Y, 1.0 / (2 * SIGMA_GMM ** 2))
    L.backward()
    return time.perf_counter() - t0, pairs_per_step(len(idx))


def cpu_arm(variant, xA, y, p0, threads, steps=None, warmup=1, budget_s=None):
    """Runs the CPU arm: `steps` timed steps (or as many as fit in budget_s seconds, at least 2).  Returns the
    cpu_baseline dict and (total seconds, steps done)."""
    ref = load_reference_or_none()
    step = (lambda: cpu_step_reference(ref, variant, xA, y, p0, CPU_SAMPLE_M, threads)) if ref is not None else \
        (lambda: cpu_step_port(variant, xA, y, p0, CPU_SAMPLE_M, threads))
    for _ in range(warmup):
        step()
    tot_t, tot_pairs, n = 0.0, 0.0, 0
    t_start = time.perf_counter()
    while (steps is not None and n < steps) or (steps is None and (n < 2 or time.perf_counter() - t_start < budget_s)):
        dt, npairs = step()
        tot_t, tot_pairs, n = tot_t + dt, tot_pairs + npairs, n + 1
    kind = "reference" if ref is not None else "port"
    what = ("the unmodified reference package (diffICP.core.GMM.EM_step + LDDMMModel.Shoot / trajloss + backward(), torch twin: "
            "pykeops absent)" if ref is not None else "the oracle port of the reference's algorithm (reference package not found)")
    sample = (f"the bench step (1 EM step + 1 closure evaluation: shoot Euler nt={NT}, {variant} model, loss, autograd backward) at "
              f"M=N={CPU_SAMPLE_M} (every {M_POINTS // CPU_SAMPLE_M}-th point of the same 20k clouds; dense 20k x 20k does not fit the "
              f"reference's dense torch path: 4.8 GB per (M,N,D) temporary, tools/kernel.py:104), {n} steps, {what}")
    return {"value": tot_pairs / tot_t, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}, tot_t, n


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    xA, y, p0 = make_workload(1234)
    cpu, tot_t, n = cpu_arm(args.variant, xA, y, p0, threads, steps=args.steps, warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.variant),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(variant):
    return {"workload": "two_set_3d_20k_dense: LBFGS closure evaluation (shoot + loss + adjoint), "
                        f"M=N={M_POINTS}, D={DIM}, Euler nt={NT}, sigma_LDDMM={SIGMA_LDDMM}, lambda={LAMBDA_LDDMM}",
            "model_variant": variant, "pairs_per_step": pairs_per_step(M_POINTS),
            "pairs_note": "ordered (i,j) pairs of the algorithm; the adjoint evaluates each UNORDERED pair once (symmetric engine), so the "
                          f"physical pair evaluations per step are M*C + nt*M^2 + nt*M^2/2 = {float(M_POINTS) ** 2 * (1 + 1.5 * NT):.3g}", "gmm": f"C={M_POINTS} centroids frozen, sigma optimised",
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

    # ---- the configuration BASELINE.json names for 8 GPUs (configs[3], "diffICP_full": 3 structures, 3-D, 50k points per
    # frame, frames sharded k mod N), at a FIXED total of C4_FRAMES frames whatever N: strong scaling of compute-bound work
    if groupwise is not None and not args.no_c4:
        from groupwise_c4 import run_c4
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):
            c4 = run_c4(rank, world, dev, GMM.comm, n_frames=C4_FRAMES, n_points=50000, iters=2)
        if c4 is not None:
            fe_rel = abs(c4["FE"] - C4_FE_1GPU) / abs(C4_FE_1GPU) if C4_FE_1GPU else None
            groupwise["c4_shaped"] = {
                "config": f"configs[3]-shaped: {C4_FRAMES} frames x 3 structures x 50k points, 3-D, {c4['support_points']} support points, "
                          f"{c4['model']}; the SAME {C4_FRAMES} frames at every N (frames k mod N) = strong scaling",
                "n_gpus": world, "frames": C4_FRAMES, "iteration_ms": c4["iteration_ms_steady"],
                "gmm_opt_ms": c4["gmm_opt_ms"][-1], "reg_opt_ms": c4["reg_opt_ms"][-1],
                "first_iteration_ms": c4["gmm_opt_ms"][0] + c4["reg_opt_ms"][0],
                "FE": c4["FE"], "FE_1gpu_reference": C4_FE_1GPU, "FE_rel_diff_vs_1gpu": fe_rel,
                # N ranks reduce the EM statistics in another order than one rank; after two L-BFGS registrations of every frame
                # that rounding shows up at ~1e-5 of the free energy (measured 1.5e-5 at N = 2): the bar is 1e-4
                "FE_matches_1gpu_to_1e-4": None if fe_rel is None else bool(fe_rel < 1e-4),
                "sigma": c4["sigma"], "timing": "host wall clock around synchronised GMM_opt + Reg_opt, max over ranks, 2nd iteration"}

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
        cpu, _, _ = cpu_arm(args.variant, xA, y, p0, threads, steps=None, warmup=1, budget_s=12.0)

    if groupwise is not None and cpu is not None:
        # CPU reference for the groupwise iteration: ONE closure evaluation of one frame (the reference's algorithm: separate
        # reductions + autograd through the Euler loop), scaled by the measured ~23.5 closures per frame
        from groupwise_iteration import spiral_frames
        torch.set_num_threads(threads)
        fr = spiral_frames(1, groupwise["points_per_frame"])[0]
        nq = groupwise["support_points"]
        side = int(round(nq ** 0.5))
        gx = torch.linspace(float(fr[:, 0].min()), float(fr[:, 0].max()), side)
        gy = torch.linspace(float(fr[:, 1].min()), float(fr[:, 1].max()), max(1, nq // side))
        qg = torch.stack(torch.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2).contiguous()
        pc = 1e-3 * torch.randn(qg.shape)
        ref = load_reference_or_none()
        t0 = time.perf_counter()
        if ref is not None:
            reference_closure(ref, "hybrid", "Euler", 10, 0.2, 500.0, qg, pc, fr, fr + 0.01, 50.0)
            how = "the reference package itself (torch twin)"
        else:
            from oracle.lddmm import LDDMMOracle
            OR = LDDMMOracle(sigma=0.2, D=2, lambd=500.0, version="hybrid", scheme="Euler", nt=10)
            pc.requires_grad_(True)
            Lc, _ = OR.loss(qg, pc, fr, fr + 0.01, 50.0)
            Lc.backward()
            how = "oracle port"
        tc = time.perf_counter() - t0
        groupwise["cpu_reference_estimate_ms"] = 1e3 * tc * 23.5 * groupwise["frames"]
        groupwise["cpu_reference_note"] = (f"{how}, {threads} threads: one closure evaluation of one frame took "
                                           f"{1e3 * tc:.0f} ms; x 23.5 closures/frame (SURVEY.md 3.3) x {groupwise['frames']} frames (EM excluded)")

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


REF_REDUCTIONS = {"classic": 2, "hybrid": 3, "logdet": 7}


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

    def graph_time(fn, reps=10):
        """Device time per call as the call runs inside the step: `reps` calls captured in one CUDA graph (the closure of the
        timed step is such a graph), events around the replay -- no host launch latency between the call's launches."""
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        return timeit(g.replay, n=10) / reps

    t_adj_eager = timeit(lambda: shooting._vjp(sp, state, lam, G, ws))
    t_fwd_eager = timeit(lambda: shooting._rhs(sp, state, F, ws))
    t_adj = graph_time(lambda: shooting._vjp(sp, state, lam, G, ws))
    t_fwd = graph_time(lambda: shooting._rhs(sp, state, F, ws))

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
    sm_mhz = 1965.0
    try:
        sm_mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
    except Exception:
        pass
    lanes_clock = sms * 128 * sm_mhz * 1e6                     # FP32 lane-operations / s at the maximum SM clock
    return {
        "frac_lanes_clock": fp_rate_adj / lanes_clock,
        "lanes_clock_note": f"hardware-theoretical denominator: {sms} SMs x 128 FP32 lanes x {sm_mhz:.0f} MHz (MEASURED_PEAKS.json "
                            "sm_max_mhz) = %.3g lane-ops/s; `frac` uses the live FFMA probe instead" % lanes_clock,
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
        "timing_note": "s_per_launch = device time of one dicp_rhs_adjoint / dicp_rhs_forward call (pack + pair kernel + merge "
                       "launches) as it runs inside the step's CUDA graph: 10 calls per captured graph, CUDA events around "
                       "the replay, median of 10; s_per_launch_eager = the same call launched from the host one by one "
                       "(includes host launch latency between its launches)",
        "adjoint": {"s_per_launch_eager": t_adj_eager, "s_per_launch": t_adj, "pairs_per_s": pairs / t_adj, "fp32_per_pair": alg["adj_fp32"], "frac": frac_adj,
                    "frac_lanes_clock": fp_rate_adj / lanes_clock,
                    "physical_pair_evaluations_per_s": (0.5 if symmetric else 1.0) * pairs / t_adj},
        "forward": {"s_per_launch_eager": t_fwd_eager, "s_per_launch": t_fwd, "pairs_per_s": pairs / t_fwd, "fp32_per_pair": alg["fwd_fp32"], "frac": frac_fwd,
                    "frac_lanes_clock": fp_rate_fwd / lanes_clock,
                    # SURVEY.md 8(d), metric 1: physical pairs/s x the reference reductions ONE fused visit of a pair replaces
                    # (x = None: classic KRed + GenDKRed = 2, hybrid + GradKRed = 3, logdet 7; core/LDDMM.py:176-227)
                    "reference_reductions_replaced": REF_REDUCTIONS[args.variant],
                    "reference_equivalent_pairs_per_s": REF_REDUCTIONS[args.variant] * pairs / t_fwd},
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
                   "replay; algorithmic bytes per point: 12D+8 per EM step in three sweeps (SURVEY.md 8d): pass 1 reads 4D and "
                   "writes 4, the column pass reads 4D+4, pass 2 reads 4D+4 and writes 4D; with <= 64 components the first two "
                   "are ONE sweep that reads 4D (first_sweep_fused), i.e. 12D+4 per EM step; em_step_s = the sweeps an EM "
                   "step actually runs"}
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
            # row LSE + column statistics from ONE read of X (C <= 64: one launch; more components: the two sweeps above)
            "first_sweep_fused": (lambda: em_ops.lse_colstats(sig, X, mu, wl2), 22, 4 * D),
        }
        total = 0.0
        fused = C <= em_ops.SMALL_C
        for name, (fn, fp, nbytes) in passes.items():
            t = graph_time(fn)
            if name == "row_full" or (name == "first_sweep_fused") == fused:
                total += t          # what an EM step runs: fused first sweep + full row pass, or the three sweeps
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
# configs[3]-shaped strong-scaling entry: total frames (divisible by 8) and the free energy the 1-GPU run reaches after its
# 2 iterations (deterministic kernels, seeded synthetic frames): every N must reproduce it to 1e-4
C4_FRAMES = 64
C4_FE_1GPU = -6932241.603818208         # measured: 1 x B200, profiles/r02_bench_1gpu.json (lock-step registration of all frames)

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
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[3]-shaped strong-scaling entry (~25 s on one GPU)")
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
