"""CPU: the lock-step L-BFGS (csrc/lbfgs_batch.cu, host arithmetic behind the C ABI) against torch.optim.LBFGS driven by
the sequential LBFGS_optimization (the reference's driver, tools/optim.py:10-110), frame by frame, on analytic test
functions: same trajectories up to rounding, same control flow (step counts, evaluation counts +-1), frames independent."""
import numpy as np
import torch

from diff_icp_b200.tools.optim import LBFGS_optimization, LBFGS_optimization_lockstep, LockstepLBFGS


def rosen(x):
    return (100 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2).sum()


def bowl(x):
    a = torch.arange(1, x.numel() + 1, dtype=x.dtype)
    return (a * x ** 2).sum() + 0.1 * (x ** 4).sum() + torch.sin(x).sum()


def logcosh(x):
    return torch.log(torch.cosh(x - 0.3)).sum() + 0.5 * (x[0] * x[-1]) ** 2


FUNCS = [rosen, bowl, logcosh, bowl, rosen]
SIZES = [5, 8, 12, 3, 2]


class Evaluator:
    """numpy stand-in for shooting.BatchedClosurePlan: same buffers, closures evaluated with torch autograd in fp32."""

    def __init__(self, funcs, sizes):
        K, stride = len(funcs), max(sizes) + 3
        self.funcs, self.sizes = funcs, sizes
        self.X = np.zeros((K, stride), np.float32)
        self.active = np.zeros(K, np.uint8)
        self.losses = np.zeros(K, np.float32)
        self.grads = np.full((K, stride), np.nan, np.float32)
        self.calls = [0] * K
        self.rounds = 0

    def evaluate(self):
        self.rounds += 1
        for k, f in enumerate(self.funcs):
            if self.active[k]:
                n = self.sizes[k]
                x = torch.from_numpy(self.X[k, :n].copy()).requires_grad_(True)
                L = f(x)
                L.backward()
                self.losses[k] = L.item()
                self.grads[k, :n] = x.grad.numpy()
                self.calls[k] += 1


def start_points():
    torch.manual_seed(0)
    return [0.5 * torch.randn(n) for n in SIZES]


def test_one_step_matches_torch_lbfgs():
    x0 = start_points()
    ev = Evaluator(FUNCS, SIZES)
    bp, bL, steps, change, rounds = LBFGS_optimization_lockstep([t.numpy() for t in x0], ev, nmax=1, tol=1e-6)
    assert rounds == ev.rounds == max(ev.calls)             # lock step: rounds = the slowest frame's evaluations
    for k, f in enumerate(FUNCS):
        cnt = [0]

        def lf(p):
            cnt[0] += 1
            return f(p)
        p, Lb, st, ch = LBFGS_optimization([x0[k]], lf, nmax=1, tol=1e-6)
        assert st == steps[k] == 1
        assert abs(cnt[0] - ev.calls[k]) <= 1, (k, cnt[0], ev.calls[k])
        assert np.abs(p[0].numpy() - bp[k]).max() < 2e-5 * max(1.0, np.abs(bp[k]).max()), k
        assert abs(Lb - bL[k]) <= 1e-5 * max(1.0, abs(Lb)), k
        assert abs(float(ch) - change[k]) < 1e-4 * max(1.0, float(ch)), k


def test_converges_like_torch_over_several_steps():
    x0 = start_points()
    ev = Evaluator(FUNCS, SIZES)
    bp, bL, steps, change, _ = LBFGS_optimization_lockstep([t.numpy() for t in x0], ev, nmax=10, tol=1e-4)
    for k, f in enumerate(FUNCS):
        p, Lb, st, ch = LBFGS_optimization([x0[k]], f, nmax=10, tol=1e-4)
        assert abs(Lb - bL[k]) < 1e-4 * max(1.0, abs(Lb)) + 1e-5, (k, Lb, bL[k])
        assert abs(st - steps[k]) <= 1


def test_frames_are_independent():
    """A frame's iterates do not depend on which other frames run beside it (bit for bit)."""
    x0 = start_points()
    ev = Evaluator(FUNCS, SIZES)
    bp, *_ = LBFGS_optimization_lockstep([t.numpy() for t in x0], ev, nmax=2, tol=1e-6)
    for k in (0, 2):
        ev1 = Evaluator([FUNCS[k]], [SIZES[k]])
        bp1, *_ = LBFGS_optimization_lockstep([x0[k].numpy()], ev1, nmax=2, tol=1e-6)
        assert np.array_equal(bp1[0], bp[k])


def test_plain_steps_and_bad_usage():
    opt = LockstepLBFGS([3], stride=4)
    opt.set_x(0, np.array([1.0, -2.0, 0.5], np.float32))
    opt.reset(0, line_search=False)                       # fixed unit steps (restart mode of tools/optim.py:77)
    ev = Evaluator([bowl], [3])
    ev.X = np.zeros((1, 4), np.float32)
    ev.grads = np.zeros((1, 4), np.float32)
    r = opt.step(np.ones(1, np.uint8), ev.evaluate, ev.X, ev.active, ev.losses, ev.grads)
    st = opt.stats(0)
    assert r == st["func_evals"] and 1 <= st["n_iter"] <= 20
    assert st["best"] < float(bowl(torch.tensor([1.0, -2.0, 0.5])))
    assert opt.get_x(0).shape == (3,)


def test_nan_loss_takes_the_fallback_path(capsys):
    """Divergence guard of the driver (tools/optim.py:59-77): a NaN closure value ends the step, the best point so far is
    kept and the frame restarts without line search."""
    calls = [0]

    def bad(x):
        calls[0] += 1
        return bowl(x) if calls[0] < 4 else bowl(x) * float("nan")
    ev = Evaluator([bad, bowl], [4, 4])
    x0 = [np.array([1.0, 2.0, -1.0, 0.5], np.float32)] * 2
    bp, bL, steps, change, _ = LBFGS_optimization_lockstep(x0, ev, nmax=2, tol=1e-6)
    assert np.isfinite(bL[0]) and np.isfinite(bp[0]).all()
    assert "NaN" in capsys.readouterr().out
    assert isinstance(change[0], str) and isinstance(change[1], float)
