"""L-BFGS driver used by LDDMMModel.Optimize: host-side control flow, same behaviour as the reference's
``LBFGS_optimization`` (tools/optim.py:10-110):

* ``torch.optim.LBFGS(max_iter=20, max_eval=100, history_size=100, line_search_fn="strong_wolfe")`` (:26);
* at most ``nmax`` optimizer steps, stop when the RMS change of every parameter is below ``tol`` x its RMS (:98-104);
* divergence guard: NaN / loss above ``errthresh`` / loss increase -> fall back to the best parameters seen so far,
  or perturb by 1 % noise, and restart L-BFGS *without* line search (:60-97);
* returns the BEST parameters seen over all closure evaluations, not the last ones (:108-110).

It stays on the host on purpose (parity of iterates with the reference); every closure evaluation is one fused
shoot + one adjoint sweep on the device, and one ``.item()`` synchronisation as in the reference (:39).
"""

import math

import torch


def LBFGS_optimization(p0, lossfunc, nmax=10, tol=1e-3, errthresh=1e8, lossgrad=None):
    """`lossgrad` (optional, B200 build): callable returning (loss, [gradients]) directly -- used by LDDMMModel.Optimize when
    the whole closure (shoot + quadratic loss + adjoint) runs as one captured launch sequence; the optimiser then sees
    exactly the same loss / gradient values as through `lossfunc` + `.backward()`, without building an autograd graph."""
    params = [t.clone().contiguous().detach().requires_grad_(True) for t in p0]

    def new_optimizer(line_search):
        return torch.optim.LBFGS(params, max_iter=20, max_eval=100, history_size=100, line_search_fn=line_search)

    optimizer = new_optimizer("strong_wolfe")
    history = []
    best = {"L": math.inf, "p": None}

    def closure():
        optimizer.zero_grad()
        if lossgrad is not None:
            val, grads = lossgrad(*[t.detach() for t in params])
            val = float(val)
            loss = torch.tensor(val)
        else:
            loss = lossfunc(*params)
            val = loss.detach().item()
        history.append(val)
        if val < best["L"]:
            best["L"] = val
            best["p"] = [t.clone().detach() for t in params]
        if lossgrad is not None:
            for t, g in zip(params, grads):
                t.grad = g.to(device=t.device, dtype=t.dtype).clone()
        else:
            loss.backward()
        return loss

    step, go_on, L, change = 0, True, math.inf, None
    while step < nmax and go_on:
        step += 1
        before = [t.clone().detach() for t in params]
        optimizer.step(closure)
        L_before, L = L, history[-1]

        if L > L_before or L > errthresh or math.isnan(L):
            if math.isnan(L):
                print("WARNING: NaN value for loss L during L-BFGS optimization.")
            elif L > errthresh:
                print("WARNING: Aberrantly large value for loss L during L-BFGS optimization.")
            else:
                print("WARNING: Increase of loss L during L-BGFS optimization.")
            if best["L"] < L_before:
                params = [t.clone() for t in best["p"]]
                L = best["L"]
                print("L-BFGS optimization. Found an intermediate 'best_p' value for this iteration.")
            else:
                rmod = 0.01
                params = [t + rmod * t.std() * torch.randn(t.shape, dtype=t.dtype, device=t.device) for t in best["p"]]
                L = lossfunc(*params) if lossgrad is None else lossgrad(*params)[0]
                print(f"L-BFGS optimization. Trying a random perturbation of parameter from its current value, with relative strength {rmod}.")
            change = "None (divergent iteration step)"
            params = [t.requires_grad_(True) for t in params]
            optimizer = new_optimizer(None)
        else:
            deltas = [((t - b) ** 2).mean().sqrt().detach().cpu().numpy() for t, b in zip(params, before)]
            scales = [(b ** 2).mean().sqrt().detach().cpu().numpy() for b in before]
            go_on = any(dl > tol * sc for dl, sc in zip(deltas, scales))
            change = max(deltas)

    return [t.detach() for t in best["p"]], best["L"], step, change
