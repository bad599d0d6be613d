"""Generic explicit integrators over tuples of tensors, same contract as the reference's tools/integrators.py:
``traj = Integrator(ODESystem, x0_tuple, nt, deltat)`` returns the list of nt+1 state tuples.

These Python loops are only used for user-supplied ODE systems; LDDMMModel.Shoot runs the same two schemes
fused on the device (diff_icp_b200/shooting.py).
"""


def _advance(state, *terms):
    """state + sum_k coef_k * deriv_k, component-wise over the tuple."""
    out = []
    for i, s in enumerate(state):
        acc = s
        for coef, deriv in terms:
            acc = acc + coef * deriv[i]
        out.append(acc)
    return tuple(out)


def EulerIntegrator(ODESystem, x0, nt=11, deltat=1.0):
    """x_{n+1} = x_n + dt f(x_n)   (reference: tools/integrators.py:20-31)."""
    dt = deltat / nt
    cur = tuple(t.clone() for t in x0)
    traj = [cur]
    for _ in range(nt):
        cur = _advance(cur, (dt, ODESystem(*cur)))
        traj.append(cur)
    return traj


def RalstonIntegrator(ODESystem, x0, nt=11, deltat=1.0):
    """Ralston's 2nd-order scheme: k1 = f(x), k2 = f(x + 2dt/3 k1), x += dt/4 (k1 + 3 k2)
    (reference: tools/integrators.py:36-51)."""
    dt = deltat / nt
    cur = tuple(t.clone() for t in x0)
    traj = [cur]
    for _ in range(nt):
        k1 = ODESystem(*cur)
        k2 = ODESystem(*_advance(cur, (2 * dt / 3, k1)))
        cur = _advance(cur, (0.25 * dt, k1), (0.75 * dt, k2))
        traj.append(cur)
    return traj
