"""Input normalisation for point sets (reference: tools/in_out.py:7-47)."""

import torch


def _is_points(t):
    return isinstance(t, torch.Tensor) and t.dtype == torch.float32


def read_point_sets(x):
    """Accepts a single (N,D) tensor, a list of per-frame tensors, or a list (frames) of lists (structures);
    returns (x[k][s], K, S, D)."""
    if _is_points(x):
        x = [[x]]
    elif isinstance(x, list):
        x = [[xk] for xk in x] if _is_points(x[0]) else [list(xk) for xk in x]
    else:
        raise ValueError("Wrong format for input x")
    K = len(x)
    counts = {len(xk) for xk in x}
    if len(counts) > 1:
        raise ValueError("All frames should have same number of structures")
    S = counts.pop()
    dims = {xks.shape[1] for xk in x for xks in xk}
    if len(dims) > 1:
        raise ValueError("All point sets should have same axis-1 dimension")
    return x, K, S, dims.pop()
