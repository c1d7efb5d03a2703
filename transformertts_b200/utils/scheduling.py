"""Host-side scalar schedules (reference: utils/scheduling.py:11-47)."""
from __future__ import annotations

import numpy as np


def piecewise_linear(step, X, Y):
    """Value at `step` of the piecewise-linear function through the points (X_i, Y_i); constant outside."""
    assert len(X) == len(Y)
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if step < X[0]:
        return float(Y[0])
    if step >= X[-1]:
        return float(Y[-1])
    i = int(np.searchsorted(X, step, side='right') - 1)
    frac = (step - X[i]) / (X[i + 1] - X[i])
    return float(Y[i] + frac * (Y[i + 1] - Y[i]))


def piecewise_linear_schedule(step, schedule) -> float:
    """Learning rate at `step` for a [[step, lr], ...] schedule (reference: utils/scheduling.py:31-36)."""
    s = np.asarray(schedule, dtype=np.float64)
    return float(np.float32(piecewise_linear(step, s[:, 0], s[:, 1])))


def reduction_schedule(step, schedule) -> int:
    """Step-wise reduction factor (Aligner only; reference: utils/scheduling.py:39-47)."""
    s = np.asarray(schedule)
    r = s[0, 1]
    for start, val in s:
        if start <= step:
            r = val
        else:
            break
    return int(r)
