"""Host-side scalar schedules (reference: utils/scheduling.py:5-47).  Checked bit-for-bit against the reference module in
tests/test_reference_shim.py::test_scheduling_bitwise."""
from __future__ import annotations

import numpy as np


def piecewise_linear(step, X, Y):
    """Value at `step` of the piecewise-linear function through the points (X_i, Y_i); constant outside.  Inside a segment
    the reference evaluates slope * step + intercept (utils/scheduling.py:5-8) -- kept, so the float64 result is identical."""
    assert len(X) == len(Y)
    X = np.asarray(X)
    Y = np.asarray(Y)
    if step < X[0]:
        return Y[0]
    i = int(np.nonzero(step >= X)[0][-1])
    if i == len(Y) - 1:
        return Y[-1]
    slope = (Y[i + 1] - Y[i]) / (X[i + 1] - X[i])
    intercept = Y[i] - slope * X[i]
    return slope * step + intercept


def piecewise_linear_schedule(step, schedule) -> float:
    """Learning rate at `step` for a [[step, lr], ...] schedule, rounded to float32 as the reference's tf.cast does
    (utils/scheduling.py:31-36)."""
    s = np.array(schedule)
    return float(np.float32(piecewise_linear(step, s[:, 0], s[:, 1])))


def reduction_schedule(step, schedule) -> int:
    """Step-wise reduction factor (Aligner only; reference: utils/scheduling.py:39-47).  Like the reference, a step below the
    first breakpoint yields the first row's FIRST column (its `r = schedule[0, 0]` initialisation)."""
    s = np.array(schedule)
    r = s[0, 0]
    for start, val in s:
        if start <= step:
            r = val
        else:
            break
    return int(r)
