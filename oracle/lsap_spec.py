"""Oracle (test infrastructure): restatement of ``scipy.optimize.linear_sum_assignment``.

The reference calls SciPy's rectangular LSAP at bpc/inference/epipolar_matching.py:107 (pinned
scipy==1.14.0 in docker/requirements.txt:3; this image has 1.18.1).  SciPy's source is not under
/root/reference, so the published algorithm is restated here -- D. F. Crouse, "On implementing 2D
rectangular assignment algorithms", IEEE T-AES 52(4), 2016, as implemented by SciPy's
``rectangular_lsap`` -- including the details that decide WHICH optimum is returned when optima
are not unique (duplicate detections, 9999 sentinels):

  * if there are more rows than columns the problem is transposed;
  * rows are augmented in ascending order; each shortest-augmenting-path search scans the
    not-yet-visited columns in the order held by ``remaining`` (initially nc-1 .. 0, later
    permuted by swap-removal);
  * among equal-lowest columns an unassigned one wins (the last such in scan order), otherwise
    the first in scan order;
  * reduced costs are ``((minVal + C[i,j]) - u[i]) - v[j]`` in float64, evaluated left to right.

``tests/test_lsap_spec.py`` pins this model against the installed SciPy on random rectangular
instances with heavy exact ties.  The CUDA assignment kernel follows the same rules.
"""
from __future__ import annotations

import numpy as np


def lsap(cost):
    """Return (row_ind, col_ind) exactly as scipy.optimize.linear_sum_assignment(cost) does."""
    C = np.asarray(cost, dtype=np.float64)
    nr0, nc0 = C.shape
    if nr0 == 0 or nc0 == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transpose = nc0 < nr0
    if transpose:
        C = C.T.copy()
    nr, nc = C.shape
    u = np.zeros(nr)
    v = np.zeros(nc)
    col4row = np.full(nr, -1, np.int64)
    row4col = np.full(nc, -1, np.int64)
    stats = {'steps': 0, 'max_chain': 0}
    for cur in range(nr):
        spc = np.full(nc, np.inf)
        path = np.full(nc, -1, np.int64)
        SR = np.zeros(nr, bool)
        SC = np.zeros(nc, bool)
        remaining = list(range(nc - 1, -1, -1))
        num_remaining = nc
        min_val = 0.0
        i = cur
        sink = -1
        chain = 0
        while sink == -1:
            index = -1
            lowest = np.inf
            SR[i] = True
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + C[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == np.inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
            chain += 1
        stats['steps'] += chain
        stats['max_chain'] = max(stats['max_chain'], chain)
        u[cur] += min_val
        for i2 in range(nr):
            if SR[i2] and i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if SC[j2]:
                v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur:
                break
    lsap.last_stats = stats
    if transpose:
        order = np.argsort(col4row, kind='stable')
        return col4row[order].astype(np.int64), order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row.astype(np.int64)


lsap.last_stats = {}
