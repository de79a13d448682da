"""oracle.lsap_spec (Crouse restatement) pinned against the installed scipy."""
import numpy as np
from scipy.optimize import linear_sum_assignment

from oracle.lsap_spec import lsap


def _instances(rng, n):
    for t in range(n):
        nr, nc = rng.integers(1, 16, 2)
        mode = t % 5
        if mode == 0:
            C = rng.random((nr, nc))
        elif mode == 1:
            C = rng.integers(0, 4, (nr, nc)).astype(float)            # heavy exact ties
        elif mode == 2:
            C = rng.integers(0, 3, (nr, nc)).astype(float)
            C[rng.random((nr, nc)) < 0.3] = 9999                       # sentinel ties
        elif mode == 3:
            C = rng.random((nr, nc)).astype(np.float32)                # duplicate rows / columns
            if nr > 2:
                C[1] = C[0]
            if nc > 2:
                C[:, 2] = C[:, 0]
        else:
            C = np.abs(rng.normal(size=(nr, nc))).astype(np.float32) * 100
        yield C


def test_matches_scipy_including_ties():
    rng = np.random.default_rng(7)
    for C in _instances(rng, 1200):
        a, b = linear_sum_assignment(C)
        a2, b2 = lsap(C)
        assert np.array_equal(a, a2) and np.array_equal(b, b2), C


def test_tall_flattened_shape():
    """The shape match_objects produces: (N*M) x P with N*M > P -> transposed inside."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        N, M, P = rng.integers(2, 7, 3)
        C = (rng.random((N * M, P)) * 50).astype(np.float32)
        a, b = linear_sum_assignment(C)
        a2, b2 = lsap(C)
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
        assert np.all(np.diff(a2) > 0)                                # ascending r, as match_objects iterates
