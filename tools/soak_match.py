"""One-off soak of the matcher against the SciPy oracle on conflict-heavy scenes of many sizes:

    python tools/soak_match.py [scenes_per_case]

Every scene is compared with oracle.geometry.match_scene (indices incl. order, float32 costs bit-equal, X to 1e-9)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from bpc_baseline_b200 import synth                                   # noqa: E402
from tests.test_gpu_fullsize import _check_against_oracle, _match      # noqa: E402

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    total = 0
    for case, (D, p_drop, sigma, n_dup, n_false) in enumerate([(5, 0.4, 3.0, 2, 2), (12, 0.3, 3.0, 3, 3), (24, 0.3, 2.0, 4, 4),
                                                               (33, 0.2, 4.0, 6, 2), (40, 0.35, 3.0, 4, 8), (64, 0.25, 2.5, 8, 8),
                                                               (100, 0.2, 2.0, 10, 10), (150, 0.3, 3.0, 0, 20)]):
        k = max(4, n if D <= 64 else n // 8)
        batch = synth.make_scenes(k, D, p_drop=p_drop, sigma=sigma, n_dup=n_dup, n_false=n_false, seed=synth.SEED + 900 + case)
        out = _match(batch)
        _check_against_oracle(batch, out, range(k))
        total += k
        print(f'D={D} p_drop={p_drop} dup={n_dup} false={n_false}: {k} scenes identical to SciPy', flush=True)
    print(f'SOAK_OK {total} scenes')
