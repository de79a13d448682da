"""One point of the config-4 crop sweep (tools/bench_configs.py: config4), for profiling a single side range:

    python tools/crop_sweep.py LO HI [T] [ROIS]      # e.g. 448 900 224 16384 = every ROI takes the any-tap-count class
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import bench_configs as bc      # noqa: E402

if __name__ == '__main__':
    a = [int(v) for v in sys.argv[1:]]
    lo, hi = a[0], a[1]
    T = a[2] if len(a) > 2 else 224
    R = a[3] if len(a) > 3 else 16384
    print(json.dumps(bc.config4(lo, hi, T, R)))
