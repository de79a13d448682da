"""bpc_baseline_b200 -- B200 (sm_100a) implementation of the bpc_baseline match + ROI-crop hot path.

Layers
  csrc/ + include/bpc_b200.h   hand-written CUDA kernels behind a C ABI (libbpc_b200.so, built in-tree)
  _lib, batched, pipeline      ctypes binding, batched device-resident API, chunked match -> crop pipeline
  inference/, utils/           the reference's Python call surface (same names and arguments) on top of it
  synth                        synthetic IPD-like scenes for tests and bench.py

``install()`` rebinds the hot-path names inside an already imported reference package (``bpc``) so that
existing scripts run on the GPU path unchanged; see INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import sys

__version__ = '0.1.0'

# (reference module, attribute) -> (our module, attribute)
_PATCHES = [
    ('inference.utils.camera_utils', 'compute_fundamental_matrix', 'inference.utils.camera_utils', 'compute_fundamental_matrix'),
    ('inference.utils.triangulation', 'triangulate_multi_view', 'inference.utils.triangulation', 'triangulate_multi_view'),
    ('inference.utils.triangulation', 'compute_reprojection_error', 'inference.utils.triangulation', 'compute_reprojection_error'),
    ('inference.epipolar_matching', 'epipolar_error', 'inference.epipolar_matching', 'epipolar_error'),
    ('inference.epipolar_matching', 'epipolar_error_full', 'inference.epipolar_matching', 'epipolar_error_full'),
    ('inference.epipolar_matching', 'compute_cost_matrix', 'inference.epipolar_matching', 'compute_cost_matrix'),
    ('inference.epipolar_matching', 'match_objects', 'inference.epipolar_matching', 'match_objects'),
    ('inference.epipolar_matching', 'triangulate_multi_view', 'inference.epipolar_matching', 'triangulate_multi_view'),
    ('utils.data_utils', 'letterbox_preserving_aspect_ratio', 'utils.data_utils', 'letterbox_preserving_aspect_ratio'),
    # process_pose bound these with `from ... import` (process_pose.py:23-26): patch its namespace too
    ('inference.process_pose', 'compute_cost_matrix', 'inference.epipolar_matching', 'compute_cost_matrix'),
    ('inference.process_pose', 'match_objects', 'inference.epipolar_matching', 'match_objects'),
    ('inference.process_pose', 'triangulate_multi_view', 'inference.epipolar_matching', 'triangulate_multi_view'),
    ('inference.process_pose', 'compute_fundamental_matrix', 'inference.utils.camera_utils', 'compute_fundamental_matrix'),
    ('inference.process_pose', 'letterbox_preserving_aspect_ratio', 'utils.data_utils', 'letterbox_preserving_aspect_ratio'),
    ('inference.process_pose', 'PosePrediction', 'inference.process_pose', 'PosePrediction'),
]
_saved: list = []


def install(reference_package: str = 'bpc', patch_estimator: bool = True) -> list:
    """Rebind the hot-path functions of the imported reference package to the CUDA implementations.

    Only modules that are already imported are touched.  With ``patch_estimator`` the reference's
    ``PoseEstimator._match`` is replaced by the single-launch batched version as well.  Returns the list of
    ``module.attribute`` names that were rebound; ``uninstall()`` restores them.
    """
    from . import _lib
    _lib.load()                                          # fail now, loudly, if the CUDA library is missing
    done = []
    for ref_mod, ref_attr, our_mod, our_attr in _PATCHES:
        mod = sys.modules.get(f'{reference_package}.{ref_mod}')
        if mod is None or not hasattr(mod, ref_attr):
            continue
        ours = getattr(importlib.import_module(f'{__name__}.{our_mod}'), our_attr)
        _saved.append((mod, ref_attr, getattr(mod, ref_attr)))
        setattr(mod, ref_attr, ours)
        done.append(f'{mod.__name__}.{ref_attr}')
    pp = sys.modules.get(f'{reference_package}.inference.process_pose')
    if patch_estimator and pp is not None and hasattr(pp, 'PoseEstimator'):
        from .inference.process_pose import PoseEstimator as Ours
        ref_cls = pp.PoseEstimator
        _saved.append((ref_cls, '_match', ref_cls._match))

        def _match(self, capture, detections, _impl=Ours._match):
            if not hasattr(self, 'verbose'):
                self.verbose = False
            return _impl(self, capture, detections)
        ref_cls._match = _match
        done.append(f'{pp.__name__}.PoseEstimator._match')
    return done


def uninstall() -> None:
    while _saved:
        obj, attr, old = _saved.pop()
        setattr(obj, attr, old)
