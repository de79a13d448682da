"""CPU oracle for the bpc_baseline match + ROI-crop hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bpc_baseline_b200/`` may import this package; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, always as the checker or the timed CPU baseline, never as a product
path.

What it is: a NumPy restatement of the reference's own algorithm for the path
(``bpc/inference/epipolar_matching.py``, ``bpc/inference/utils/camera_utils.py``,
``bpc/inference/utils/triangulation.py``, ``bpc/inference/process_pose.py:79-94,144-210``,
``bpc/utils/data_utils.py:34-44,383-387``), every function citing the lines it follows.

Third-party arithmetic on the path (not under /root/reference) is restated too and pinned
against the installed libraries in ``tests/``:
  * ``scipy.optimize.linear_sum_assignment`` (reference pins scipy==1.14.0, image has 1.18.1;
    call site epipolar_matching.py:107)  -> ``oracle.lsap_spec`` (Crouse shortest augmenting path).
  * ``cv2.resize(..., INTER_AREA)`` (reference pins opencv-python==4.8.0.74, image has
    opencv-python-headless 4.13.0; call site data_utils.py:39) -> ``oracle.area_spec``.
  * ``torchvision.transforms.functional.to_tensor/normalize`` (process_pose.py:207-209)
    -> ``oracle.crop.normalise_lut``.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: the
script ``oracle/make_golden.py`` imports the unmodified reference from /root/reference in the
authoring container, runs it on seeded synthetic scenes and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against those files.
"""
