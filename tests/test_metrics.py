"""ASA / UE (Report.pdf p6 eq. 1-2) on hand-checkable label images."""
import importlib

import numpy as np

PKG = "graph-algorithm-image-segmentation-gpgpu_b200"


def test_asa_ue_hand_cases():
    m = importlib.import_module(PKG + ".metrics")
    gt = np.array([[0, 0, 1, 1], [0, 0, 1, 1]])
    assert m.asa(gt, gt) == 1.0 and m.undersegmentation_error(gt, gt) == 0.0
    # label renaming and over-segmentation do not hurt either score
    over = np.array([[5, 7, 9, 9], [5, 7, 2, 2]])
    assert m.asa(over, gt) == 1.0 and m.undersegmentation_error(over, gt) == 0.0
    # one segment covering everything: best overlap 4 of 8; leakage min(4, 4) for both regions
    one = np.zeros_like(gt)
    assert m.asa(one, gt) == 0.5 and m.undersegmentation_error(one, gt) == 1.0
    # a segment leaking one pixel across the boundary
    leak = np.array([[0, 0, 0, 1], [0, 0, 1, 1]])
    assert m.asa(leak, gt) == 7 / 8
    assert m.undersegmentation_error(leak, gt) == (1 + 1) / 8  # segment 0: min(4,1) + min(1,4) = 2
