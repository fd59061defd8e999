"""Segmentation quality scores of the reference's evaluation: achievable segmentation accuracy (ASA)
and undersegmentation error (UE) of a label image against a ground-truth label image
(Report.pdf p6 section 4.2, eq. 1-2; the reference's `comparetool` branch, README.md:22).

Host-side analysis of label images the engine already produced: numpy
only, nothing here is on the segmentation hot path.

    ASA(S, G) = sum_k max_i |s_k ∩ g_i| / sum_i |g_i|
    UE(S, G)  = sum_i sum_{k : s_k ∩ g_i != ∅} min(|s_k ∩ g_i|, |s_k − g_i|) / sum_i |g_i|
"""
import numpy as np


def _overlap(seg, gt):
    seg = np.asarray(seg).reshape(-1)
    gt = np.asarray(gt).reshape(-1)
    if seg.shape != gt.shape:
        raise ValueError("label images differ in size")
    _, s = np.unique(seg, return_inverse=True)
    _, g = np.unique(gt, return_inverse=True)
    ns, ng = int(s.max()) + 1, int(g.max()) + 1
    pairs, counts = np.unique(s.astype(np.int64) * ng + g, return_counts=True)
    return pairs // ng, pairs % ng, counts, ns, ng, seg.size


def asa(seg, gt):
    """Achievable segmentation accuracy: every segment is labelled with the ground-truth region it overlaps most."""
    sk, _, cnt, ns, _, n = _overlap(seg, gt)
    best = np.zeros(ns, np.int64)
    np.maximum.at(best, sk, cnt)
    return float(best.sum()) / n


def undersegmentation_error(seg, gt):
    """Leakage of segments across ground-truth boundaries: min(inside, outside) of every overlapping pair."""
    sk, _, cnt, ns, _, n = _overlap(seg, gt)
    size = np.zeros(ns, np.int64)
    np.add.at(size, sk, cnt)
    return float(np.minimum(cnt, size[sk] - cnt).sum()) / n
