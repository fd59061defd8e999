"""Segmentation quality on the synthetic generator's own ground truth (CPU only: the oracle; the GPU partitions are
bit-identical to the oracle's, tests/test_gpu_parity.py).  ASA and undersegmentation error (Report.pdf p6 eq.1-2) of
  - Kruskal Felzenszwalb            (felzenswlab_baseline, BASELINE configs[0])      k = 300, min_size = 20
  - Boruvka Felzenszwalb            (cuda-mst-naive / the engine's GSEG_FELZ)        k = 300, min_size = 20
  - segmentation hierarchy, level 4 (fastmst_segment / GSEG_HIER)
  - superpixel hierarchy, level 4   (superpixel_gpu / GSEG_SUPERPIX)
against the Voronoi regions the images were generated from, next to the medians the report reads off BSDS500
(Report.pdf p6 Fig.4: CPU 0.97 / 0.05, Atomic 0.90 / 0.19, DPP segmentation 0.92 / 0.14, DPP superpixel 0.93 / 0.14).
A sanity anchor for the semantics the build had to choose, not a parity claim: other images.
Usage: python tools/quality.py [n_images] [w h]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

metrics = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200.metrics")


def true_regions(w, h, seed):
    """Region map of O.synth(w, h, seed): the nearest of the jittered 64-pixel grid seeds (same hash as the generator)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def sm64(x):
        with np.errstate(over="ignore"):
            z = (x + np.uint64(0x9E3779B97F4A7C15)) & M
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            return z ^ (z >> np.uint64(31))

    def hash2(a, b):
        with np.errstate(over="ignore"):
            return sm64(sm64(np.uint64(seed) ^ (a * np.uint64(0xD6E8FEB86659FD93))) + np.uint64(b))

    ys, xs = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    cx, cy = xs >> 6, ys >> 6
    bestd = np.full((h, w), np.iinfo(np.int64).max, np.int64)
    best = np.zeros((h, w), np.uint64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            ccx, ccy = cx + dx, cy + dy
            cell = ((ccy + 1).astype(np.uint64) << np.uint64(20)) | (ccx + 1).astype(np.uint64)
            hs = hash2(cell, 1)
            sx = ccx * 64 + (hs & np.uint64(63)).astype(np.int64)
            sy = ccy * 64 + ((hs >> np.uint64(6)) & np.uint64(63)).astype(np.int64)
            d = (xs - sx) ** 2 + (ys - sy) ** 2
            m = d < bestd
            bestd[m] = d[m]
            best[m] = cell[m]
    return np.unique(best, return_inverse=True)[1].reshape(h, w)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (481, 321)   # BSDS500's image size
    rows = {k: [] for k in ("kruskal", "boruvka", "kruskal_k80", "boruvka_k80", "boruvka_k10", "hier_l4", "superpix_l4")}
    agree = []
    for i in range(n):
        seed = 4000 + i
        img = O.synth(w, h, seed)
        gt = true_regions(w, h, seed)
        kr = O.pipeline(img, 0.8, 300.0, 20, 8, O.KRUSKAL)["labels"]
        bo = O.pipeline(img, 0.8, 300.0, 20, 8, O.FELZ)["labels"]
        hi = O.pipeline(img, 0.8, 0.0, 0, 8, O.HIER, max_levels=64)
        sp = O.pipeline(img, 0.8, 0.0, 0, 8, O.SUPERPIX, max_levels=64)
        extra = (("kruskal_k80", O.pipeline(img, 0.8, 80.0, 20, 8, O.KRUSKAL)["labels"]),
                 ("boruvka_k80", O.pipeline(img, 0.8, 80.0, 20, 8, O.FELZ)["labels"]),
                 ("boruvka_k10", O.pipeline(img, 0.8, 10.0, 20, 8, O.FELZ)["labels"]))
        for key, lab in extra + (("kruskal", kr), ("boruvka", bo), ("hier_l4", hi["levels"][min(3, hi["nlevels"] - 1)]),
                         ("superpix_l4", sp["levels"][min(3, sp["nlevels"] - 1)])):
            rows[key].append((metrics.asa(lab, gt), metrics.undersegmentation_error(lab, gt), len(np.unique(lab))))
        agree.append((metrics.asa(bo, kr), metrics.asa(kr, bo)))
    print("# %d synthetic %dx%d images (seeds 4000..), sigma 0.8, 8-connected; ground truth = the generator's Voronoi regions (%d per image)" %
          (n, w, h, len(np.unique(true_regions(w, h, 4000)))))
    print("%-14s %10s %10s %12s   %s" % ("variant", "ASA median", "UE median", "#segments", "Report.pdf p6 Fig.4 (BSDS500, read off the box plots)"))
    ref = {"kruskal": "(k = 300)", "boruvka": "(k = 300)", "kruskal_k80": "CPU 0.97 / 0.05  (K = 80)", "boruvka_k80": "Atomic 0.90 / 0.19  (K = 80)",
           "boruvka_k10": "(k = 10)", "hier_l4": "DPP segmentation 0.92 / 0.14", "superpix_l4": "DPP superpixel 0.93 / 0.14"}
    for key, v in rows.items():
        a = np.array(v)
        print("%-14s %10.4f %10.4f %12d   %s" % (key, np.median(a[:, 0]), np.median(a[:, 1]), int(np.median(a[:, 2])), ref[key]))
    ag = np.array(agree)
    print("Boruvka vs Kruskal partitions (same k, min_size): ASA(Boruvka | Kruskal) median %.4f, ASA(Kruskal | Boruvka) median %.4f "
          "-- different partitions, as the report says (p6)" % (np.median(ag[:, 0]), np.median(ag[:, 1])))
    print("Reading: round-synchronous Boruvka lets every pixel take its lightest edge while k/|C| is still huge, so the blurred\n"
          "transition pixels of a sharp synthetic boundary join a region early and raise its Int(C) to ~0.23 x the boundary's\n"
          "contrast; the region then accepts lighter crossings to other regions: under-segmentation that grows with k.  Kruskal\n"
          "meets the same edges late, with tight thresholds, and leaves them as slivers (over-segmentation).  The report sees the\n"
          "same direction on BSDS500 (Boruvka + predicate less accurate than Kruskal); natural images have no such ramps, so its\n"
          "gap is smaller.  The hierarchy variants (no predicate, level 4) are accurate on this data.")


if __name__ == "__main__":
    main()
