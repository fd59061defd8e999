"""Adversarial full-size inputs (constant, noise, checkerboard, stripes, ramp) against the CPU oracle."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
from oracle import oracle as O
import numpy as np
w, h = 1920, 1080
rng = np.random.default_rng(1)
yy, xx = np.mgrid[0:h, 0:w]
imgs = {
    "constant": np.full((h, w, 3), 128, np.uint8),
    "noise": rng.integers(0, 256, (h, w, 3), dtype=np.uint8),
    "checker1": np.repeat((((xx + yy) & 1) * 255).astype(np.uint8)[..., None], 3, 2),
    "stripes": np.repeat((((xx // 7) & 1) * 200).astype(np.uint8)[..., None], 3, 2),
    "ramp": np.stack([(xx * 255 // (w - 1)), (yy * 255 // (h - 1)), ((xx + yy) % 256)], -1).astype(np.uint8),
}
seg = gseg.Segmenter(w, h)
ok = True
for name, img in imgs.items():
    img = np.ascontiguousarray(img)
    for conn, variant in [(4, 0), (8, 0), (8, 1), (4, 2)]:
        t0 = time.perf_counter()
        seg.segment(img, sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant)
        dt = time.perf_counter() - t0
        ref, n = O.segment(img, 0.8, 300.0, 20, conn, variant, max_rounds=48)
        a, na = O.canon(seg.labels()); b, nb = O.canon(ref)
        same = bool(na == nb and np.array_equal(a, b))
        ok &= same
        print("%-9s conn %d variant %d: %7.2f ms, %7d components, %2d rounds, identical to oracle: %s" %
              (name, conn, variant, dt * 1e3, seg.num_components(), len(seg.stats()), same), flush=True)
print("ALL OK" if ok else "MISMATCH")
