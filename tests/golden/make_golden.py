"""Generates tests/golden/golden_v1.npz from the CPU oracle (oracle/gseg_oracle.c).

The reference ships no golden vectors, so these are *self-generated* fixtures:
they pin the oracle (and through it the CUDA path) against drift, they do not pin it against the
reference.  Run from the repo root:  python -m tests.golden.make_golden
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# (w, h, seed, sigma, k, min_size, conn, variant)  variant: 0 FELZ 1 HIER 2 SUPERPIX 3 KRUSKAL
CASES = [
    (48, 36, 1, 0.8, 300.0, 20, 8, 3),
    (48, 36, 1, 0.8, 300.0, 20, 8, 0),
    (48, 36, 1, 0.8, 300.0, 20, 4, 0),
    (64, 40, 2, 0.8, 30.0, 10, 4, 0),
    (64, 40, 2, 0.8, 3.0, 0, 8, 0),
    (37, 29, 3, 0.5, 0.0, 4, 4, 0),
    (48, 36, 4, 0.8, 0.0, 0, 8, 1),
    (48, 36, 4, 1.2, 0.0, 0, 4, 1),
    (48, 36, 5, 0.8, 0.0, 0, 8, 2),
    (31, 17, 6, 0.8, 0.0, 0, 4, 2),
]


def run_case(O, case):
    w, h, seed, sigma, k, ms, conn, variant = case
    img = O.synth(w, h, seed)
    r = O.pipeline(img, sigma, k, ms, conn, variant, max_levels=32)
    out = {
        "img_sha": np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8).copy(),
        "wts_bits": r["wts"].view(np.uint32).copy(),
        "labels": O.canon(r["labels"])[0].astype(np.int32),
    }
    if variant in (1, 2):
        out["ncomp"] = np.array(r["ncomp"], np.int32)
        out["levels"] = np.stack([O.canon(l)[0] for l in r["levels"]]).astype(np.int32)
    return out


if __name__ == "__main__":
    from oracle import oracle as O
    blob = {}
    for i, case in enumerate(CASES):
        for key, val in run_case(O, case).items():
            blob["c%d_%s" % (i, key)] = val
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")
