"""Randomised parity sweep: random sizes / parameters / variants / schedules against the CPU oracle.
Usage: python tools/fuzz.py [n_cases] [seed]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
from oracle import oracle as O
import numpy as np
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
seg = gseg.Segmenter(1200, 1200)
bad = compactions = dedups = 0
for case in range(n_cases):
    if rng.random() < 0.15:
        w, h = int(rng.integers(1, 1200)), int(rng.integers(1, 40))
        if rng.random() < 0.5:
            w, h = h, w
    else:
        w, h = int(rng.integers(1, 700)), int(rng.integers(1, 700))
    kind = rng.integers(0, 4)
    if kind == 0:
        img = O.synth(w, h, int(rng.integers(1, 1 << 30)))
    elif kind == 1:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == 2:
        img = (rng.integers(0, 4, (h, w, 3)) * 60).astype(np.uint8)       # few levels: many exact ties
    else:
        img = np.repeat(np.repeat(rng.integers(0, 256, (h // 9 + 1, w // 9 + 1, 3), dtype=np.uint8), 9, 0), 9, 1)[:h, :w].copy()
    conn = int(rng.choice([4, 8]))
    variant = int(rng.choice([0, 0, 1, 2]))
    sigma = float(rng.choice([0.0, 0.5, 0.8, 1.3, 2.4]))
    k = float(rng.choice([0.0, 1.0, 30.0, 300.0, 3000.0]))
    ms = int(rng.choice([0, 1, 2, 20, 200]))
    flags = int(rng.choice([0, 0, 1]))
    tail = [(262144, 65536), (0, 0), (2000, 300)][int(rng.integers(0, 3))]
    # every fourth case: a context sized exactly to the image (arena, edge list and scratch at their tightest), noise input,
    # small k -- the slow-converging predicate rounds that can outgrow the map arena and go through its compaction
    matched = case % 4 == 3
    if matched:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        k, variant = float(rng.choice([0.0, 1.0, 30.0])), 0
    s = gseg.Segmenter(w, h) if matched else seg
    s.set_tail(*tail)
    s.set_blocks_per_sm(int(rng.choice([1, 2, 4])))
    s.set_dedup(int(rng.integers(0, 2)), int(rng.choice([1, 8192])), int(rng.choice([1, 8])), int(rng.choice([65536, 4096, 300])))
    s.segment(np.ascontiguousarray(img), sigma=sigma, k=k, min_size=ms, connectivity=conn, variant=variant, flags=flags)
    got = s.labels()
    compactions += s.compaction_count() if matched else 0
    dedups += len(s.dedup_rounds())
    if matched:
        s.close()
    ref, n = O.segment(np.ascontiguousarray(img), sigma, k, ms, conn, variant, max_rounds=48)
    a, na = O.canon(got); b, nb = O.canon(ref)
    ok = na == nb and np.array_equal(a, b)
    if not ok:
        bad += 1
        print("MISMATCH case %d: %dx%d kind %d conn %d variant %d sigma %.1f k %.0f ms %d flags %d tail %s: %d vs %d components" %
              (case, w, h, kind, conn, variant, sigma, k, ms, flags, tail, na, nb), flush=True)
print("%d cases (%d on capacity-matched contexts), %d mismatches; %d arena compactions, %d sort-based duplicate eliminations ran" %
      (n_cases, n_cases // 4, bad, compactions, dedups))
