import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
from oracle import oracle as O
import numpy as np
w, h, conn, variant, flags = (int(x) for x in sys.argv[1:6])
seg = gseg.Segmenter(max(w, 64), max(h, 64))
img = O.synth(w, h, 100 + w)
print("start", w, h, conn, variant, flags, flush=True)
seg.segment(img, sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant, flags=flags)
print("done; stats", seg.stats(), flush=True)
ref = O.pipeline(img, 0.8, 300.0, 20, conn, variant)
a, na = O.canon(seg.labels()); b, nb = O.canon(ref["labels"])
print("match", na == nb and np.array_equal(a, b), flush=True)
