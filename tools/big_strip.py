"""One gigapixel-config strip (BASELINE.json configs[4]: 32768x32768 tiled in 8 strips of 32768x4096 =
2^27 pixels) on ONE GPU: capacity, timing, and bit-exact partition parity against the CPU oracle.
Usage: python tools/big_strip.py [w h conn variant check]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
gseg.build()
import numpy as np, torch
w, h, conn, variant, check = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (32768, 4096, 4, 0, 1)))
t0 = time.perf_counter()
seg = gseg.Segmenter(w, h, max_connectivity=conn)
dimg = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
seg.synth(w, h, 5, out=dimg)
torch.cuda.synchronize()
print("context + synthetic image: %.1f s, device memory in use %.1f GB" % (time.perf_counter() - t0, (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2**30), flush=True)
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant)
ts = []
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg.segment(dimg, **kw)
    ts.append(time.perf_counter() - t0)
print("%dx%d (%.1f Mpixel) conn %d variant %d: %.1f ms best of 3 -> %.0f Mpixel/s; %d components, %d rounds" %
      (w, h, w * h / 1e6, conn, variant, min(ts) * 1e3, w * h / 1e6 / min(ts), seg.num_components(), len(seg.stats())), flush=True)
print("rounds (V, E, merged, phase):", seg.stats(), flush=True)
lab = seg.labels()
cnt = np.bincount(lab.reshape(-1))
assert cnt.sum() == w * h and cnt.min() > 0 and len(cnt) == seg.num_components()
if variant == 0:
    assert cnt.min() >= 20
print("labels dense, sizes sum to V, min component size %d" % cnt.min(), flush=True)
if check:
    from oracle import oracle as O
    img = dimg.cpu().numpy()
    t0 = time.perf_counter()
    ref = O.segment(img, 0.8, 300.0, 20, conn, variant, max_rounds=48)
    print("CPU oracle: %.1f s" % (time.perf_counter() - t0), flush=True)
    a, na = O.canon(lab)
    b, nb = O.canon(ref[0] if isinstance(ref, tuple) else ref)
    print("partition identical to the CPU oracle:", bool(na == nb and np.array_equal(a, b)), "(%d components)" % nb, flush=True)
