"""Whole-job throughput of a batch of images over S concurrent contexts (one stream each).
Usage: python tools/throughput.py [w h conn variant nimg]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
gseg.build()
import torch

w, h, conn, variant, nimg = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (1920, 1080, 4, 0, 32)))
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant)
s0 = gseg.Segmenter(w, h)
dimgs = torch.empty((nimg, h, w, 3), dtype=torch.uint8, device="cuda")
for i in range(nimg):
    s0.synth(w, h, 2000 + i, out=dimgs[i])
dlab = torch.empty((nimg, h, w), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for S in (1, 2, 3, 4, 6, 8):
    segs = [gseg.Segmenter(w, h) for _ in range(S)]
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for base in range(0, nimg, S):
            n = min(S, nimg - base)
            for j in range(n):
                segs[j].segment(dimgs[base + j], wait=False, **kw)
            for j in range(n):
                segs[j].wait()
                segs[j].labels(out=dlab[base + j])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep >= 2:
            best = min(best, dt)
    print("S=%d contexts: %.3f ms/image  %.1f Mpixel/s  (stats rounds=%d)" %
          (S, best / nimg * 1e3, nimg * w * h / 1e6 / best, len(segs[0].stats())), flush=True)
    for s in segs:
        s.close()
