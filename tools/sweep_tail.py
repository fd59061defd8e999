"""Throughput vs tail hand-over thresholds (S contexts in flight). Usage: python tools/sweep_tail.py [S]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
w, h, nimg = 1920, 1080, 32
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
segs = [gseg.Segmenter(w, h) for _ in range(S)]
dimgs = torch.empty((nimg, h, w, 3), dtype=torch.uint8, device="cuda")
for i in range(nimg):
    segs[0].synth(w, h, 2000 + i, out=dimgs[i])
dlab = torch.empty((S, h, w), dtype=torch.int32, device="cuda")
for tE, tV in [(0, 0), (16384, 4096), (65536, 16384), (131072, 65536), (262144, 65536), (524288, 65536), (1 << 20, 1 << 17)]:
    for s in segs:
        s.set_tail(tE, tV)
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
                segs[j].labels(out=dlab[j])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rep >= 2:
            best = min(best, dt)
    print("S=%d tail_E=%7d tail_V=%6d: %.3f ms/image  %.1f Mpixel/s" % (S, tE, tV, best / nimg * 1e3, nimg * w * h / 1e6 / best), flush=True)
