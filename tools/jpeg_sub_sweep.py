"""Marker-less JPEG decode in the pool: sweep of the sub-sequence size (GSEG_JPEG_SUB) and of the tail cluster size
(GSEG_TAIL_CLUSTER), cluster kernel vs grid of small blocks with a software barrier (GSEG_JPEG_SYNC).  Usage: python tools/jpeg_sub_sweep.py [nimg]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "graph-algorithm-image-segmentation-gpgpu_b200"
gseg = importlib.import_module(PKG)
batch = importlib.import_module(PKG + ".batch")
import cv2
import numpy as np
import torch

w, h, nimg = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 64
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
s0 = gseg.Segmenter(w, h)
imgs = [s0.synth(w, h, 3000 + i) for i in range(nimg)]
s0.close()
out = torch.empty((nimg, h * w), dtype=torch.int32).pin_memory()


def files(sf):
    encs = [cv2.imencode(".jpg", np.ascontiguousarray(im[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])[1] for im in imgs]
    tot = sum((e.size + 63) // 64 * 64 for e in encs)
    hj = torch.empty(tot, dtype=torch.uint8).pin_memory()
    items, o = [], 0
    for e in encs:
        hj.numpy()[o:o + e.size] = e.reshape(-1)
        items.append(batch.Jpeg(hj[o:o + e.size], e.size))
        o += (e.size + 63) // 64 * 64
    return items


def wall(fn, reps=4):
    best = 1e9
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        if r:
            best = min(best, time.perf_counter() - t0)
    return best


sets = {"4:2:0": files(cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), "4:4:4": files(cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)}
for sub, tc, mode in [(128, None, "cluster"), (128, None, "grid"), (64, None, "grid"), (96, None, "grid"), (192, None, "grid"), (256, None, "grid"), (128, None, "cluster")]:
    os.environ["GSEG_JPEG_SUB"] = str(sub)
    os.environ["GSEG_JPEG_SYNC"] = mode
    if tc is None:
        os.environ.pop("GSEG_TAIL_CLUSTER", None)
    else:
        os.environ["GSEG_TAIL_CLUSTER"] = str(tc)
    pool = batch.Pool(gseg, w, h, contexts=8, max_connectivity=4, caps=gseg.CAP_JPEG)
    line = "%-7s sub-sequence %4d bytes, tail cluster %s:" % (mode, sub, tc if tc else "8 (pool default)")
    for name, items in sets.items():
        jobs = pool.jobs(items, [out[i] for i in range(nimg)], **kw)
        t = wall(lambda: pool.run(jobs))
        line += "  %s %.3f ms/image %7.1f Mpixel/s" % (name, t / nimg * 1e3, nimg * w * h / 1e6 / t)
    print(line, flush=True)
    pool.close()
