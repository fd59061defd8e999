"""A short JPEG-fed pool run for an ncu launch list: N 1080p files (4:2:0, quality 90, restart interval RST) through
gseg_pool_run, REPS times.  Usage: python tools/jpeg_launches.py [nimg rst reps]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "graph-algorithm-image-segmentation-gpgpu_b200"
gseg = importlib.import_module(PKG)
batch = importlib.import_module(PKG + ".batch")
import cv2
import numpy as np
import torch

nimg, rst, reps = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (16, 1, 2)))
w, h = 1920, 1080
s0 = gseg.Segmenter(w, h)
imgs = [s0.synth(w, h, 3000 + i) for i in range(nimg)]
s0.close()
encs = [cv2.imencode(".jpg", np.ascontiguousarray(im[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])[1] for im in imgs]
hj = [torch.from_numpy(e.reshape(-1).copy()).pin_memory() for e in encs]
out = torch.empty((nimg, h * w), dtype=torch.int32).pin_memory()
pool = batch.Pool(gseg, w, h, contexts=8, max_connectivity=4, caps=gseg.CAP_JPEG)
jobs = pool.jobs([batch.Jpeg(b, b.numel()) for b in hj], [out[i] for i in range(nimg)], sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
for _ in range(reps):
    res = pool.run(jobs)
torch.cuda.synchronize()
print("ok", [int(r.n_components) for r in res][:4])
pool.close()
