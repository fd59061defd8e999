"""JPEG-fed batched mode through the C++ pool: in-house decoder vs raw RGB vs nvJPEG.
For each restart interval: decode latency of one image alone (CUDA events around gseg_jpeg_decode-only work via
segment_jpeg minus segment), pool throughput with pinned JPEG bytes in, narrowest labels out.
Usage: python tools/jpeg_pool.py [w h nimg contexts]"""
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

w, h, nimg, S = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1920, 1080, 64, 8)))
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
s0 = gseg.Segmenter(w, h)
imgs = [s0.synth(w, h, 3000 + i) for i in range(nimg)]


def enc(img, rst, sampling, q=90):
    ok, e = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]),
                         [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    return e


def wall(fn, reps=5):
    best = 1e9
    for r in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        if r:
            best = min(best, time.perf_counter() - t0)
    return best


# one image alone: what the decode adds to the latency of a context
s0.set_jpeg_backend(gseg.JPEG_OWN)
raw_ms = wall(lambda: s0.segment(imgs[0], **kw)) * 1e3
print("one image alone, raw RGB from pageable host memory: %.3f ms" % raw_ms, flush=True)
for name, sf in (("4:2:0", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), ("4:4:4", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)):
    for rst in (1, 2, 4, 8, 16, 32, 64, 0):
        e = enc(imgs[0], rst, sf).tobytes()
        ms = wall(lambda: s0.segment_jpeg(e, **kw), 3) * 1e3
        ok = np.array_equal(s0.input_rgb(), cv2.imdecode(np.frombuffer(e, np.uint8), cv2.IMREAD_COLOR)[..., ::-1])
        print("  %s rst %3d: %7d bytes, segment_jpeg %.3f ms (decode adds %.3f ms), pixels == libjpeg: %s" %
              (name, rst, len(e), ms, ms - raw_ms, ok), flush=True)
s0.close()

# pool throughput
out = torch.empty((nimg, h * w), dtype=torch.int32).pin_memory()
raw = torch.from_numpy(np.stack(imgs)).pin_memory()
for S_ in (S,):
    pool = batch.Pool(gseg, w, h, contexts=S_, max_connectivity=4, caps=gseg.CAP_JPEG)
    jobs = pool.jobs([raw[i] for i in range(nimg)], [out[i] for i in range(nimg)], **kw)
    t = wall(lambda: pool.run(jobs))
    print("pool %2d contexts, raw RGB (pinned):              %.3f ms/image  %8.1f Mpixel/s  (%.2f MB/image in)" %
          (S_, t / nimg * 1e3, nimg * w * h / 1e6 / t, w * h * 3 / 1e6), flush=True)
    for name, sf in (("4:2:0", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), ("4:4:4", cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444)):
        for rst in (0, 1, 2, 4, 8, 16):
            encs = [enc(im, rst, sf) for im in imgs]
            tot = sum((e.size + 63) // 64 * 64 for e in encs)
            hj = torch.empty(tot, dtype=torch.uint8).pin_memory()
            items, o = [], 0
            for e in encs:
                hj.numpy()[o:o + e.size] = e.reshape(-1)
                items.append(batch.Jpeg(hj[o:o + e.size], e.size))
                o += (e.size + 63) // 64 * 64
            jobs = pool.jobs(items, [out[i] for i in range(nimg)], **kw)
            t = wall(lambda: pool.run(jobs))
            print("pool %2d contexts, JPEG %s rst %2d (in-house):      %.3f ms/image  %8.1f Mpixel/s  (%.2f MB/image in)" %
                  (S_, name, rst, t / nimg * 1e3, nimg * w * h / 1e6 / t, sum(e.size for e in encs) / nimg / 1e6), flush=True)
    pool.close()
